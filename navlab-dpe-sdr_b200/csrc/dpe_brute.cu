// dpe_brute.cu -- the north-star kernel: every (candidate, PRN) pair correlates the
// whole 20 ms block against its own blended C/A replica (SURVEY.md section 8 a').
//
//   v(j,c) = sum_m xw_c[m] * ( (1-a) r_c[(m-k) mod S] + a r_c[(m-k-1) mod S] )
//          = sum_p xw_c[(p+k) mod S] * ( r_c[p] + a d_c[p] ),   d_c[p] = r_c[p-1] - r_c[p]
//
// with integer lag k and fraction a from the candidate's FP64 geometry
// (batchcorrmanifold.cu:1779-1800).  By linearity this equals the reference's
// lerp of two correlogram bins (:1806-1812) to rounding.
//
// Mapping (B200; the kernel is bound by FFMA2 issue -- one per 2 cycles per sub-partition, and
// nothing co-issues with it -- so the design minimises every other instruction):
//   * pairs are bucketed by (PRN, k); a warp ("group") holds 32 pairs of one bucket in
//     registers, a CTA slot = 8 groups of ONE bucket, so the whole CTA shares the lag;
//   * the kernel walks the replica position p.  The replica is staged as (d, d, r, r) position
//     pairs, the samples from the copy of the sample plane that is shifted by k (dpe_prepare.cu:
//     k_sample_planes), so both tiles are aligned on p for every lag: a lane's 8 positions are
//     4 + 4 conflict-free LDS.128 at immediate offsets, and there is no address arithmetic and
//     no subtraction in the loop;
//   * per pair of positions and candidate: 1 FFMA2 (blend, alpha as broadcast operand) +
//     2 FFMA2 (re / im accumulate);
//   * tiles of 1024 positions are staged by 1-D TMA bulk copies (cp.async.bulk + mbarrier
//     full/empty pairs, 4 stages); the planes are stored pre-skewed in HBM;
//   * persistent CTAs (one per SM), 8 warps (2 per scheduler, up to 255 registers); the TMA
//     refill duty rotates over the warps instead of living in a 9th producer warp;
//   * lane partials are combined with warp shuffles.
#include "dpe_geom.cuh"

namespace dpe {

// ---------------------------------------------------------------------------
// The pair sort: three launches (round 1: five).
//   k_pair_bins   bins of every pair + per-CTA (PRN, lag) histograms
//   k_block_scan  per bucket: exclusive scan of the per-CTA counts; the CTA that takes the last
//                 ticket lays the buckets out as groups / slots (bucket scan)
//   k_scatter     group headers + the pairs into their bucket, in candidate order
// It is a stable counting sort (position in the bucket = rank by candidate index), so the
// composition of every group -- and with it the summation order of the slots that a CTA boundary
// of k_brute cuts -- is the same on every run.  No kernel here waits for another CTA (last-ticket
// tails only): with two epochs in flight these kernels share the SMs with a running k_brute.
// ---------------------------------------------------------------------------
template <int SAT_MODE>
__global__ void DPE_SIDE128
k_pair_bins(const double* __restrict__ grid, const EpochDev* __restrict__ ep, const double* __restrict__ sat,
            double fs, int S, int W, int T, int64_t G, int64_t grid_offset, int16_t* __restrict__ pair_k,
            float* __restrict__ pair_a, float2* __restrict__ pair_v, int32_t* __restrict__ blk_hist,
            const SatGeo* __restrict__ geo_tab) {
    extern __shared__ int32_t hs[];
    __shared__ ChanConst cc[DPE_MAX_CHAN];
    __shared__ SatGeo geo_mid[DPE_MAX_CHAN];
    const EpochDev& e = *ep;
    const int NB = 2 * W + 1;
    const int nbuck = e.C * NB;
    for (int i = threadIdx.x; i < nbuck; i += blockDim.x) hs[i] = 0;
    chan_consts(e, fs, cc);
    if (SAT_MODE == DPE_SAT_MIDDLE)
        for (int c = threadIdx.x; c < e.C; c += blockDim.x) geo_mid[c] = make_sat_geo(e, sat + ((size_t)c * T + T / 2) * 8);
    __syncthreads();
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < G) {
        CandRel rel;
        const Cand p = cand_ecef(e, grid + 4 * j, &rel);
        const int it = (SAT_MODE == DPE_SAT_PER_TIME) ? (int)((j + grid_offset) % T) : T / 2;
#pragma unroll 2
        for (int c = 0; c < e.C; ++c) {
            const SatGeo& sg = (SAT_MODE == DPE_SAT_PER_TIME) ? geo_tab[(size_t)c * T + it] : geo_mid[c];
            const double idx = code_index_fast(e, cc[c], sg, p, rel, sat + ((size_t)c * T + it) * 8, c, (double)S);
            const BinFast b = make_bin_fast(idx, c, S, W);
            pair_k[(size_t)c * G + j] = b.ok ? (int16_t)b.l : (int16_t)-1;
            pair_a[(size_t)c * G + j] = (float)b.wg;
            if (b.ok) atomicAdd(&hs[c * NB + b.l], 1);
            else pair_v[(size_t)c * G + j] = make_float2(__int_as_float(0x7fc00000), 0.f);   // NaN: "not scored"
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nbuck; i += blockDim.x)
        blk_hist[(size_t)i * gridDim.x + blockIdx.x] = hs[i];    // bucket-major: the scan below is coalesced
}

// Every non-empty bucket is padded to whole groups of kBfNC pairs and to whole CTA slots of kBfWarps
// groups (a slot has one PRN and one lag).
__device__ __forceinline__ int groups_of_count(int cnt) {
    const int g = (cnt + kBfNC - 1) / kBfNC;
    return ((g + kBfWarps - 1) / kBfWarps) * kBfWarps;
}

// One CTA per bucket: exclusive scan of the bucket's per-CTA counts (in place), 256 CTAs per step, bucket
// total -> hist.  Last CTA: group base of every bucket + total group count.
__global__ void DPE_SIDE256
k_block_scan(int32_t* __restrict__ blk_hist, int nblk, int32_t* __restrict__ hist, int nbuck,
             int32_t* __restrict__ group_base, int64_t* __restrict__ bucket_base, int32_t* __restrict__ n_groups,
             int64_t max_groups, unsigned int* __restrict__ ticket) {
    __shared__ int32_t wsum[8];
    __shared__ int32_t s_carry;
    int32_t* row = blk_hist + (size_t)blockIdx.x * nblk;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nblk; b0 += 256) {
        const int b = b0 + threadIdx.x;
        const int32_t v = (b < nblk) ? row[b] : 0;
        int32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int32_t woff = 0;
        for (int w = 0; w < warp; ++w) woff += wsum[w];
        const int32_t carry = s_carry;
        if (b < nblk) row[b] = carry + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == 255) s_carry = carry + woff + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) hist[blockIdx.x] = s_carry;
    if (!take_last_ticket(ticket)) return;
    // ---- bucket -> group layout (all bucket totals are visible now) ----
    const int per = (nbuck + 255) / 256;                   // consecutive buckets per thread
    const int b0 = threadIdx.x * per, b1 = min(b0 + per, nbuck);
    int32_t mine = 0;
    for (int i = b0; i < b1; ++i) mine += groups_of_count(__ldcg(hist + i));
    int32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int32_t u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    __syncthreads();
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int32_t base = incl - mine;
    for (int w = 0; w < warp; ++w) base += wsum[w];
    for (int i = b0; i < b1; ++i) {
        group_base[i] = base;
        bucket_base[i] = (int64_t)base * kBfNC;
        base += groups_of_count(__ldcg(hist + i));
    }
    if (threadIdx.x == 255) {                              // the last thread ends on the total
        group_base[nbuck] = base;
        *n_groups = (base <= max_groups) ? base : 0;
    }
}

// Scatter the pairs into their bucket, in candidate order: bucket base + pairs of earlier CTAs
// (k_block_scan) + pairs of earlier warps of this CTA + rank among the warp's lanes.  The CTAs of channel
// row 0 also write the group headers {channel, lag, valid pairs}.
__global__ void DPE_SIDE128
k_scatter(const int16_t* __restrict__ pair_k, const float* __restrict__ pair_a, int64_t G, int W, int nbuck,
          const int32_t* __restrict__ hist, const int32_t* __restrict__ group_base,
          const int64_t* __restrict__ bucket_base, const int32_t* __restrict__ blk_base,
          int32_t* __restrict__ ent_j, float* __restrict__ ent_a, int4* __restrict__ hdr, int64_t max_groups) {
    extern __shared__ int32_t sm[];                        // [kSortBlock/32][NB] warp counts, then [nbuck+1] group bases
    constexpr int NW = kSortBlock / 32;
    const int NB = 2 * W + 1;
    int32_t* wcnt = sm;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < NW * NB; i += blockDim.x) wcnt[i] = 0;
    if (c == 0) {
        int32_t* gb = sm + NW * NB;
        for (int i = threadIdx.x; i <= nbuck; i += blockDim.x) gb[i] = group_base[i];
        __syncthreads();
        const int total = gb[nbuck];
        if (total <= max_groups) {
            for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x) {
                int lo = 0, hi = nbuck - 1;                // last bucket with base <= g
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (gb[mid] <= g) lo = mid; else hi = mid - 1;
                }
                const int q = g - gb[lo];
                const int cnt = hist[lo];
                const int ng = (cnt + kBfNC - 1) / kBfNC;
                int n = 0;
                if (q < ng) { n = cnt - q * kBfNC; if (n > kBfNC) n = kBfNC; }
                hdr[g] = make_int4(lo / NB, lo % NB, n, 0);
            }
        }
    }
    __syncthreads();
    const int k = (j < G) ? (int)pair_k[(size_t)c * G + j] : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, k);
    if (k >= 0 && lane == __ffs(peers) - 1) wcnt[warp * NB + k] = __popc(peers);
    __syncthreads();
    if (k < 0) return;
    int before = __popc(peers & ((1u << lane) - 1u));
    for (int w = 0; w < warp; ++w) before += wcnt[w * NB + k];
    const int key = c * NB + k;
    const int64_t pos = bucket_base[key] + blk_base[(size_t)key * gridDim.x + blockIdx.x] + before;
    ent_j[pos] = (int32_t)j;
    ent_a[pos] = pair_a[(size_t)c * G + j];
}

// ---------------------------------------------------------------------------
// mbarrier / TMA bulk-copy primitives (PTX; SASS: SYNCS.*, UBLKCP)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s2u(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s2u(b)), "r"(count) : "memory");
}
// barrier and bulk-copy operations on 32-bit shared-window addresses (computed once per kernel: no address
// arithmetic per tile)
__device__ __forceinline__ void mbar_expect_tx_u(uint32_t b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_u(uint32_t b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory");
}
__device__ __forceinline__ void mbar_wait_u(uint32_t b, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok) : "r"(b), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_1d_u(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int kXTileF = (int)skewX(kBfTile);               // 2560 floats: float4-skewed tile of 1024 elements

// Register image of one warp-chunk (256 positions): per lane 8 contiguous positions as 4 sample
// float4 (re,im,re,im) and 4 replica float4 (d,d,r,r).
struct BruteChunk {
    float4 x[4], rd[4];
    __device__ __forceinline__ void load(const float4* __restrict__ px, int ch) {
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = px[ch * 160 + i];
#pragma unroll
        for (int i = 0; i < 4; ++i) rd[i] = px[kXTileF / 4 + ch * 160 + i];
    }
    // FFMA2 operand economy (B200: an FFMA2 with three uncached 64-bit sources needs a third
    // register-file cycle): the blend takes alpha as a scalar .F32 operand and (d, r) pairs
    // shared by all candidates; the accumulate takes the blended chip as a scalar .F32 operand
    // and the (re,im) sample pair shared by all candidates.  Per 2 positions and candidate:
    // 1 FFMA2 blend + 2 FFMA2 accumulate = 12 FLOP.
    __device__ __forceinline__ void accumulate(const float (&al)[kBfNC], float2 (&acc)[kBfNC]) const {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 dp = make_float2(rd[q].x, rd[q].y), r0 = make_float2(rd[q].z, rd[q].w);
            const float2 xa = make_float2(x[q].x, x[q].y), xb = make_float2(x[q].z, x[q].w);
#pragma unroll
            for (int j = 0; j < kBfNC; ++j) {
                const float2 bp = __ffma2_rn(make_float2(al[j], al[j]), dp, r0);   // r + alpha d
                acc[j] = __ffma2_rn(make_float2(bp.x, bp.x), xa, acc[j]);
                acc[j] = __ffma2_rn(make_float2(bp.y, bp.y), xb, acc[j]);
            }
        }
    }
};

// ---------------------------------------------------------------------------
// k_brute: persistent; CTA slot = kBfWarps groups of one channel.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBfWarps * 32, 1)
k_brute(const float* __restrict__ bx, const float* __restrict__ brd,
        int64_t bx_stride, int64_t brd_stride, const int4* __restrict__ hdr,
        const int32_t* __restrict__ ent_j, const float* __restrict__ ent_a,
        const int32_t* __restrict__ n_groups, float2* __restrict__ pair_v, int64_t G, int S_pad,
        int H, int W, float2* __restrict__ tail_part, unsigned int* __restrict__ tail_ticket, int skip_pad) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bars[2 * kBfStages];  // full[s] = bars[s], empty[s] = bars[stages + s]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int stage_f = 2 * kXTileF;                   // floats per stage: sample tile, replica tile
    static_assert((kBfStages & (kBfStages - 1)) == 0 && (kBfWarps & (kBfWarps - 1)) == 0, "powers of two");
    const uint32_t full0 = s2u(&bars[0]), empty0 = s2u(&bars[kBfStages]), smem0 = s2u(smem);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kBfStages; ++s) { mbar_init(&bars[s], 1); mbar_init(&bars[kBfStages + s], kBfWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // Work decomposition: the CTA's share of the tile sequence of all slots (slot-major), cut at tile
    // granularity so that every CTA streams the same number of tiles +-1 whatever the slot count
    // (8 GPUs, demo: 10.3 waves of whole slots would cost 11).  A slot cut by a CTA boundary is summed in
    // parts: every contributing warp stores its partial sums and takes a ticket, the last one adds the
    // parts in CTA order (deterministic).
    const int n_slots = *n_groups / kBfWarps;
    const int ntiles = S_pad / kBfTile;
    const int64_t U = (int64_t)n_slots * ntiles;             // tiles of the whole launch
    const int n_cta = (int)(U < (int64_t)gridDim.x ? U : (int64_t)gridDim.x);   // CTAs that get work
    auto cta_begin = [&](int i) -> int64_t { return U * i / n_cta; };
    auto cta_of = [&](int64_t u) -> int {                    // the CTA whose share holds tile u
        int i = (int)(u * n_cta / U);
        while (cta_begin(i + 1) <= u) ++i;
        while (cta_begin(i) > u) --i;
        return i;
    };
    const bool working = (int)blockIdx.x < n_cta;
    const int64_t u0 = working ? cta_begin(blockIdx.x) : 0, u1 = working ? cta_begin(blockIdx.x + 1) : 0;
    const uint32_t my_tiles = (uint32_t)(u1 - u0);            // tiles this CTA streams, 0..my_tiles-1
    uint32_t it = 0;                                       // running tile counter (same on all warps)

    // TMA producer duty (one elected lane): stream tile `jt` of this CTA's share into stage `stg`.
    // There is no dedicated producer warp: 9 warps would put 3 on one scheduler and cap the kernel at
    // 168 registers per thread (16K registers per SM sub-partition); the duty rotates over the 8 warps.
    auto issue_tile = [&](uint32_t jt, uint32_t stg) {
        if (jt >= my_tiles) return;
        const int64_t u = u0 + jt;
        const int slot = (int)(u / ntiles), t = (int)(u - (int64_t)slot * ntiles);
        const int4 h = hdr[(size_t)slot * kBfWarps];         // one PRN and one lag per slot
        const int c = h.x, k = h.y - W;
        const int ks = k & 7, kq = (k - ks) / 8;             // k = 8 kq + ks, ks in 0..7
        constexpr uint32_t bytes = kXTileF * 4;
        const uint32_t dst = smem0 + stg * (stage_f * 4), bar = full0 + 8 * stg;
        mbar_expect_tx_u(bar, 2 * bytes);
        tma_load_1d_u(dst, bx + ((size_t)c * 8 + ks) * bx_stride + skewX((int64_t)t * kBfTile + 8 * kq + H), bytes, bar);
        tma_load_1d_u(dst + bytes, brd + c * brd_stride + (size_t)t * kXTileF, bytes, bar);
    };
    if (threadIdx.x == 0)
        for (uint32_t j = 0; j < kBfStages - kBfLag; ++j) issue_tile(j, j);   // prologue

    // ===== consumer warps =====
    const float4* const px0 = reinterpret_cast<const float4*>(smem) + 5 * lane;   // lane's run in stage 0 (skewX)
    const float4* px = px0;
    uint32_t stg = 0, par = 0;                             // stage and full-barrier parity of tile `it`
    for (int64_t u = u0; u < u1;) {
        const int slot = (int)(u / ntiles);
        const int t_begin = (int)(u - (int64_t)slot * ntiles);
        const int t_end = (u1 - u < (int64_t)(ntiles - t_begin)) ? t_begin + (int)(u1 - u) : ntiles;
        u += t_end - t_begin;
        const int g = slot * kBfWarps + warp;
        const int4 h = hdr[g];
        const int c = h.x, n_valid = h.z;
        // one coalesced load of the group's 32 alphas, then register broadcast (FFMA2 takes alpha as a
        // scalar .F32 operand)
        const float a_mine = (lane < n_valid) ? ent_a[(size_t)g * kBfNC + lane] : 0.f;
        float al[kBfNC];
#pragma unroll
        for (int j = 0; j < kBfNC; ++j) al[j] = __shfl_sync(0xffffffffu, a_mine, j);
        float2 acc[kBfNC];                                  // (re, im) of this lane's samples
#pragma unroll
        for (int j = 0; j < kBfNC; ++j) acc[j] = make_float2(0.f, 0.f);
        for (int t = t_begin; t < t_end; ++t) {
            mbar_wait_u(full0 + 8 * stg, par);
            // software pipeline over the 4 chunks of the tile: the shared-memory operands of chunk
            // ch+1 are in flight while chunk ch is computed (2 warps per scheduler are not enough to
            // hide the LDS latency otherwise: they run in lock step)
            if (n_valid > 0 || !skip_pad) {   // a padding group (bucket rounded up to a whole slot) only keeps the barriers going
                BruteChunk cur, nxt;
                cur.load(px, 0);
#pragma unroll
                for (int ch = 0; ch < kBfTile / kBfChunk; ++ch) {
                    if (ch + 1 < kBfTile / kBfChunk) nxt.load(px, ch + 1);
                    cur.accumulate(al, acc);
                    cur = nxt;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_u(empty0 + 8 * stg);
            // refill duty of tile `it`: reload the stage tile it-lag used (everyone released it about a
            // tile ago) with tile it+stages-lag
            if (warp == (int)(it & (kBfWarps - 1))) {
                const uint32_t rs = (it - kBfLag) & (kBfStages - 1);
                if (it >= kBfLag) mbar_wait_u(empty0 + 8 * rs, ((it - kBfLag) / kBfStages) & 1);
                if (lane == 0) issue_tile(it + kBfStages - kBfLag, rs);
                __syncwarp();
            }
            ++it;
            px += stage_f / 4;
            if (++stg == kBfStages) { stg = 0; par ^= 1; px = px0; }
        }

        // lane partials -> one candidate per lane: halving butterfly (31 shuffles per component instead of
        // 160): at step o the lanes with bit o set keep the upper half of the candidates, the others the
        // lower half; after 5 steps lane l holds the full sum of candidate l.  FP32 here (32 addends of
        // equal weight: ~1e-7); the scores are formed in FP64 from these FP32 sums (k_score_pairs).
#pragma unroll
        for (int o = 16, n = kBfNC / 2; o > 0; o >>= 1, n >>= 1) {
            const bool up = (lane & o) != 0;
#pragma unroll
            for (int j = 0; j < n; ++j) {
                const float2 keep = up ? acc[j + n] : acc[j];
                const float2 send = up ? acc[j] : acc[j + n];
                acc[j].x = keep.x + __shfl_xor_sync(0xffffffffu, send.x, o);
                acc[j].y = keep.y + __shfl_xor_sync(0xffffffffu, send.y, o);
            }
        }
        if (t_begin == 0 && t_end == ntiles) {             // the whole slot was mine
            if (lane < n_valid) {
                const int64_t j = ent_j[(size_t)g * kBfNC + lane];
                pair_v[(size_t)c * G + j] = acc[0];
            }
        } else {
            // part of a slot that a CTA boundary cuts.  Buffer 0 of a CTA holds the slot that began before
            // its share, buffer 1 the slot that continues after it.
            const int64_t s0 = (int64_t)slot * ntiles;
            const int i_first = cta_of(s0), i_last = cta_of(s0 + ntiles - 1);
            constexpr int kPart = kBfWarps * kBfNC;
            tail_part[((size_t)blockIdx.x * 2 + (t_begin != 0 ? 0 : 1)) * kPart + threadIdx.x] = acc[0];
            __threadfence();
            __syncwarp();
            unsigned int tk = 0;
            if (lane == 0) {
                unsigned int* ticket = &tail_ticket[i_first * kBfWarps + warp];
                tk = atomicAdd(ticket, 1u);
                if (tk == (unsigned int)(i_last - i_first)) *ticket = 0u;     // self-resetting for the next launch
            }
            tk = __shfl_sync(0xffffffffu, tk, 0);
            if (tk == (unsigned int)(i_last - i_first)) {    // last part in: add all of them in CTA order
                __threadfence();
                float2 v = make_float2(0.f, 0.f);
                for (int i = i_first; i <= i_last; ++i) {
                    const int buf = (cta_begin(i) > s0) ? 0 : 1;
                    const float2 w = __ldcg(&tail_part[((size_t)i * 2 + buf) * kPart + threadIdx.x]);
                    v.x += w.x; v.y += w.y;
                }
                if (lane < n_valid) {
                    const int64_t j = ent_j[(size_t)g * kBfNC + lane];
                    pair_v[(size_t)c * G + j] = v;
                }
            }
        }
    }
}

// pass 5: per-candidate score from the pair correlations + fused reductions.  A pair that fell outside
// the lag window carries NaN in pair_v (k_pair_bins).  WITH_SUMS = 0 (arg-max estimate): the candidate
// states are not needed per candidate -- 8 B x C + 8 B per candidate instead of 10 B x C + 40 B.
template <int WITH_SUMS>
__global__ void __launch_bounds__(kReduceBlock, 8)      // 64 registers: a streaming kernel, occupancy is what hides its loads
k_score_pairs(const double* __restrict__ grid, const EpochDev* __restrict__ ep,
              const float2* __restrict__ pair_v, int lpower,
              int64_t G, int64_t grid_offset, double* __restrict__ scores, double* __restrict__ blk_partial,
              unsigned int* __restrict__ ticket, double* __restrict__ partial, const FoldEst fold) {
    const EpochDev& e = *ep;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = j < G;
    double score = 0.0;
    int oow = 0;
    Cand p = {0, 0, 0, 0};
    if (active) {
        if (WITH_SUMS) p = cand_ecef(e, grid + 4 * j);
        // batches of 4 channels: the four streamed loads are in flight together, the sums stay in channel order
        for (int c0 = 0; c0 < e.C; c0 += 4) {
            float2 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
                v[q] = (c0 + q < e.C) ? __ldcs(&pair_v[(size_t)(c0 + q) * G + j]) : make_float2(0.f, 0.f);   // streamed once
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (c0 + q >= e.C) break;
                if (v[q].x == v[q].x) score += mag_pow((double)v[q].x, (double)v[q].y, lpower);
                else ++oow;
            }
        }
        __stcs(&scores[j], score);
    }
    block_reduce_store(score, j + grid_offset, p, active, oow, blk_partial);
    if (take_last_ticket(ticket)) finish_position_partial<2>(blk_partial, gridDim.x, grid, e, grid_offset, partial, fold);
}

size_t brute_smem_bytes(int) {
    return (size_t)kBfStages * 2 * kXTileF * sizeof(float);
}

// second and third launch of the pair sort (per-bucket scan of the per-CTA counts + bucket layout; headers + scatter),
// shared by the position pairs (buckets = PRN x lag) and the velocity pairs (PRN x Doppler bin, dpe_vel.cu)
int launch_sort_tail(dpe_ctx* c, const SortLists& L, int64_t G, int W, int C, unsigned int* ticket, cudaStream_t s) {
    const int NB = 2 * W + 1, nbuck = C * NB;
    const int nblk = (int)((G + kSortBlock - 1) / kSortBlock);
    k_block_scan<<<nbuck, 256, 0, s>>>(L.blk_hist, nblk, L.hist, nbuck, L.group_base, L.bucket_base, L.n_groups,
                                       L.max_groups, ticket);
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    dim3 gs(nblk, C);
    k_scatter<<<gs, kSortBlock, sizeof(int32_t) * ((kSortBlock / 32) * NB + nbuck + 1), s>>>(
        L.pair_k, L.pair_a, G, W, nbuck, L.hist, L.group_base, L.bucket_base, L.blk_hist,
        reinterpret_cast<int32_t*>(L.ent_j), L.ent_a, reinterpret_cast<int4*>(L.hdr), L.max_groups);
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    return DPE_OK;
}

// bins + stable counting sort of the pairs into (PRN, lag) buckets / groups / slots; needs the epoch
// parameters only (dpe_brute_presort may run it on another stream than the sample pre-pass)
int launch_brute_sort(dpe_ctx* c, int sat_mode, cudaStream_t s) {
    const int C = c->epoch_C, NB = 2 * c->W + 1, nbuck = C * NB;
    prof_begin(c, DPE_STAGE_BRUTE_BINS, s);
    const int nblk = (int)((c->G + kSortBlock - 1) / kSortBlock);
    const size_t hs_bytes = sizeof(int32_t) * nbuck;
    if (sat_mode == DPE_SAT_PER_TIME) {
        int rc = launch_sat_geo(c, s);
        if (rc) return rc;
        k_pair_bins<DPE_SAT_PER_TIME><<<nblk, kSortBlock, hs_bytes, s>>>(c->grid, c->ep, c->sat, c->cfg.fs, (int)c->S,
            c->W, c->T, c->G, c->cfg.grid_offset, c->pair_k, c->pair_a, c->pair_v, c->blk_hist,
            reinterpret_cast<const SatGeo*>(c->sat_geo));
    } else
        k_pair_bins<DPE_SAT_MIDDLE><<<nblk, kSortBlock, hs_bytes, s>>>(c->grid, c->ep, c->sat, c->cfg.fs, (int)c->S,
            c->W, c->T, c->G, c->cfg.grid_offset, c->pair_k, c->pair_a, c->pair_v, c->blk_hist, nullptr);
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    SortLists L = {c->pair_k, c->pair_a, c->blk_hist, c->hist, c->group_base, c->bucket_base, c->n_groups, c->hdr, c->ent_j,
                   c->ent_a, c->max_groups};
    int rc = launch_sort_tail(c, L, c->G, c->W, C, c->ticket + 1, s);
    if (rc) return rc;
    prof_end(c, s);
    return DPE_OK;
}

int brute_set_attributes(dpe_ctx* c) {     // per context: the attributes belong to the device the context lives on
    DPE_CUDA(cudaFuncSetAttribute(k_brute, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    // the whole unified L1 as shared memory: one carve-out for k_brute and the side kernels that share its SMs
    DPE_CUDA(cudaFuncSetAttribute(k_brute, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared));
    c->brute_attr_set = 1;
    return DPE_OK;
}

// k_brute alone (the pair sort and the planes are in place)
int launch_brute_corr(dpe_ctx* c, cudaStream_t s) {
    const size_t smem = brute_smem_bytes(c->H);
    if (!c->brute_attr_set) { int rc = brute_set_attributes(c); if (rc) return rc; }
    prof_begin(c, DPE_STAGE_BRUTE_CORR, s);
    // With a communicator, k_brute leaves `comm_reserve_sms` SMs alone: an NCCL kernel needs a whole SM, and with
    // every SM under a persistent k_brute CTA the broadcast of the NEXT epoch (and the all-gather of the previous one)
    // would wait for this launch to end -- the pre-pass they gate could then not run under it.
    const int n_cta = c->sm_count - (c->comm ? c->comm_reserve_sms : 0);
    k_brute<<<n_cta > 0 ? n_cta : 1, kBfWarps * 32, smem, s>>>(
        c->bx, c->brd, c->bx_stride, c->brd_stride, reinterpret_cast<const int4*>(c->hdr),
        reinterpret_cast<const int32_t*>(c->ent_j), c->ent_a, c->n_groups, c->pair_v, c->G, (int)c->S_pad,
        c->H, c->W, c->tail_part, c->tail_ticket, c->brute_skip_pad);
    prof_end(c, s);
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    return DPE_OK;
}

int launch_brute_score(dpe_ctx* c, cudaStream_t s) {
    prof_begin(c, DPE_STAGE_BRUTE_SCORE, s);
    const FoldEst fold = {c->fold_est_mode, c->zval, c->rval, c->result};
    const int nblk = (int)((c->G + kReduceBlock - 1) / kReduceBlock);
    if (c->want_sums)
        k_score_pairs<1><<<nblk, kReduceBlock, 0, s>>>(c->grid, c->ep, c->pair_v, c->cfg.lpower, c->G,
                                                       c->cfg.grid_offset, c->scores, c->blk_partial, c->ticket, c->partial, fold);
    else
        k_score_pairs<0><<<nblk, kReduceBlock, 0, s>>>(c->grid, c->ep, c->pair_v, c->cfg.lpower, c->G,
                                                       c->cfg.grid_offset, c->scores, c->blk_partial, c->ticket, c->partial, fold);
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    c->n_blk_partial = nblk;
    prof_end(c, s);
    return DPE_OK;
}

int launch_score_brute(dpe_ctx* c, int sat_mode, cudaStream_t s) {
    int rc;
    if (!c->have_planes) {
        if ((rc = launch_brute_planes(c, s))) return rc;
        c->have_planes = 1;
    }
    if (c->sort_pending) { DPE_CUDA(cudaStreamWaitEvent(s, c->ev_sort, 0)); c->sort_pending = 0; }
    if (c->sort_valid != 1 + sat_mode) {                   // not presorted (maybe on another stream) this epoch
        if ((rc = launch_brute_sort(c, sat_mode, s))) return rc;
        c->sort_valid = 1 + sat_mode;
    }
    if ((rc = launch_brute_corr(c, s))) return rc;
    return launch_brute_score(c, s);
}

int kernel_attr_brute(const char* name, cudaFuncAttributes* a) {
    DPE_KATTR("k_brute", k_brute);
    DPE_KATTR("k_pair_bins", k_pair_bins<DPE_SAT_MIDDLE>);
    DPE_KATTR("k_block_scan", k_block_scan);
    DPE_KATTR("k_scatter", k_scatter);
    DPE_KATTR("k_score_pairs", k_score_pairs<0>);
    DPE_KATTR("k_score_pairs_sums", k_score_pairs<1>);
    return 0;
}

}  // namespace dpe
