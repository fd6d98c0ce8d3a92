// dpe_brute.cu -- the north-star kernel: every (candidate, PRN) pair correlates the
// whole 20 ms block against its own blended C/A replica (SURVEY.md section 8 a').
//
//   v(j,c) = sum_m xw_c[m] * ( (1-a) r_c[(m-k) mod S] + a r_c[(m-k-1) mod S] )
//
// with integer lag k and fraction a from the candidate's FP64 geometry
// (batchcorrmanifold.cu:1779-1800).  By linearity this equals the reference's
// lerp of two correlogram bins (:1806-1812) to rounding.
//
// Mapping (B200, FP32-pipe bound; the block and the replica live in SMEM / L2):
//   * pairs are bucketed by (PRN, k) so that the 16 pairs of a warp ("group")
//     share the replica window; lanes own runs of 8 contiguous samples, the 16
//     candidates live in registers: per sample-pair 1 FFMA (blend) + 1 FFMA2
//     (complex accumulate, packed over two consecutive samples);
//   * sample / replica tiles of 1024 samples are staged by 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier, 4 stages) issued by a dedicated producer warp;
//     the planes are stored pre-skewed in HBM so the staged tiles are read with
//     conflict-free LDS.128 (samples) and lane-stride-9 LDS.32 (replica);
//   * persistent CTAs (one per SM), 8 consumer warps + 1 producer warp;
//   * lane partials are combined with warp shuffles in FP64.
#include "dpe_geom.cuh"

namespace dpe {

// ---------------------------------------------------------------------------
// pass 1: bins of every pair + (PRN, lag) histogram
// ---------------------------------------------------------------------------
template <int SAT_MODE>
__global__ void __launch_bounds__(256)
k_pair_bins(const double* __restrict__ grid, const EpochDev* __restrict__ ep, const double* __restrict__ sat,
            double fs, int S, int W, int T, int64_t G, int64_t grid_offset, int16_t* __restrict__ pair_k,
            float* __restrict__ pair_a, int32_t* __restrict__ hist) {
    extern __shared__ int32_t hs[];
    const EpochDev& e = *ep;
    const int NB = 2 * W + 1;
    const int nbuck = e.C * NB;
    for (int i = threadIdx.x; i < nbuck; i += blockDim.x) hs[i] = 0;
    __syncthreads();
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < G) {
        const Cand p = cand_ecef(e, grid + 4 * j);
        const int it = (SAT_MODE == DPE_SAT_PER_TIME) ? (int)((j + grid_offset) % T) : T / 2;
        for (int c = 0; c < e.C; ++c) {
            const double idx = code_index(e, p, sat + ((size_t)c * T + it) * 8, c, fs, (double)S);
            const Bin b = make_bin(idx, c, S, W);
            pair_k[(size_t)c * G + j] = b.ok ? (int16_t)b.l : (int16_t)-1;
            pair_a[(size_t)c * G + j] = (float)b.wg;
            if (b.ok) atomicAdd(&hs[c * NB + b.l], 1);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nbuck; i += blockDim.x)
        if (hs[i]) atomicAdd(&hist[i], hs[i]);
}

// ---------------------------------------------------------------------------
// pass 2: bucket -> group layout.  Every bucket is padded to whole groups of
// kBfNC pairs, every channel to whole CTAs of kBfWarps groups.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_bucket_scan(const int32_t* __restrict__ hist, const EpochDev* __restrict__ ep, int W,
              int64_t* __restrict__ bucket_base, int4* __restrict__ hdr, int32_t* __restrict__ n_groups,
              int64_t max_groups) {
    extern __shared__ int32_t sm[];        // [nbuck] counts, then [nbuck+1] group bases
    const int NB = 2 * W + 1;
    const int C = ep->C;
    const int nbuck = C * NB;
    int32_t* cnt = sm;
    int32_t* gb = sm + nbuck;
    for (int i = threadIdx.x; i < nbuck; i += blockDim.x) cnt[i] = hist[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        int32_t g = 0;
        for (int c = 0; c < C; ++c) {
            for (int b = 0; b < NB; ++b) {
                gb[c * NB + b] = g;
                g += (cnt[c * NB + b] + kBfNC - 1) / kBfNC;
            }
            g = ((g + kBfWarps - 1) / kBfWarps) * kBfWarps;
        }
        gb[nbuck] = g;
        *n_groups = (g <= max_groups) ? g : 0;
    }
    __syncthreads();
    const int total = gb[nbuck];
    if (total > max_groups) return;
    for (int i = threadIdx.x; i < nbuck; i += blockDim.x) bucket_base[i] = (int64_t)gb[i] * kBfNC;
    for (int g = threadIdx.x; g < total; g += blockDim.x) {
        int lo = 0, hi = nbuck - 1;            // last bucket with gb <= g
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (gb[mid] <= g) lo = mid; else hi = mid - 1;
        }
        const int q = g - gb[lo];
        const int ng = (cnt[lo] + kBfNC - 1) / kBfNC;
        int n = 0;
        if (q < ng) { n = cnt[lo] - q * kBfNC; if (n > kBfNC) n = kBfNC; }
        hdr[g] = make_int4(lo / NB, lo % NB, n, 0);
    }
}

// pass 3: scatter the pairs into their bucket
__global__ void __launch_bounds__(256)
k_scatter(const int16_t* __restrict__ pair_k, const float* __restrict__ pair_a, int64_t G, int W,
          const int64_t* __restrict__ bucket_base, int32_t* __restrict__ cursor,
          int32_t* __restrict__ ent_j, float* __restrict__ ent_a) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (j >= G) return;
    const int k = pair_k[(size_t)c * G + j];
    if (k < 0) return;
    const int i = c * (2 * W + 1) + k;
    const int64_t pos = bucket_base[i] + atomicAdd(&cursor[i], 1);
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
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s2u(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s2u(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok) : "r"(s2u(b)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(s2u(dst)), "l"(src), "r"(bytes), "r"(s2u(bar)) : "memory");
}

constexpr int kXTileF = kBfTile + kBfTile / 8;            // 1152 floats: float4-skewed 1024-sample plane tile

// ---------------------------------------------------------------------------
// k_brute: persistent; CTA slot = kBfWarps groups of one channel.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__((kBfWarps + 1) * 32, 1)
k_brute(const float* __restrict__ bxr, const float* __restrict__ bxi, const float* __restrict__ brr,
        int64_t bx_stride, int64_t br_stride, const int4* __restrict__ hdr,
        const int32_t* __restrict__ ent_j, const float* __restrict__ ent_a,
        const int32_t* __restrict__ n_groups, double2* __restrict__ pair_v, int64_t G, int S_pad,
        int H, int W) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bar[kBfStages], empty_bar[kBfStages];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rr_len = (int)skewR(kBfTile + 2 * H);        // floats, multiple of 4
    const int stage_f = 2 * kXTileF + rr_len;              // floats per stage
    float* const stage0 = reinterpret_cast<float*>(smem);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kBfStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kBfWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int n_slots = *n_groups / kBfWarps;
    const int ntiles = S_pad / kBfTile;
    uint32_t it = 0;                                       // running tile counter (same on all warps)

    if (warp == kBfWarps) {
        // ===== producer warp: one elected lane streams the tiles =====
        if (lane == 0) {
            const uint32_t bytes_x = kXTileF * 4, bytes_r = (uint32_t)rr_len * 4;
            for (int slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
                const int c = hdr[(size_t)slot * kBfWarps].x;
                const float* sxr = bxr + c * bx_stride;
                const float* sxi = bxi + c * bx_stride;
                const float* srr = brr + c * br_stride;
                for (int t = 0; t < ntiles; ++t, ++it) {
                    const int s = it % kBfStages;
                    mbar_wait(&empty_bar[s], ((it / kBfStages) & 1) ^ 1);
                    float* dst = stage0 + (size_t)s * stage_f;
                    mbar_expect_tx(&full_bar[s], 2 * bytes_x + bytes_r);
                    tma_load_1d(dst, sxr + (size_t)t * kXTileF, bytes_x, &full_bar[s]);
                    tma_load_1d(dst + kXTileF, sxi + (size_t)t * kXTileF, bytes_x, &full_bar[s]);
                    tma_load_1d(dst + 2 * kXTileF, srr + (size_t)t * kXTileF, bytes_r, &full_bar[s]);
                }
            }
        }
        return;
    }

    // ===== consumer warps =====
    const int lane_f4 = 2 * lane + (lane >> 2);            // float4 index of this lane's run (skewX)
    for (int slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
        const int g = slot * kBfWarps + warp;
        const int4 h = hdr[g];
        const int c = h.x, k = h.y - W, n_valid = h.z;
        float2 al[kBfNC];                                   // (alpha, alpha): FFMA2 has no scalar broadcast
#pragma unroll
        for (int j = 0; j < kBfNC; ++j) {
            const float a = (j < n_valid) ? ent_a[(size_t)g * kBfNC + j] : 0.f;
            al[j] = make_float2(a, a);
        }
        float2 are[kBfNC], aim[kBfNC];
#pragma unroll
        for (int j = 0; j < kBfNC; ++j) { are[j] = make_float2(0.f, 0.f); aim[j] = make_float2(0.f, 0.f); }
        // replica window: lane run starts at local x' = Lu + chunk*256 + lane*8, Lu = H - k - 1
        const int Lu = H - k - 1;
        int off[kBfNS + 1];
#pragma unroll
        for (int i = 0; i <= kBfNS; ++i) off[i] = (int)skewR(Lu + i) + 9 * lane;

        for (int t = 0; t < ntiles; ++t, ++it) {
            const int s = it % kBfStages;
            mbar_wait(&full_bar[s], (it / kBfStages) & 1);
            const float* st = stage0 + (size_t)s * stage_f;
            const float4* pxr = reinterpret_cast<const float4*>(st) + lane_f4;
            const float4* pxi = reinterpret_cast<const float4*>(st + kXTileF) + lane_f4;
            const float* prr = st + 2 * kXTileF;
#pragma unroll 1
            for (int ch = 0; ch < kBfTile / kBfChunk; ++ch) {
                const float4 a0 = pxr[ch * 72], a1 = pxr[ch * 72 + 1];
                const float4 b0 = pxi[ch * 72], b1 = pxi[ch * 72 + 1];
                float rr[kBfNS + 1];
#pragma unroll
                for (int i = 0; i <= kBfNS; ++i) rr[i] = prr[off[i] + ch * 288];
                float2 dp[4], r0[4];                        // (r1 - r0) and r0 of samples 2q, 2q+1
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    dp[q] = make_float2(rr[2 * q] - rr[2 * q + 1], rr[2 * q + 1] - rr[2 * q + 2]);
                    r0[q] = make_float2(rr[2 * q + 1], rr[2 * q + 2]);
                }
                const float2 xre[4] = {make_float2(a0.x, a0.y), make_float2(a0.z, a0.w),
                                       make_float2(a1.x, a1.y), make_float2(a1.z, a1.w)};
                const float2 xim[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w),
                                       make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
                for (int j = 0; j < kBfNC; ++j) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        // blended replica of samples 2q, 2q+1: r0 + alpha (r1 - r0)
                        const float2 bp = __ffma2_rn(al[j], dp[q], r0[q]);
                        are[j] = __ffma2_rn(xre[q], bp, are[j]);
                        aim[j] = __ffma2_rn(xim[q], bp, aim[j]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
        }

        // lane partials -> FP64 -> warp butterfly; lane j keeps candidate j
        double outr = 0.0, outi = 0.0;
#pragma unroll
        for (int j = 0; j < kBfNC; ++j) {
            double re = (double)are[j].x + (double)are[j].y;
            double im = (double)aim[j].x + (double)aim[j].y;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                re += __shfl_xor_sync(0xffffffffu, re, o);
                im += __shfl_xor_sync(0xffffffffu, im, o);
            }
            if (lane == j) { outr = re; outi = im; }
        }
        if (lane < n_valid) {
            const int64_t j = ent_j[(size_t)g * kBfNC + lane];
            pair_v[(size_t)c * G + j] = make_double2(outr, outi);
        }
    }
}

// pass 5: per-candidate score from the pair correlations + fused reductions
__global__ void __launch_bounds__(kReduceBlock)
k_score_pairs(const double* __restrict__ grid, const EpochDev* __restrict__ ep,
              const int16_t* __restrict__ pair_k, const double2* __restrict__ pair_v, int lpower,
              int64_t G, int64_t grid_offset, double* __restrict__ scores, double* __restrict__ blk_partial) {
    const EpochDev& e = *ep;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = j < G;
    double score = 0.0;
    int oow = 0;
    Cand p = {0, 0, 0, 0};
    if (active) {
        p = cand_ecef(e, grid + 4 * j);
        for (int c = 0; c < e.C; ++c) {
            if (pair_k[(size_t)c * G + j] >= 0) {
                const double2 v = pair_v[(size_t)c * G + j];
                score += mag_pow(v.x, v.y, lpower);
            } else {
                ++oow;
            }
        }
        scores[j] = score;
    }
    block_reduce_store(score, j + grid_offset, p, active, oow, blk_partial);
}

size_t brute_smem_bytes(int H) {
    return (size_t)kBfStages * (2 * kXTileF + skewR(kBfTile + 2 * H)) * sizeof(float);
}

int launch_brute_passes(dpe_ctx* c, int sat_mode, cudaStream_t s) {
    const int C = c->epoch_C, NB = 2 * c->W + 1, nbuck = C * NB;
    prof_begin(c, DPE_STAGE_BRUTE_BINS, s);
    DPE_CUDA(cudaMemsetAsync(c->hist, 0, sizeof(int32_t) * nbuck, s));
    DPE_CUDA(cudaMemsetAsync(c->cursor, 0, sizeof(int32_t) * nbuck, s));
    const int nblk = (int)((c->G + 255) / 256);
    const size_t hs_bytes = sizeof(int32_t) * nbuck;
    if (sat_mode == DPE_SAT_PER_TIME)
        k_pair_bins<DPE_SAT_PER_TIME><<<nblk, 256, hs_bytes, s>>>(c->grid, c->ep, c->sat, c->cfg.fs, (int)c->S,
            c->W, c->T, c->G, c->cfg.grid_offset, c->pair_k, c->pair_a, c->hist);
    else
        k_pair_bins<DPE_SAT_MIDDLE><<<nblk, 256, hs_bytes, s>>>(c->grid, c->ep, c->sat, c->cfg.fs, (int)c->S,
            c->W, c->T, c->G, c->cfg.grid_offset, c->pair_k, c->pair_a, c->hist);
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    k_bucket_scan<<<1, 1024, sizeof(int32_t) * (2 * nbuck + 1), s>>>(c->hist, c->ep, c->W, c->bucket_base,
                                                                   reinterpret_cast<int4*>(c->hdr), c->n_groups,
                                                                   c->max_groups);
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    dim3 gs(nblk, C);
    k_scatter<<<gs, 256, 0, s>>>(c->pair_k, c->pair_a, c->G, c->W, c->bucket_base, c->cursor,
                                 reinterpret_cast<int32_t*>(c->ent_j), c->ent_a);
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    prof_end(c, s);
    return DPE_OK;
}

int launch_score_brute(dpe_ctx* c, int sat_mode, cudaStream_t s) {
    int rc = launch_brute_passes(c, sat_mode, s);
    if (rc) return rc;
    const size_t smem = brute_smem_bytes(c->H);
    static bool attr_set = false;
    if (!attr_set) {
        DPE_CUDA(cudaFuncSetAttribute(k_brute, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    prof_begin(c, DPE_STAGE_BRUTE_CORR, s);
    k_brute<<<c->sm_count, (kBfWarps + 1) * 32, smem, s>>>(
        c->bxr, c->bxi, c->brr, c->bx_stride, c->br_stride, reinterpret_cast<const int4*>(c->hdr),
        reinterpret_cast<const int32_t*>(c->ent_j), c->ent_a, c->n_groups, c->pair_v, c->G, (int)c->S_pad,
        c->H, c->W);
    prof_end(c, s);
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    prof_begin(c, DPE_STAGE_BRUTE_SCORE, s);
    const int nblk = (int)((c->G + kReduceBlock - 1) / kReduceBlock);
    k_score_pairs<<<nblk, kReduceBlock, 0, s>>>(c->grid, c->ep, c->pair_k, c->pair_v, c->cfg.lpower, c->G,
                                                c->cfg.grid_offset, c->scores, c->blk_partial);
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    c->n_blk_partial = nblk;
    rc = launch_reduce_partials(c, s);
    prof_end(c, s);
    return rc;
}

}  // namespace dpe
