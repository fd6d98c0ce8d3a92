// dpe_prepare.cu -- per-PRN pre-pass of the DPE hot path (sm_100a):
//   C/A code table, int16 I/Q unpack + carrier wipe-off + replica sign, and the
//   windowed circular code correlogram with the flip / no-flip choice.
//
// Replaces (reference paths relative to cudarecv/modules/src):
//   BCS_GenCACode            batchcorrscores.cu:117-177
//   BCS_GenTimeIdcs          batchcorrscores.cu:185-196
//   BCS_Load                 batchcorrscores.cu:209-221
//   BCS_NavBitBoundary       batchcorrscores.cu:237-258
//   BCS_ComputeDopplerWipeoff batchcorrscores.cu:277-305
//   BCS_ComputeCodeReplica   batchcorrscores.cu:323-372
//   cuFFT chain + BCS_ChooseCodeCorr + BCS_cufftBatchShift  :1099-1153
// The reference computes all S lags with five batched FFTs; a position grid only
// ever looks at lags within +-W samples of the prompt, so the correlogram is
// produced directly in the lag domain on that window (identical values:
// c[k] = sum_n xw[(n+k) mod S] r[n]).
#include "dpe_internal.cuh"

namespace dpe {

// ---------------------------------------------------------------------------
// C/A Gold codes.  G1 = 1 + x^3 + x^10, G2 = 1 + x^2 + x^3 + x^6 + x^8 + x^9 + x^10,
// G2 output = XOR of two phase-select stages (IS-GPS-200 table 3-I).  chip = +1
// where the XOR of the two sequences is 1 (same polarity as BCS_GenCACode's
// -g1*g2 with registers initialised to -1).
// ---------------------------------------------------------------------------
__constant__ uint8_t kPhaseSel[37][2] = {
    {2, 6}, {3, 7}, {4, 8}, {5, 9}, {1, 9}, {2, 10}, {1, 8}, {2, 9}, {3, 10}, {2, 3},
    {3, 4}, {5, 6}, {6, 7}, {7, 8}, {8, 9}, {9, 10}, {1, 4}, {2, 5}, {3, 6}, {4, 7},
    {5, 8}, {6, 9}, {1, 3}, {4, 6}, {5, 7}, {6, 8}, {7, 9}, {8, 10}, {1, 6}, {2, 7},
    {3, 8}, {4, 9}, {5, 10}, {4, 10}, {1, 7}, {2, 8}, {4, 10}};

__global__ void k_gen_ca(int8_t* __restrict__ ca) {
    int p = threadIdx.x;
    if (p >= DPE_MAX_CHAN) return;
    // bit s-1 of the word holds stage s (stage 1 = input side, stage 10 = output side)
    uint32_t g1 = 0x3FF, g2 = 0x3FF;
    const int s1 = kPhaseSel[p][0] - 1, s2 = kPhaseSel[p][1] - 1;
    for (int i = 0; i < K_L_CA; ++i) {
        uint32_t o1 = (g1 >> 9) & 1u;
        uint32_t o2 = ((g2 >> s1) ^ (g2 >> s2)) & 1u;
        ca[p * 1024 + i] = (o1 ^ o2) ? 1 : -1;
        uint32_t f1 = ((g1 >> 2) ^ (g1 >> 9)) & 1u;
        uint32_t f2 = ((g2 >> 1) ^ (g2 >> 2) ^ (g2 >> 5) ^ (g2 >> 7) ^ (g2 >> 8) ^ (g2 >> 9)) & 1u;
        g1 = ((g1 << 1) | f1) & 0x3FF;
        g2 = ((g2 << 1) | f2) & 0x3FF;
    }
    ca[p * 1024 + 1023] = 0;
}

// BCS_NavBitBoundary (batchcorrscores.cu:237-258), same expression order.
__device__ __forceinline__ int nav_edge_index(const EpochDev& e, int c, double fs) {
    int since = (((e.cp_start[c] - e.cp_ref[c]) % 20) + 20) % 20;
    int to_next = 20 - since;
    return (int)(floor((K_L_CA * to_next - e.rc_start[c]) * (fs / e.fc[c])) + 1);
}

// ---------------------------------------------------------------------------
// BCS_GenTimeIdcs (batchcorrscores.cu:185-196): t_n = round(n / fs * 1e9) / 1e9, once per context.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gen_time(double fs, int S, double* __restrict__ tidx) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= S) return;
    double t = (double)n / fs;                                // :190-193
    tidx[n] = round(t * 1.0e9) / 1.0e9;
}

// ---------------------------------------------------------------------------
// k_prep_corr: the whole per-PRN pre-pass of a lookup epoch in ONE launch -- int16 unpack, carrier wipe-off,
// C/A replica, windowed circular correlogram, flip / no-flip choice, fft-shifted CodeScores window.
// (Round 1: k_prepare + k_corr_partial + k_corr_finalize, the wiped samples making a round trip through HBM.)
//
// One CTA per (chunk of 1024 samples, channel), 128 threads (a side kernel: shares an SM with k_brute).
//   1. the CTA wipes its 1024 samples plus the circular halo of NLp + 8 the lag window needs, straight into
//      the word-skewed shared tile the correlation reads.  Time index, carrier phase and chip index are FP64
//      (t * fc ~ 2e4 chips, fi * t ~ 1e2 cycles do not survive FP32; the chip index is a bit-exact target);
//      the phase is reduced to [0, 1) cycles in FP64 and only then handed to the FP32 sincospif -- the
//      range-reduced FP32 NCO the north star asks for (phase error <= 6e-8 cycle per sample, incoherent).
//   2. each thread owns an 8-sample x 8-lag register tile (64 complex MACs per 23 shared loads), FP32 inside
//      the chunk, exact 64-bit fixed point across chunks (integer atomics: order-independent); part A = samples
//      before the nav-bit edge, part B = from the edge on, so no-flip = A + B and flipped = A - B without a second pass.
//   3. the CTA that takes the last ticket of its channel reads the channel's totals, decides flip / no-flip on
//      lag 0 (BCS_ChooseCodeCorr, batchcorrscores.cu:499-543) and writes the window (BCS_cufftBatchShift,
//      :554-584: cs[l] <-> shifted bin l - W + S/2).
// xw / rs (and the conjugate carrier cc for the carrier branch) still go to HBM once: the brute-force planes and the
// velocity manifold read them.
// ---------------------------------------------------------------------------
__global__ void DPE_SIDE128
k_prep_corr(const int16_t* __restrict__ iq, const int8_t* __restrict__ ca, const double* __restrict__ tidx,
            const EpochDev* __restrict__ ep, double fs, int S, int W, int NL, int NLp, int nchunk,
            float2* __restrict__ xw, int8_t* __restrict__ rs, int16_t* __restrict__ chip_idx,
            int32_t* __restrict__ idx_next, long long* __restrict__ dc_part, float2* __restrict__ cc,
            long long* __restrict__ cacc, double2* __restrict__ cs, int32_t* __restrict__ no_flip,
            unsigned int* __restrict__ chan_ticket) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_edge;
    __shared__ bool s_last;
    __shared__ int s_dc[4][2];
    const int c = blockIdx.y;
    const EpochDev& e = *ep;
    if (c >= e.C) return;
    DPE_PT_DECL;
    DPE_PT_MARK();                                    // [0] start
    grid_dep_trigger();                               // a dependent scoring kernel may set itself up beside this one
    const int chunk = blockIdx.x;
    const int n0 = chunk * kCorrChunk;
    const int nx = kCorrChunk + NLp + 8;             // halo: lags -W .. -W+NLp-1 (+7 slack)
    float2* xs = reinterpret_cast<float2*>(smem_raw);                 // [skew(nx)]
    float* r_s = reinterpret_cast<float*>(xs + (nx + (nx >> 3) + 1)); // [skew(1024)]
    int8_t* code_s = reinterpret_cast<int8_t*>(r_s + (kCorrChunk + (kCorrChunk >> 3) + 1));   // [1024]

    const int prn = e.prn[c];
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        reinterpret_cast<int32_t*>(code_s)[i] = reinterpret_cast<const int32_t*>(ca + (prn - 1) * 1024)[i];
    if (threadIdx.x == 0) {
        const int en = nav_edge_index(e, c, fs);
        s_edge = en;
        if (chunk == 0) idx_next[c] = en;
    }
    __syncthreads();
    DPE_PT_MARK();                                    // [1] code table + edge in shared memory

    const double fc = e.fc[c], rc = e.rc_start[c], fi = e.fi[c], ri = e.ri_start[c];
    // carrier branch (velocity manifold): the CTAs of channel 0 also sum the raw samples of their chunk -- the DC sum of
    // the block (thrust::reduce, batchcorrscores.cu:1065) as exact integers, chunk by chunk -- and every channel keeps the
    // conjugate carrier, so that (x - mean) conj(carrier) = xw - mean cc is formed where the mean is known (dpe_vel.cu)
    int dc_i = 0, dc_q = 0;
    // batches of kPrepBatch samples per thread: the global loads of a batch are in flight together (one sample at a time
    // the kernel sat in "long scoreboard": 3.3 stalled warps per issue at 16 % occupancy); 2 x 5 x 128 covers the 1024
    // samples of the chunk and the halo of the usual lag windows in two rounds of loads
    constexpr int kPrepBatch = 5;
    for (int i0 = threadIdx.x; i0 < nx; i0 += kPrepBatch * (int)blockDim.x) {
        double tq[kPrepBatch];
        short2 vq[kPrepBatch];
        int nq[kPrepBatch];
#pragma unroll
        for (int q = 0; q < kPrepBatch; ++q) {
            const int i = i0 + q * blockDim.x;
            int n = n0 - W + i;
            n %= S; if (n < 0) n += S;
            nq[q] = n;
            tq[q] = (i < nx) ? __ldg(tidx + n) : 0.0;
            vq[q] = (i < nx) ? __ldg(reinterpret_cast<const short2*>(iq) + n) : make_short2(0, 0);
        }
#pragma unroll
        for (int q = 0; q < kPrepBatch; ++q) {
            const int i = i0 + q * blockDim.x;
            if (i >= nx) break;
            const int n = nq[q];
            const double t = tq[q];
            const short2 v = vq[q];
            const double ph = fi * t + ri;                        // cycles (:293)
            const float fr = (float)(ph - floor(ph));             // [0, 1): range reduction in FP64, NCO in FP32
            float sn, cn;
            sincospif(2.0f * fr, &sn, &cn);
            const float I = (float)v.x, Q = (float)v.y;
            const float2 x = make_float2(fmaf(I, cn, Q * sn), fmaf(Q, cn, -I * sn));   // x * conj(exp(j phi)) (cuCmul, :385-407)
            xs[i + (i >> 3)] = x;
            const int m = i - W;                                  // sample of the chunk's own 1024
            if (m >= 0 && m < kCorrChunk) {
                float r = 0.f;
                if (n0 + m < S) {                                 // n == n0 + m here (no wrap inside the block)
                    int chip = (int)floor(t * fc + rc);           // :347-348
                    chip = ((chip % K_L_CA) + K_L_CA) % K_L_CA;
                    const int8_t rv = code_s[chip];
                    r = (float)rv;
                    const size_t o = (size_t)c * S + n;
                    rs[o] = rv;
                    if (chip_idx) chip_idx[o] = (int16_t)chip;
                    xw[o] = x;
                    if (cc) cc[o] = make_float2(cn, -sn);
                    dc_i += v.x; dc_q += v.y;
                }
                r_s[m + (m >> 3)] = r;
            }
        }
    }
    __syncthreads();
    DPE_PT_MARK();                                    // [2] samples wiped into the shared tile

    int edge = s_edge;
    if (!(edge > 0 && edge < S)) edge = S;           // no edge in block: everything is part A
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (dc_part && c == 0) {                          // <= 9 samples per thread: the sums fit an int with room to spare
        dc_i = __reduce_add_sync(0xffffffffu, dc_i);
        dc_q = __reduce_add_sync(0xffffffffu, dc_q);
        if (lane == 0) { s_dc[warp][0] = dc_i; s_dc[warp][1] = dc_q; }
        __syncthreads();
        if (threadIdx.x == 0) {
            long long si = 0, sq = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { si += s_dc[w][0]; sq += s_dc[w][1]; }
            dc_part[2 * chunk] = si; dc_part[2 * chunk + 1] = sq;
        }
    }
    // One tile = (run of 8 lags, range of 8-sample runs): 64 complex MACs per lane and step in registers, then the 32 sums
    // of the tile -- 8 lags x (A, B) x (re, im) -- go through a halving butterfly (31 shuffles; lane L ends up with the
    // whole-warp sum of value L) and every lane adds its value to the channel's FIXED-POINT accumulator with one 64-bit
    // integer atomic: integer addition is associative, so the total does not depend on the order the CTAs arrive in
    // (bit-identical from run to run), and no per-chunk partials travel through HBM to a serial reduction at the end.
    // (Before: cpart[C][nchunk][2][NLp] in FP64 and a last CTA per channel that walked the 49 chunks of its 34 lags --
    // 10 of the kernel's 28 us, profiles/r02ae_phase_stamps.txt.)  Scale 2^19: an FP32 partial of magnitude >= 16 is
    // represented exactly, |partial| < 2^27 and nchunk <= 2^16 keep the total below 2^62.
    // Lag runs are dealt to the warps whole while a full round of them is left; the remaining ones (5 runs on 4 warps at
    // W = 16) are cut into one quarter of the samples per warp, so that no warp works while the others wait.
    const int n_lag_runs = NLp / kLagTile;
    const int nw = (int)(blockDim.x >> 5);
    long long* const acc_c = cacc + (size_t)c * NLp * 4;
    auto run_tile = [&](const int lr, const int sr_begin, const int sr_end) {
        float2 accA[kLagTile], accB[kLagTile];
#pragma unroll
        for (int l = 0; l < kLagTile; ++l) { accA[l] = make_float2(0.f, 0.f); accB[l] = accA[l]; }
#pragma unroll 1
        for (int sr = sr_begin + lane; sr < sr_end; sr += 32) {
            const int m = sr * 8;                     // local sample of the run
            float r[8];
#pragma unroll
            for (int s_ = 0; s_ < 8; ++s_) r[s_] = r_s[m + s_ + sr];        // skew: (m+s) + (m+s)/8
            float2 x[15];
            const int xb = m + lr * kLagTile;         // xs index of (sample m, lag lr*8)
#pragma unroll
            for (int s_ = 0; s_ < 15; ++s_) { int i = xb + s_; x[s_] = xs[i + (i >> 3)]; }
            const int nabs = n0 + m;
            const bool allA = (nabs + 8 <= edge), allB = (nabs >= edge);
            if (allA || allB) {
                float2* acc = allA ? accA : accB;
#pragma unroll
                for (int l = 0; l < kLagTile; ++l)
#pragma unroll
                    for (int s_ = 0; s_ < 8; ++s_) {
                        acc[l].x = fmaf(x[s_ + l].x, r[s_], acc[l].x);
                        acc[l].y = fmaf(x[s_ + l].y, r[s_], acc[l].y);
                    }
            } else {
#pragma unroll
                for (int s_ = 0; s_ < 8; ++s_) {
                    const bool a = (nabs + s_) < edge;
#pragma unroll
                    for (int l = 0; l < kLagTile; ++l) {
                        if (a) { accA[l].x = fmaf(x[s_ + l].x, r[s_], accA[l].x); accA[l].y = fmaf(x[s_ + l].y, r[s_], accA[l].y); }
                        else   { accB[l].x = fmaf(x[s_ + l].x, r[s_], accB[l].x); accB[l].y = fmaf(x[s_ + l].y, r[s_], accB[l].y); }
                    }
                }
            }
        }
        float v[4 * kLagTile];                        // value 4 l + {0, 1, 2, 3} = A.re, A.im, B.re, B.im of lag lr*8 + l
#pragma unroll
        for (int l = 0; l < kLagTile; ++l) { v[4 * l] = accA[l].x; v[4 * l + 1] = accA[l].y; v[4 * l + 2] = accB[l].x; v[4 * l + 3] = accB[l].y; }
        static_assert(4 * kLagTile == 32, "one value per lane after the butterfly");
#pragma unroll
        for (int o = 16, n = 16; o > 0; o >>= 1, n >>= 1) {
            const bool up = (lane & o) != 0;          // lanes with bit o set keep the upper half of the values
#pragma unroll
            for (int j = 0; j < n; ++j) {
                const float keep = up ? v[j + n] : v[j];
                const float send = up ? v[j] : v[j + n];
                v[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
        }
        const long long q = __double2ll_rn((double)v[0] * kFixScale);
        if (q != 0) atomicAdd(reinterpret_cast<unsigned long long*>(acc_c + lr * 32 + lane), (unsigned long long)q);
    };
    const int n_whole = (n_lag_runs / nw) * nw;
    for (int lr = warp; lr < n_whole; lr += nw) run_tile(lr, 0, kCorrChunk / 8);
    constexpr int kSteps = kCorrChunk / 8 / 32;       // 32-lane steps per lag run
    static_assert(kSteps >= 1 && kCorrChunk % 256 == 0, "whole warps of 8-sample runs");
    for (int it = warp; it < (n_lag_runs - n_whole) * kSteps; it += nw)
        run_tile(n_whole + it / kSteps, (it % kSteps) * 32, (it % kSteps) * 32 + 32);

    // ---- last CTA of this channel: chunk partials -> CodeScores window ----
    DPE_PT_MARK();                                    // [3] thread 0's (warp 0's) lag runs done
    __threadfence();
    __syncthreads();
    DPE_PT_MARK();                                    // [4] every warp done
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(&chan_ticket[c], 1u);
        s_last = (t == (unsigned int)nchunk - 1);
        if (s_last) chan_ticket[c] = 0;               // self-resetting for the next launch
    }
    __syncthreads();
    DPE_PT_MARK();                                    // [5] ticket taken
    DPE_PT_PRINT("prep", !s_last && (blockIdx.x % 12) == 0 && (blockIdx.y % 7) == 0);
    if (!s_last) return;
    __threadfence();
    const bool has_edge = (s_edge > 0) && (s_edge < S);
    // the channel's accumulators are complete: back to FP64 (exact: a power-of-two scale), and cleared for the next launch
    double2* sumA = reinterpret_cast<double2*>(smem_raw);      // the shared tiles are dead by now: [NL] + [NL]
    double2* sumB = sumA + NL;
    for (int l = threadIdx.x; l < NLp; l += blockDim.x) {
        longlong2* p = reinterpret_cast<longlong2*>(acc_c + 4 * l);
        if (l < NL) {
            const longlong2 a = __ldcg(p), b2 = __ldcg(p + 1);
            sumA[l] = make_double2((double)a.x * (1.0 / kFixScale), (double)a.y * (1.0 / kFixScale));
            sumB[l] = make_double2((double)b2.x * (1.0 / kFixScale), (double)b2.y * (1.0 / kFixScale));
        }
        p[0] = make_longlong2(0, 0);
        p[1] = make_longlong2(0, 0);
    }
    __syncthreads();
    // lag 0 decides (:512); every thread evaluates the same expression on the same sums
    const double2 a0 = sumA[W], b0 = sumB[W];
    const bool keep = !has_edge || (hypot(a0.x + b0.x, a0.y + b0.y) > hypot(a0.x - b0.x, a0.y - b0.y));
    if (threadIdx.x == 0) no_flip[c] = keep;
    for (int l = threadIdx.x; l < NL; l += blockDim.x) {
        const double2 a = sumA[l], b2 = sumB[l];
        cs[(size_t)c * NL + l] = keep ? make_double2(a.x + b2.x, a.y + b2.y)     // no-flip = A + B
                                      : make_double2(a.x - b2.x, a.y - b2.y);    // flipped = A - B (only chosen when an edge exists)
    }
    DPE_PT_MARK();                                    // [6] window written (last CTA of the channel)
    DPE_PT_PRINT("prep-last", true);
}

// ---------------------------------------------------------------------------
// Planes of the brute-force kernel (dpe_brute.cu).  The kernel walks the replica
// position p; a pair with lag k multiplies replica position p with sample
// (p + k) mod S.  k_sample_planes writes the wiped samples 8 times, copy s shifted
// by s samples, each with a circular halo of H elements, so that lag k = 8 q + s is
// copy s read at element offset 8 q: 16-byte aligned for the TMA bulk copy and in
// phase with the float4 skew for every lag.  One thread per element pair.
// ---------------------------------------------------------------------------
__global__ void DPE_SIDE256
k_sample_planes(const float2* __restrict__ xw, const EpochDev* __restrict__ ep, int S, int n_elem, int H,
                float* __restrict__ bx, int64_t bx_stride) {
    const int c = blockIdx.y, s = blockIdx.z;
    if (c >= ep->C) return;
    const int e = 2 * (blockIdx.x * blockDim.x + threadIdx.x);   // element = p + H
    if (e >= n_elem) return;
    int n0 = (e - H + s) % S; if (n0 < 0) n0 += S;
    int n1 = n0 + 1; if (n1 == S) n1 = 0;
    const float2 a = xw[(size_t)c * S + n0], b = xw[(size_t)c * S + n1];
    *reinterpret_cast<float4*>(bx + ((size_t)c * 8 + s) * bx_stride + skewX(e)) = make_float4(a.x, a.y, b.x, b.y);
}

// k_replica_rd: chosen replica (flip applied) by position pair for the brute-force kernel:
// (d[p], d[p+1], r[p], r[p+1]) with d[p] = r[(p-1) mod S] - r[p], so that the blended chip of a
// candidate is r + alpha d; zero beyond S (the padded tail of the last tile contributes nothing).
__global__ void DPE_SIDE256
k_replica_rd(const int8_t* __restrict__ rs, const int32_t* __restrict__ idx_next,
             const int32_t* __restrict__ no_flip, const EpochDev* __restrict__ ep,
             int S, int S_pad, float* __restrict__ brd, int64_t brd_stride) {
    const int c = blockIdx.y;
    if (c >= ep->C) return;
    const int p = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (p >= S_pad) return;
    const bool flip = !no_flip[c];
    const int edge = idx_next[c];
    auto rep = [&](int n) -> float {                      // n in [0, S)
        const float r = (float)rs[(size_t)c * S + n];
        return (flip && n >= edge) ? -r : r;
    };
    float r[2] = {0.f, 0.f}, d[2] = {0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int n = p + q;
        if (n < S) {
            r[q] = rep(n);
            d[q] = rep(n == 0 ? S - 1 : n - 1) - r[q];
        }
    }
    *reinterpret_cast<float4*>(brd + c * brd_stride + skewX(p)) = make_float4(d[0], d[1], r[0], r[1]);
}

// ---------------------------------------------------------------------------
int launch_gen_ca(dpe_ctx* c, cudaStream_t s) {
    k_gen_ca<<<1, 64, 0, s>>>(c->ca);
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    return DPE_OK;
}

int launch_gen_time(dpe_ctx* c, cudaStream_t s) {
    k_gen_time<<<(int)((c->S + 255) / 256), 256, 0, s>>>(c->cfg.fs, (int)c->S, c->tidx);
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    return DPE_OK;
}

// pre-pass + correlogram: one launch (k_prep_corr)
int launch_prepare(dpe_ctx* c, cudaStream_t s) {
    const int S = (int)c->S;
    const int nx = kCorrChunk + c->NLp + 8;
    size_t smem = (size_t)(nx + (nx >> 3) + 1) * sizeof(float2) +
                  (size_t)(kCorrChunk + (kCorrChunk >> 3) + 1) * sizeof(float) + 1024;
    if (smem < 2 * sizeof(double2) * (size_t)c->NL) smem = 2 * sizeof(double2) * (size_t)c->NL;     // the last CTA's sums reuse the tiles
    dim3 grid(c->nchunk, c->epoch_C);
    prof_begin(c, DPE_STAGE_PREPARE, s);
    k_prep_corr<<<grid, 128, smem, s>>>(c->iq, c->ca, c->tidx, c->ep, c->cfg.fs, S, c->W, c->NL, c->NLp, c->nchunk,
                                        c->xw, c->rs, c->chip_idx, c->idx_next, c->Gv > 0 ? c->dc_part : nullptr,
                                        c->Gv > 0 ? c->bb : nullptr, c->cacc, c->cs, c->no_flip, c->chan_ticket);
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    prof_end(c, s);
    return DPE_OK;
}

// the correlogram is finished by the last CTA of every channel inside k_prep_corr: nothing left to launch
int launch_correlogram(dpe_ctx*, cudaStream_t) { return DPE_OK; }

// The planes k_brute streams (built on the first brute-force scoring of an epoch, so that a
// context which only looks up pays nothing for them).
int launch_brute_planes(dpe_ctx* c, cudaStream_t s) {
    const int S = (int)c->S, S_pad = (int)c->S_pad;
    const int n_elem = S_pad + 2 * c->H;
    dim3 g1((n_elem / 2 + 255) / 256, c->epoch_C, 8);
    k_sample_planes<<<g1, 256, 0, s>>>(c->xw, c->ep, S, n_elem, c->H, c->bx, c->bx_stride);
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    dim3 g2((S_pad / 2 + 255) / 256, c->epoch_C);
    k_replica_rd<<<g2, 256, 0, s>>>(c->rs, c->idx_next, c->no_flip, c->ep, S, S_pad, c->brd, c->brd_stride);
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    return DPE_OK;
}

int kernel_attr_prepare(const char* name, cudaFuncAttributes* a) {
    DPE_KATTR("k_gen_ca", k_gen_ca);
    DPE_KATTR("k_prep_corr", k_prep_corr);
    DPE_KATTR("k_gen_time", k_gen_time);
    DPE_KATTR("k_sample_planes", k_sample_planes);
    DPE_KATTR("k_replica_rd", k_replica_rd);
    return 0;
}

}  // namespace dpe
