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
// k_prepare: one thread per (channel, sample).  16 B vector load of 4 I/Q pairs
// per thread, FP64 time index / carrier phase / chip index (bit-exactness of
// the chip index and of the wiped samples does not survive FP32: t*fc ~ 2e4
// chips, fi*t ~ 1e2 cycles), FP32 results.
// ---------------------------------------------------------------------------
__global__ void DPE_SIDE128
k_prepare(const int16_t* __restrict__ iq, const int8_t* __restrict__ ca,
          const EpochDev* __restrict__ ep, double fs, int S, int S_pad,
          float2* __restrict__ xw, int8_t* __restrict__ rs, int16_t* __restrict__ chip_idx,
          int32_t* __restrict__ idx_next, const long long* __restrict__ dc, float2* __restrict__ zw) {
    __shared__ int8_t code_s[1024];
    const int c = blockIdx.y;
    const EpochDev& e = *ep;
    if (c >= e.C) return;
    const int prn = e.prn[c];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) code_s[i] = ca[(prn - 1) * 1024 + i];
    if (blockIdx.x == 0 && threadIdx.x == 0) idx_next[c] = nav_edge_index(e, c, fs);
    __syncthreads();

    const double fc = e.fc[c], rc = e.rc_start[c], fi = e.fi[c], ri = e.ri_start[c];
    const int n0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (n0 >= S_pad) return;
    int16_t v[8];
    if (n0 + 4 <= S) {
        *reinterpret_cast<int4*>(v) = *reinterpret_cast<const int4*>(iq + 2 * (size_t)n0);
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            bool in = (n0 + q) < S;
            v[2 * q] = in ? iq[2 * (size_t)(n0 + q)] : (int16_t)0;
            v[2 * q + 1] = in ? iq[2 * (size_t)(n0 + q) + 1] : (int16_t)0;
        }
    }
    // DC mean for the carrier branch: sum / (float)S (ComplexDivide, batchcorrscores.cu:1065,1210-1216)
    const double inv = 1.0 / (double)(float)S;
    const double mr = dc ? (double)dc[0] * inv : 0.0, mi = dc ? (double)dc[1] * inv : 0.0;
    float xr[4], xi[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int n = n0 + q;
        xr[q] = 0.f; xi[q] = 0.f;
        if (n < S) {
            double t = (double)n / fs;                        // BCS_GenTimeIdcs :190-193
            t = round(t * 1.0e9) / 1.0e9;
            double sn, cs;
            sincos(2 * K_PI * (fi * t + ri), &sn, &cs);       // :293
            const double I = (double)v[2 * q], Q = (double)v[2 * q + 1];
            // x * conj(exp(j phi)) (cuCmul, :385-407)
            xr[q] = (float)(I * cs + Q * sn);
            xi[q] = (float)(Q * cs - I * sn);
            int chip = (int)floor(t * fc + rc);               // :347-348
            chip = ((chip % K_L_CA) + K_L_CA) % K_L_CA;
            rs[(size_t)c * S + n] = code_s[chip];
            if (chip_idx) chip_idx[(size_t)c * S + n] = (int16_t)chip;
            xw[(size_t)c * S + n] = make_float2(xr[q], xi[q]);
            if (zw)   // (x - mean) * conj(carrier), BCS_SubtractDCOffset :470-485
                zw[(size_t)c * S + n] = make_float2((float)((I - mr) * cs + (Q - mi) * sn),
                                                    (float)((Q - mi) * cs - (I - mr) * sn));
        }
    }
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

// ---------------------------------------------------------------------------
// k_corr_partial: one CTA per (chunk of 1024 samples, channel).  The chunk of xw
// (with a +-W halo, circular) and of r is staged in shared memory in a
// word-skewed layout (x + x/8) so that lanes holding runs of 8 contiguous samples
// read conflict-free; each thread owns an 8-sample x 8-lag register tile
// (64 complex MACs per 23 shared loads), FP32 inside the chunk, FP64 across chunks.
// Part A = samples before the nav-bit edge, part B = from the edge on, so that
// no-flip = A + B and flipped = A - B without a second pass.
// ---------------------------------------------------------------------------
__global__ void DPE_SIDE128
k_corr_partial(const float2* __restrict__ xw, const int8_t* __restrict__ rs,
               const int32_t* __restrict__ idx_next, const EpochDev* __restrict__ ep,
               int S, int W, int NLp, int nchunk, double2* __restrict__ cpart) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int c = blockIdx.y;
    if (c >= ep->C) return;
    const int chunk = blockIdx.x;
    const int n0 = chunk * kCorrChunk;
    const int nx = kCorrChunk + NLp + 8;             // halo: lags -W .. -W+NLp-1 (+7 slack)
    float2* xs = reinterpret_cast<float2*>(smem_raw);                 // [skew(nx)]
    float* r_s = reinterpret_cast<float*>(xs + (nx + (nx >> 3) + 1)); // [skew(1024)]

    const float2* xc = xw + (size_t)c * S;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
        int n = n0 - W + i;
        n %= S; if (n < 0) n += S;
        xs[i + (i >> 3)] = xc[n];
    }
    for (int i = threadIdx.x; i < kCorrChunk; i += blockDim.x) {
        int n = n0 + i;
        r_s[i + (i >> 3)] = (n < S) ? (float)rs[(size_t)c * S + n] : 0.f;
    }
    __syncthreads();

    int edge = idx_next[c];
    if (!(edge > 0 && edge < S)) edge = S;           // no edge in block: everything is part A
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_lag_runs = NLp / kLagTile;
    for (int lr = warp; lr < n_lag_runs; lr += (int)(blockDim.x >> 5)) {
        float2 accA[kLagTile], accB[kLagTile];
#pragma unroll
        for (int l = 0; l < kLagTile; ++l) { accA[l] = make_float2(0.f, 0.f); accB[l] = accA[l]; }
#pragma unroll 1
        for (int sr = lane; sr < kCorrChunk / 8; sr += 32) {
            const int m = sr * 8;                     // local sample of the run
            float r[8];
#pragma unroll
            for (int s = 0; s < 8; ++s) r[s] = r_s[m + s + sr];        // skew: (m+s) + (m+s)/8
            float2 x[15];
            const int xb = m + lr * kLagTile;         // xs index of (sample m, lag lr*8)
#pragma unroll
            for (int s = 0; s < 15; ++s) { int i = xb + s; x[s] = xs[i + (i >> 3)]; }
            const int nabs = n0 + m;
            const bool allA = (nabs + 8 <= edge), allB = (nabs >= edge);
            if (allA || allB) {
                float2* acc = allA ? accA : accB;
#pragma unroll
                for (int l = 0; l < kLagTile; ++l)
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        acc[l].x = fmaf(x[s + l].x, r[s], acc[l].x);
                        acc[l].y = fmaf(x[s + l].y, r[s], acc[l].y);
                    }
            } else {
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const bool a = (nabs + s) < edge;
#pragma unroll
                    for (int l = 0; l < kLagTile; ++l) {
                        if (a) { accA[l].x = fmaf(x[s + l].x, r[s], accA[l].x); accA[l].y = fmaf(x[s + l].y, r[s], accA[l].y); }
                        else   { accB[l].x = fmaf(x[s + l].x, r[s], accB[l].x); accB[l].y = fmaf(x[s + l].y, r[s], accB[l].y); }
                    }
                }
            }
        }
#pragma unroll
        for (int l = 0; l < kLagTile; ++l) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                accA[l].x += __shfl_xor_sync(0xffffffffu, accA[l].x, o);
                accA[l].y += __shfl_xor_sync(0xffffffffu, accA[l].y, o);
                accB[l].x += __shfl_xor_sync(0xffffffffu, accB[l].x, o);
                accB[l].y += __shfl_xor_sync(0xffffffffu, accB[l].y, o);
            }
        }
        if (lane == 0) {
            double2* out = cpart + (((size_t)c * nchunk + chunk) * 2) * NLp + lr * kLagTile;
#pragma unroll
            for (int l = 0; l < kLagTile; ++l) {
                out[l] = make_double2((double)accA[l].x, (double)accA[l].y);
                out[NLp + l] = make_double2((double)accB[l].x, (double)accB[l].y);
            }
        }
    }
}

// k_corr_finalize: fixed-order FP64 sum of the chunk partials, flip / no-flip
// choice on lag 0 (BCS_ChooseCodeCorr, batchcorrscores.cu:499-543), fft-shifted
// window out (BCS_cufftBatchShift, :554-584: cs[l] <-> shifted bin l - W + S/2).
// One warp per lag (lanes stride the chunks, xor-tree: a fixed summation order); every warp also
// sums lag 0 so the decision needs no cross-CTA exchange.
__device__ __forceinline__ void sum_chunks(const double2* __restrict__ cpart, int c, int l, int NLp, int nchunk,
                                           int lane, double (&r)[4]) {
    r[0] = r[1] = r[2] = r[3] = 0.0;
    for (int ch = lane; ch < nchunk; ch += 32) {
        const double2* p = cpart + (((size_t)c * nchunk + ch) * 2) * NLp + l;
        r[0] += p[0].x; r[1] += p[0].y; r[2] += p[NLp].x; r[3] += p[NLp].y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] += __shfl_xor_sync(0xffffffffu, r[k], o);
}

__global__ void DPE_SIDE256
k_corr_finalize(const double2* __restrict__ cpart, const int32_t* __restrict__ idx_next,
                const EpochDev* __restrict__ ep, int S, int W, int NL, int NLp,
                int nchunk, double2* __restrict__ cs, int32_t* __restrict__ no_flip) {
    const int c = blockIdx.x;
    if (c >= ep->C) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int l = blockIdx.y * 8 + warp;
    const int edge_raw = idx_next[c];
    const bool edge = (edge_raw > 0) && (edge_raw < S);
    double z[4];
    sum_chunks(cpart, c, W, NLp, nchunk, lane, z);                // lag 0
    const bool keep = !edge || (hypot(z[0] + z[2], z[1] + z[3]) > hypot(z[0] - z[2], z[1] - z[3]));
    if (blockIdx.y == 0 && threadIdx.x == 0) no_flip[c] = keep ? 1 : 0;
    if (l >= NL) return;
    double r[4];
    sum_chunks(cpart, c, l, NLp, nchunk, lane, r);
    if (lane == 0) {
        double2 v;
        if (keep) v = make_double2(r[0] + r[2], r[1] + r[3]);     // no-flip = A + B
        else v = make_double2(r[0] - r[2], r[1] - r[3]);          // flipped = A - B (only chosen when an edge exists)
        cs[(size_t)c * NL + l] = v;
    }
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

int launch_prepare(dpe_ctx* c, cudaStream_t s) {
    const int S = (int)c->S;
    const int S4 = ((S + 3) / 4) * 4;
    dim3 grid((S4 / 4 + 127) / 128, c->epoch_C);
    prof_begin(c, DPE_STAGE_PREPARE, s);
    if (c->Gv > 0) { int rc = launch_dc_sum(c, s); if (rc) return rc; }
    k_prepare<<<grid, 128, 0, s>>>(c->iq, c->ca, c->ep, c->cfg.fs, S, S4, c->xw, c->rs, c->chip_idx, c->idx_next,
                                   c->Gv > 0 ? c->dc_sum : nullptr, c->Gv > 0 ? c->bb : nullptr);
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    prof_end(c, s);
    return DPE_OK;
}

int launch_correlogram(dpe_ctx* c, cudaStream_t s) {
    const int S = (int)c->S;
    const int nx = kCorrChunk + c->NLp + 8;
    const size_t smem = (size_t)(nx + (nx >> 3) + 1) * sizeof(float2) +
                        (size_t)(kCorrChunk + (kCorrChunk >> 3) + 1) * sizeof(float);
    dim3 grid(c->nchunk, c->epoch_C);
    prof_begin(c, DPE_STAGE_CORRELOGRAM, s);
    k_corr_partial<<<grid, 128, smem, s>>>(c->xw, c->rs, c->idx_next, c->ep, S, c->W, c->NLp,
                                           c->nchunk, c->cpart);
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    dim3 gf(c->epoch_C, (c->NL + 7) / 8);
    k_corr_finalize<<<gf, 256, 0, s>>>(c->cpart, c->idx_next, c->ep, S, c->W, c->NL, c->NLp, c->nchunk, c->cs,
                                       c->no_flip);
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    prof_end(c, s);
    return DPE_OK;
}

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
    DPE_KATTR("k_prepare", k_prepare);
    DPE_KATTR("k_sample_planes", k_sample_planes);
    DPE_KATTR("k_corr_partial", k_corr_partial);
    DPE_KATTR("k_corr_finalize", k_corr_finalize);
    DPE_KATTR("k_replica_rd", k_replica_rd);
    return 0;
}

}  // namespace dpe
