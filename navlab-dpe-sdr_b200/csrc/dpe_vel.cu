// dpe_vel.cu -- velocity / clock-drift manifold.
//
// Reference (cudarecv/modules/src): the DC-removed, wiped, replica-stripped block is zero-padded to
// N_c = 8 * 2^ceil(log2 S) points and transformed by one batched Z2Z FFT -> "CarrScores"
// (batchcorrscores.cu:1158-1180, BCS_SubtractDCOffset :470-485, BCS_ChoosyBatchMultiplyAndPad
// :422-452); BCM_VelMeasML then looks every velocity candidate's Doppler bin up with a lerp
// (batchcorrmanifold.cu:1861-1963), thrust::max_element, BCM_MakeVelMeas (:2030-2068).
//
// Here: a velocity grid of +-V m/s only reaches Doppler bins within +-Wd of the prompt
// (bin width fs / N_c = 4.77 Hz at 2.5 MHz), so those 2*Wd+2 bins of the zero-padded spectrum are
// evaluated directly,  carr[m] = sum_n bb[n] exp(-j 2 pi n m / N_c),  with the twiddle angle
// reduced exactly in integers before sincospif; FP32 inside a 1024-sample chunk, FP64 across chunks.
// Those bins are a narrow band (+-310 Hz of 2.5 MHz): inside a block of 32 samples the twiddle turns by at most
// 0.012 rad, so a block enters every bin through its four complex MOMENTS  M_k = sum_b t^k bb[n0 + b]
// (k_carr_partial; Taylor remainder < 1e-9, below the FP32 rounding of the sums) -- 22 FP32 operations per
// (block, bin) instead of 256.  Windows too wide for that bound take the direct kernel (k_carr_partial_direct).
// The scoring kernel is the position one with a Doppler geometry.
#include "dpe_geom.cuh"

namespace dpe {

// DC mean of the block, sum / (float)S (ComplexDivide, batchcorrscores.cu:1065,1210-1216), from the per-chunk integer sums
// k_prep_corr left behind; every warp evaluates it for itself (lanes stride the chunks, xor-tree: no barrier needed).
__device__ __forceinline__ float2 dc_mean(const long long* __restrict__ dc_part, int nchunk, int S) {
    long long si = 0, sq = 0;
    for (int ch = threadIdx.x & 31; ch < nchunk; ch += 32) { si += dc_part[2 * ch]; sq += dc_part[2 * ch + 1]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        si += __shfl_xor_sync(0xffffffffu, si, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    const double inv = 1.0 / (double)(float)S;
    return make_float2((float)((double)si * inv), (float)((double)sq * inv));
}

// bb[n] = (x[n] - mean) conj(carrier[n]) * chosen replica[n] = (xw[n] - mean cc[n]) r[n]
// (BCS_SubtractDCOffset :470-485, BCS_ChoosyBatchMultiplyAndPad :422-452)
__device__ __forceinline__ float2 baseband(const float2* __restrict__ xw, const float2* __restrict__ cc,
                                           const int8_t* __restrict__ rs, size_t o, float2 mean, bool negate) {
    const float2 x = xw[o], k = cc[o];
    float r = (float)rs[o];
    if (negate) r = -r;
    return make_float2((x.x - (mean.x * k.x - mean.y * k.y)) * r, (x.y - (mean.x * k.y + mean.y * k.x)) * r);
}

// Chunk partials of the carrier spectrum meet in a per-channel FIXED-POINT accumulator (64-bit integer atomics, scale 2^19:
// integer addition is associative, so the spectrum is bit-identical whatever order the CTAs arrive in -- the same scheme as
// the correlogram's, dpe_prepare.cu); the CTA that takes the channel's last ticket converts the totals to CarrScores and
// clears the accumulator for the next launch.  (Before: vpart[C][nchunk][NBd] in FP64 and a k_carr_finalize launch, 6 us.)
__device__ __forceinline__ void carr_add(long long* __restrict__ acc_c, int l, double re, double im) {
    const long long qr = __double2ll_rn(re * kFixScale), qi = __double2ll_rn(im * kFixScale);
    if (qr != 0) atomicAdd(reinterpret_cast<unsigned long long*>(acc_c + 2 * l), (unsigned long long)qr);
    if (qi != 0) atomicAdd(reinterpret_cast<unsigned long long*>(acc_c + 2 * l + 1), (unsigned long long)qi);
}
// all threads of the CTA, after their carr_add calls
__device__ __forceinline__ void carr_tail(long long* __restrict__ acc_c, unsigned int* __restrict__ ticket, int nchunk, int NBd,
                                          double2* __restrict__ carr_c) {
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        s_last = (t == (unsigned int)nchunk - 1);
        if (s_last) *ticket = 0;                       // self-resetting for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int l = threadIdx.x; l < NBd; l += blockDim.x) {
        longlong2* p = reinterpret_cast<longlong2*>(acc_c + 2 * l);
        const longlong2 a = __ldcg(p);
        carr_c[l] = make_double2((double)a.x * (1.0 / kFixScale), (double)a.y * (1.0 / kFixScale));
        *p = make_longlong2(0, 0);
    }
}

// One CTA per (1024-sample chunk, channel); warp w owns bins w, w+8, ...; lanes stride the samples.
// bb[n] = zw[n] * chosen replica[n]  with zw = (x - mean) * conj(carrier) from k_prepare
// (BCS_SubtractDCOffset :470-485, BCS_ChoosyBatchMultiplyAndPad :422-452).
// Twiddles: lane handles samples n0 + lane + 32 i; exp(-j 2 pi n m / N_c) is evaluated exactly
// (integer n*m mod N_c -> sincospif) every 8th step and advanced by the exact 32-sample rotation in
// between (7 complex multiplies: < 5e-7 relative drift).
__global__ void DPE_SIDE256
k_carr_partial_direct(const float2* __restrict__ xw, const float2* __restrict__ cc, const long long* __restrict__ dc_part,
                      const int8_t* __restrict__ rs, const int32_t* __restrict__ idx_next,
                      const int32_t* __restrict__ no_flip, const EpochDev* __restrict__ ep, int S, int Wd, int NBd,
                      int n_fft, int nchunk, int n_dc, long long* __restrict__ vacc, unsigned int* __restrict__ vticket,
                      double2* __restrict__ carr) {
    __shared__ float2 xs[kCarrChunk];
    const int c = blockIdx.y;
    if (c >= ep->C) return;
    grid_dep_trigger();
    const int chunk = blockIdx.x, n0 = chunk * kCarrChunk;
    const bool flip = !no_flip[c];
    const int edge = idx_next[c];
    const float2 mean = dc_mean(dc_part, n_dc, S);
    for (int i = threadIdx.x; i < kCarrChunk; i += blockDim.x) {
        const int n = n0 + i;
        xs[i] = (n < S) ? baseband(xw, cc, rs, (size_t)c * S + n, mean, flip && n >= edge) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned mask = (unsigned)n_fft - 1u;                 // n_fft is a power of two
    const float scale = 2.0f / (float)n_fft;
    for (int l = warp; l < NBd; l += 8) {
        const int m = l - Wd;                                   // bin relative to 0 Hz
        float stp_s, stp_c;                                     // rotation of 32 samples: exp(-j 2 pi 32 m / N_c)
        sincospif((float)((32u * (unsigned)m) & mask) * scale, &stp_s, &stp_c);
        float ar = 0.f, ai = 0.f;
#pragma unroll 1
        for (int i0 = lane; i0 < kCarrChunk; i0 += 32 * 8) {
            float sn, cs;
            sincospif((float)(((unsigned)(n0 + i0) * (unsigned)m) & mask) * scale, &sn, &cs);   // exact re-sync
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float2 x = xs[i0 + 32 * k];
                ar = fmaf(x.x, cs, fmaf(x.y, sn, ar));          // x * (cs - j sn)
                ai = fmaf(x.y, cs, fmaf(-x.x, sn, ai));
                const float c2 = cs * stp_c - sn * stp_s, s2 = sn * stp_c + cs * stp_s;
                cs = c2; sn = s2;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ar += __shfl_xor_sync(0xffffffffu, ar, o);
            ai += __shfl_xor_sync(0xffffffffu, ai, o);
        }
        if (lane == 0) carr_add(vacc + (size_t)c * NBd * 2, l, (double)ar, (double)ai);
    }
    carr_tail(vacc + (size_t)c * NBd * 2, vticket + c, nchunk, NBd, carr + (size_t)c * NBd);
}

// The same partial spectrum through block moments.  One CTA per (1024-sample chunk, channel):
//   1. warp w takes blocks 4w .. 4w+3 of 32 samples: lane b holds sample b, t = (b - 15.5) / 16, the four complex
//      moments sum_b t^k bb are butterfly-summed and kept in shared memory;
//   2. one item = (quarter of 8 blocks, bin m): with phi = 32 pi m / N_c
//          sum_b bb[b] exp(-j 2 pi (n0 + b) m / N_c) = exp(-j 2 pi nc m / N_c) (M0 - j phi M1 - phi^2/2 M2 + j phi^3/6 M3) + O(phi^4/24)
//      (nc = n0 + 15.5 the block centre; its phase is reduced exactly in integers, (2 nc) m mod 2 N_c, for the first
//      block of the quarter and advanced by the exact 32-sample rotation for the other seven);
//   3. the four quarters are added in FP64.
// The launcher only takes this kernel when (2 pi (Wd + 1) 15.5 / N_c)^4 / 24 < 2e-8 and 2 N_c <= 2^24.
__global__ void DPE_SIDE256
k_carr_partial(const float2* __restrict__ xw, const float2* __restrict__ cc, const long long* __restrict__ dc_part,
               const int8_t* __restrict__ rs, const int32_t* __restrict__ idx_next,
               const int32_t* __restrict__ no_flip, const EpochDev* __restrict__ ep, int S, int Wd, int NBd,
               int n_fft, int nchunk, int n_dc, long long* __restrict__ vacc, unsigned int* __restrict__ vticket,
               double2* __restrict__ carr) {
    extern __shared__ float2 qpart[];                           // [4][NBd] quarter partials
    __shared__ float2 mom[kCarrChunk / 32][4];
    const int c = blockIdx.y;
    if (c >= ep->C) return;
    grid_dep_trigger();
    const int chunk = blockIdx.x, n0 = chunk * kCarrChunk;
    const bool flip = !no_flip[c];
    const int edge = idx_next[c];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float t = ((float)lane - 15.5f) * 0.0625f;
    const float2 mean = dc_mean(dc_part, n_dc, S);
#pragma unroll
    for (int bi = 0; bi < 4; ++bi) {
        const int blk = 4 * warp + bi;
        const int n = n0 + 32 * blk + lane;
        const float2 v = (n < S) ? baseband(xw, cc, rs, (size_t)c * S + n, mean, flip && n >= edge) : make_float2(0.f, 0.f);
        float m[8] = {v.x, v.y, t * v.x, t * v.y, 0.f, 0.f, 0.f, 0.f};
        m[4] = t * m[2]; m[5] = t * m[3]; m[6] = t * m[4]; m[7] = t * m[5];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int k = 0; k < 8; ++k) m[k] += __shfl_xor_sync(0xffffffffu, m[k], o);
        if (lane < 4) mom[blk][lane] = make_float2(lane == 0 ? m[0] : lane == 1 ? m[2] : lane == 2 ? m[4] : m[6],
                                                   lane == 0 ? m[1] : lane == 1 ? m[3] : lane == 2 ? m[5] : m[7]);
    }
    __syncthreads();
    const unsigned mask2 = 2u * (unsigned)n_fft - 1u;             // n_fft is a power of two
    const float inv_n = 1.0f / (float)n_fft;
    for (int item = threadIdx.x; item < 4 * NBd; item += blockDim.x) {
        const int q = item / NBd, l = item - q * NBd;
        const int m = l - Wd;                                   // bin relative to 0 Hz
        const float phi = (float)m * (32.0f * 3.14159265358979f * inv_n);
        const float a2 = 0.5f * phi * phi, a3 = phi * phi * (1.0f / 6.0f);
        float stp_s, stp_c, sn, cs;                             // 32-sample rotation; phase of the first block centre
        sincospif((float)((64u * (unsigned)m) & mask2) * inv_n, &stp_s, &stp_c);
        const unsigned twice_nc = 2u * (unsigned)(n0 + 256 * q) + 31u;
        sincospif((float)((twice_nc * (unsigned)m) & mask2) * inv_n, &sn, &cs);
        float ar = 0.f, ai = 0.f;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const float2* M = mom[8 * q + b];
            const float2 M0 = M[0], M1 = M[1], M2 = M[2], M3 = M[3];
            const float Ar = fmaf(-a2, M2.x, M0.x), Ai = fmaf(-a2, M2.y, M0.y);
            const float Br = fmaf(-a3, M3.x, M1.x), Bi = fmaf(-a3, M3.y, M1.y);
            const float Pr = fmaf(phi, Bi, Ar), Pi = fmaf(-phi, Br, Ai);          // A - j phi B
            ar = fmaf(Pr, cs, fmaf(Pi, sn, ar));                                   // P * (cs - j sn)
            ai = fmaf(Pi, cs, fmaf(-Pr, sn, ai));
            const float c2 = cs * stp_c - sn * stp_s, s2 = sn * stp_c + cs * stp_s;
            cs = c2; sn = s2;
        }
        qpart[item] = make_float2(ar, ai);
    }
    __syncthreads();
    for (int l = threadIdx.x; l < NBd; l += blockDim.x) {
        double re = 0, im = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) { re += (double)qpart[q * NBd + l].x; im += (double)qpart[q * NBd + l].y; }
        carr_add(vacc + (size_t)c * NBd * 2, l, re, im);
    }
    carr_tail(vacc + (size_t)c * NBd * 2, vticket + c, nchunk, NBd, carr + (size_t)c * NBd);
}

// arg-max over all block partials + BCM_MakeVelMeas (all threads of the last CTA)
// (weighted: BCM_ReduceAndVelMeas, batchcorrmanifold.cu:1658-1661 -- z = sum s v / sum s)
template <int kRound = 4>
__device__ __forceinline__ void finish_velocity(const double* __restrict__ blk_partial, int n_blk, const EpochDev& e,
                                                const double* __restrict__ vgrid_all, double* __restrict__ zval,
                                                double* __restrict__ rval, double* __restrict__ res, bool weighted = false) {
    double r[8];
    reduce_all_partials<kRound>(blk_partial, n_blk, r);
    if (threadIdx.x == 0) {
        const int64_t jm = (int64_t)r[6];
        const double* g = vgrid_all + 4 * jm;
        double z[4] = {e.R[0] * g[0] + e.R[1] * g[1] + e.R[2] * g[2] + e.center[4],
                       e.R[3] * g[0] + e.R[4] * g[1] + e.R[5] * g[2] + e.center[5],
                       e.R[6] * g[0] + e.R[7] * g[1] + e.R[8] * g[2] + e.center[6], g[3] + e.center[7]};
        if (weighted)
            for (int k = 0; k < 4; ++k) z[k] = r[k] / r[4];
        for (int k = 0; k < 4; ++k) { zval[4 + k] = z[k]; res[4 + k] = z[k]; }
        for (int rr = 4; rr < 8; ++rr)
            for (int k = 0; k < 8; ++k) rval[rr * 8 + k] = (rr == k) ? 1.0 : 0.0;
        res[12] = r[5]; res[13] = (double)jm; res[14] = r[7];
    }
}

// Per-channel terms of the Doppler geometry that do not depend on the candidate: unit line of sight to the grid CENTRE
// (batchcorrmanifold.cu:1917-1921: one per channel, not per candidate), and everything else of the reference's chain
//     rate = u . (v - v_sat);  f = F_L1 ((rate - drift_rx) / c + drift_sat) / sign;  idx = (N_c / fs) (f - f_i) + N_c / 2
// folded into ONE multiply-add of the candidate's  u . v - drift_rx :   idx = A (u . v - drift_rx) + B,
//     A = (N_c / fs) F_L1 / (c sign),   B = (N_c / fs) (F_L1 drift_sat / sign - f_i) + N_c / 2 - A (u . v_sat).
// 5 FP64 instructions per pair instead of 14.  Against the reference's order of operations the index moves by a few ulp of
// N_c / 2 -- 1e-10 of a bin -- and the lerp is continuous across a bin boundary: far inside the 1e-5 of the scores.
// `off` is the row offset N_c * c the reference adds before the floor, `lbase` the first bin of the window.
struct VelChan { double ux, uy, uz, A, B, off, lbase; };

__device__ __forceinline__ void vel_chan_consts(const EpochDev& e, const double* __restrict__ sat, int T, int n_fft, int Wd,
                                                double fs, VelChan* __restrict__ vc) {
    if (threadIdx.x < e.C) {
        const int c = threadIdx.x;
        const double* s = sat + ((size_t)c * T + T / 2) * 8;
        double los[3] = {s[0] - e.center[0], s[1] - e.center[1], s[2] - e.center[2]};
        const double range = norm(3, los);
        VelChan u;
        u.ux = los[0] / range; u.uy = los[1] / range; u.uz = los[2] / range;
        const double scale = n_fft / fs, inv_sign = 1.0 / e.doppler_sign;
        u.A = scale * K_F_L1 * inv_sign / K_C;
        u.B = scale * (K_F_L1 * s[7] * inv_sign - e.fi[c]) + n_fft / 2.0 - u.A * (u.ux * s[4] + u.uy * s[5] + u.uz * s[6]);
        u.off = (double)((int64_t)n_fft * c);
        u.lbase = u.off + (double)(n_fft / 2 - Wd);
        vc[c] = u;
    }
}

// Doppler bin of one (velocity candidate, channel) pair (batchcorrmanifold.cu:1932-1950): window entry l (bin l - Wd
// relative to 0 Hz), lerp weights of entries l + 1 and l; false when the pair falls outside the window / the spectrum.
// All integers here are below 2^53, so the window entry is formed in FP64 (exact) and converted once.
struct VelCand { double ex, ey, ez, pt; };
__device__ __forceinline__ bool vel_bin(const VelChan& u, const VelCand& v, double nf, int NBd, int* l_out, double* wg,
                                        double* wf) {
    const double dot = fma(u.ux, v.ex, fma(u.uy, v.ey, u.uz * v.ez));
    const double idx_base = fma(u.A, dot - v.pt, u.B);
    const bool valid = (idx_base < nf) && (idx_base > 0.0);
    const double idxo = idx_base + u.off;
    const double f = floor(idxo), gg = floor(idxo + 1.0);
    const double lrel = f - u.lbase;
    const bool ok = valid && lrel >= 0.0 && lrel <= (double)(NBd - 2);
    *l_out = ok ? (int)lrel : 0;
    *wg = idxo - f;
    *wf = gg - idxo;
    return ok;
}

__device__ __forceinline__ VelCand vel_cand(const EpochDev& e, const double* __restrict__ g4) {
    const double2 a = *reinterpret_cast<const double2*>(g4);
    const double2 b = *reinterpret_cast<const double2*>(g4 + 2);
    const double vx = e.R[0] * a.x + e.R[1] * a.y + e.R[2] * b.x + e.center[4];
    const double vy = e.R[3] * a.x + e.R[4] * a.y + e.R[5] * b.x + e.center[5];
    const double vz = e.R[6] * a.x + e.R[7] * a.y + e.R[8] * b.x + e.center[6];
    return VelCand{vx - K_OEDOT * e.center[1], vy + K_OEDOT * e.center[0], vz, b.y + e.center[7]};
}

// BCM_VelMeasML with the arg-max fused (batchcorrmanifold.cu:1896-1962); partial layout as the position kernels'
// (sum of scores, max, argmax, out-of-window).  kVelCand candidates per thread (j = base + tid + 128 k), branch-free over
// them so that their FP64 chains interleave -- one candidate per thread left the kernel at 45 us for 25^4 candidates,
// latency bound, with the per-CTA prologue (EpochDev copy, lines of sight) paid once per 128 candidates.
// WSUM: also the score-weighted sums of the ECEF velocity candidates (BCM_VelMeasReduction, batchcorrmanifold.cu:1090-1347:
// the same score, :1193-1197 the accumulation), for the weighted estimate.
// (7 per thread at 3 CTAs per SM, for an even 3 CTAs on every SM at 25^4 candidates, was slower: see k_score_lookup.)
constexpr int kVelCand = 6;
template <bool LP1, bool WSUM>
__global__ void __launch_bounds__(kReduceBlock, 4)
k_score_vel(const double* __restrict__ vgrid, const EpochDev* __restrict__ ep, const double* __restrict__ sat,
            const double2* __restrict__ carr, double fs, int n_fft, int Wd, int NBd, int T, int lpower, int64_t Gv,
            double* __restrict__ vscores, double* __restrict__ blk_partial, unsigned int* __restrict__ ticket,
            const double* __restrict__ vgrid_all, double* __restrict__ zval, double* __restrict__ rval,
            double* __restrict__ res) {
    __shared__ EpochDev e;
    __shared__ VelChan vch[DPE_MAX_CHAN];
    for (int i = threadIdx.x; i < (int)(sizeof(EpochDev) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t*>(&e)[i] = reinterpret_cast<const uint32_t*>(ep)[i];
    __syncthreads();
    vel_chan_consts(e, sat, T, n_fft, Wd, fs, vch);
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * (kReduceBlock * kVelCand) + threadIdx.x;
    VelCand vc[kVelCand];
    double score[kVelCand];
    bool act[kVelCand];
#pragma unroll
    for (int k = 0; k < kVelCand; ++k) {
        const int64_t j = base + (int64_t)k * kReduceBlock;
        act[k] = j < Gv;
        score[k] = 0.0;
        vc[k] = act[k] ? vel_cand(e, vgrid + 4 * j) : VelCand{0, 0, 0, 0};
    }
    const double nf = (double)n_fft;
    int oow = 0;
    grid_dep_wait();                                  // the carrier spectrum comes from the kernel before
    for (int c = 0; c < e.C; ++c) {
        const VelChan u = vch[c];
        const double2* __restrict__ cc = carr + (size_t)c * NBd;
#pragma unroll
        for (int k = 0; k < kVelCand; ++k) {
            int l;
            double wg, wf;
            const bool ok = vel_bin(u, vc[k], nf, NBd, &l, &wg, &wf);
            const double2 lo = cc[l], hi = cc[l + 1];                 // l = 0 when not ok: the loads are unconditional
            const double m = mag_pow_t<LP1>(hi.x * wg + lo.x * wf, hi.y * wg + lo.y * wf, lpower);
            score[k] += (ok && act[k]) ? m : 0.0;
            oow += (act[k] && !ok) ? 1 : 0;
        }
    }
    double v[5] = {0, 0, 0, 0, 0}, mx = -1.0, mi = 9.0e18;
#pragma unroll
    for (int k = 0; k < kVelCand; ++k) {
        if (!act[k]) continue;
        const int64_t j = base + (int64_t)k * kReduceBlock;
        vscores[j] = score[k];
        if (WSUM) {                                              // the ECEF velocity of the candidate (:1138-1141)
            v[0] += score[k] * (vc[k].ex + K_OEDOT * e.center[1]);
            v[1] += score[k] * (vc[k].ey - K_OEDOT * e.center[0]);
            v[2] += score[k] * vc[k].ez;
            v[3] += score[k] * vc[k].pt;
        }
        v[4] += score[k];
        if (score[k] > mx) { mx = score[k]; mi = (double)j; }    // increasing index order: the lowest index wins ties
    }
    block_reduce_store_vals<WSUM ? 5 : 1>(v, mx, mi, (double)oow, blk_partial);
    // last CTA: arg-max over all candidates + BCM_MakeVelMeas (zVal[4:8], RVal rows 4-7, batchcorrmanifold.cu:2030-2068)
    if (take_last_ticket(ticket)) finish_velocity(blk_partial, gridDim.x, e, vgrid_all, zval, rval, res, WSUM);
}

// =============================================================================================
// Brute-force velocity manifold (SURVEY.md section 8 a', last sentence).  By linearity the reference's lerp of two
// carrier-spectrum bins (batchcorrmanifold.cu:1950-1958) is a full-length correlation against a BLENDED carrier,
//     v(j, c) = sum_n bb_c[n] * ( (1 - a) w_m[n] + a w_{m+1}[n] ),   w_m[n] = exp(-j 2 pi n m / N_c),
// with the bin m and fraction a from the candidate's own Doppler geometry -- nothing is looked up, every
// (velocity candidate, PRN) pair runs all S samples, the same way k_brute does on the position grid.
//   k_vel_plane      bb = (x - mean) conj(carrier) * chosen replica, zero-padded to whole tiles
//   k_vel_pair_bins  bins of every pair + per-CTA (PRN, bin) histograms; then the shared sort tail
//                    (k_block_scan, k_scatter: dpe_brute.cu)
//   k_brute_vel      one warp = 32 pairs of one (PRN, bin) bucket in registers, lanes stride the samples of a
//                    shared 1024-sample tile; per candidate-sample 1 FFMA2 blends the carrier (alpha scalar,
//                    the (d, w) pair shared by all candidates) and 2 FFMA2 do the complex MAC (blended re / im
//                    as scalar operands, the sample pairs (re, im) and (-im, re) shared): 12 FLOP
//   k_score_vpairs   sum_prn |v|^L, arg-max, BCM_MakeVelMeas
// =============================================================================================
constexpr int kVelTile = 1024;
constexpr size_t kVelSmem = 2 * 2 * kVelTile * sizeof(float4);   // 64 KB: two buffers of (A, B)

__global__ void DPE_SIDE256
k_vel_plane(const float2* __restrict__ xw, const float2* __restrict__ cc, const long long* __restrict__ dc_part, int nchunk,
            const int8_t* __restrict__ rs, const int32_t* __restrict__ idx_next,
            const int32_t* __restrict__ no_flip, const EpochDev* __restrict__ ep, int S, int64_t S_pad,
            float2* __restrict__ vbb) {
    const int c = blockIdx.y;
    if (c >= ep->C) return;
    const float2 mean = dc_mean(dc_part, nchunk, S);            // before any thread leaves: the shuffles need whole warps
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= S_pad) return;
    vbb[(size_t)c * S_pad + n] = (n < S) ? baseband(xw, cc, rs, (size_t)c * S + n, mean, !no_flip[c] && n >= idx_next[c])
                                         : make_float2(0.f, 0.f);
}

__global__ void DPE_SIDE128
k_vel_pair_bins(const double* __restrict__ vgrid, const EpochDev* __restrict__ ep, const double* __restrict__ sat,
                double fs, int n_fft, int Wd, int NBd, int T, int64_t Gv, int16_t* __restrict__ pair_k,
                float* __restrict__ pair_a, float2* __restrict__ pair_v, int32_t* __restrict__ blk_hist) {
    extern __shared__ int32_t hs[];
    __shared__ EpochDev e;
    __shared__ VelChan vch[DPE_MAX_CHAN];
    for (int i = threadIdx.x; i < (int)(sizeof(EpochDev) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t*>(&e)[i] = reinterpret_cast<const uint32_t*>(ep)[i];
    __syncthreads();
    const int NB = 2 * Wd + 1, nbuck = e.C * NB;
    for (int i = threadIdx.x; i < nbuck; i += blockDim.x) hs[i] = 0;
    vel_chan_consts(e, sat, T, n_fft, Wd, fs, vch);
    __syncthreads();
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < Gv) {
        const VelCand vc = vel_cand(e, vgrid + 4 * j);
        const double nf = (double)n_fft;
        for (int c = 0; c < e.C; ++c) {
            int l;
            double wg, wf;
            const bool ok = vel_bin(vch[c], vc, nf, NBd, &l, &wg, &wf);
            pair_k[(size_t)c * Gv + j] = ok ? (int16_t)l : (int16_t)-1;
            pair_a[(size_t)c * Gv + j] = (float)wg;
            if (ok) atomicAdd(&hs[c * NB + l], 1);
            else pair_v[(size_t)c * Gv + j] = make_float2(__int_as_float(0x7fc00000), 0.f);   // NaN: "not scored"
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nbuck; i += blockDim.x) blk_hist[(size_t)i * gridDim.x + blockIdx.x] = hs[i];
}

// One CTA slot = kBfWarps groups of ONE (PRN, Doppler bin) bucket, so the two carriers of the bucket -- and their
// difference -- are the same for every pair of the CTA: the threads build them once per 1024-sample tile into shared
// memory next to the samples (thread i: samples i, i+256, i+512, i+768 -- the first from the exact integer phase
// n m mod N_c, the others by the exact 256-sample rotation), double-buffered, one __syncthreads per tile.  The inner
// loop is then 2 LDS.128 per sample and lane against 96 FFMA2.
//   A[n] = (w.re, w.im, d.re, d.im)   w = exp(-j theta_m[n]),  d = exp(-j theta_{m+1}[n]) - w
//   B[n] = (x.re, x.im, -x.im, x.re)  the sample and j times the sample
__global__ void __launch_bounds__(kBfWarps * 32, 1)
k_brute_vel(const float2* __restrict__ vbb, int64_t S_pad, const int4* __restrict__ hdr, const int32_t* __restrict__ ent_j,
            const float* __restrict__ ent_a, const int32_t* __restrict__ n_groups, float2* __restrict__ pair_v, int64_t Gv,
            int Wd, int n_fft) {
    extern __shared__ __align__(16) float4 vsm[];              // [2 buffers][A | B][kVelTile]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_slots = *n_groups / kBfWarps;
    const unsigned mask = (unsigned)n_fft - 1u;                 // n_fft is a power of two
    const float scale = 2.0f / (float)n_fft;
    const int ntiles = (int)(S_pad / kVelTile);
    constexpr int kPer = kVelTile / (kBfWarps * 32);            // samples per thread and tile
    for (int slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
        const int g = slot * kBfWarps + warp;
        const int4 h = hdr[g];                               // {channel, window entry, valid pairs}: one bucket per slot
        const int c = h.x, n_valid = h.z;
        const int m = h.y - Wd;                              // bin relative to 0 Hz; the pair blends bins m and m + 1
        const float a_mine = (lane < n_valid) ? ent_a[(size_t)g * kBfNC + lane] : 0.f;
        float al[kBfNC];
#pragma unroll
        for (int j = 0; j < kBfNC; ++j) al[j] = __shfl_sync(0xffffffffu, a_mine, j);
        float2 acc[kBfNC];
#pragma unroll
        for (int j = 0; j < kBfNC; ++j) acc[j] = make_float2(0.f, 0.f);
        float r0s, r0c, r1s, r1c;                            // rotation of 256 samples for bins m, m + 1
        sincospif((float)((256u * (unsigned)m) & mask) * scale, &r0s, &r0c);
        sincospif((float)((256u * (unsigned)(m + 1)) & mask) * scale, &r1s, &r1c);
        const float2* __restrict__ src = vbb + (size_t)c * S_pad;
        // staging of tile t, in two parts so that the global-load latency hides under the FFMA2 stream of the tile
        // before: the samples are loaded before the first half of tile t - 1 is computed, the carriers are built and
        // everything is stored between its two halves
        float2 xr[kPer];
        auto stage_load = [&](int t) {
#pragma unroll
            for (int k = 0; k < kPer; ++k) xr[k] = __ldg(src + (size_t)t * kVelTile + threadIdx.x + 256 * k);
        };
        auto stage_store = [&](int t, int buf) {
            float4* A = vsm + (size_t)buf * 2 * kVelTile;
            float4* B = A + kVelTile;
            const unsigned n = (unsigned)(t * kVelTile) + threadIdx.x;
            float s0, c0, s1, c1;                            // exact from the integer phase
            sincospif((float)((n * (unsigned)m) & mask) * scale, &s0, &c0);
            sincospif((float)((n * (unsigned)(m + 1)) & mask) * scale, &s1, &c1);
#pragma unroll
            for (int k = 0; k < kPer; ++k) {
                A[threadIdx.x + 256 * k] = make_float4(c0, -s0, c1 - c0, s0 - s1);
                B[threadIdx.x + 256 * k] = make_float4(xr[k].x, xr[k].y, -xr[k].y, xr[k].x);
                const float nc0 = c0 * r0c - s0 * r0s, ns0 = s0 * r0c + c0 * r0s;
                const float nc1 = c1 * r1c - s1 * r1s, ns1 = s1 * r1c + c1 * r1s;
                c0 = nc0; s0 = ns0; c1 = nc1; s1 = ns1;
            }
        };
        auto compute_half = [&](int t, int half) {
            const float4* A = vsm + (size_t)(t & 1) * 2 * kVelTile + lane + half * (kVelTile / 2);
            const float4* B = A + kVelTile;
#pragma unroll 4
            for (int i = 0; i < kVelTile / 64; ++i) {
                const float4 wd = A[32 * i], xx = B[32 * i];
                const float2 w = make_float2(wd.x, wd.y), d = make_float2(wd.z, wd.w);
                const float2 xa = make_float2(xx.x, xx.y), xb = make_float2(xx.z, xx.w);
#pragma unroll
                for (int j = 0; j < kBfNC; ++j) {
                    const float2 b = __ffma2_rn(make_float2(al[j], al[j]), d, w);   // w + alpha d
                    acc[j] = __ffma2_rn(make_float2(b.x, b.x), xa, acc[j]);          // x * b, real part of b
                    acc[j] = __ffma2_rn(make_float2(b.y, b.y), xb, acc[j]);          // ... imaginary part
                }
            }
        };
        __syncthreads();                                     // the previous slot is done with both buffers
        stage_load(0);
        stage_store(0, 0);
        __syncthreads();
        for (int t = 0; t < ntiles; ++t) {
            const bool more = t + 1 < ntiles;
            if (more) stage_load(t + 1);
            if (n_valid > 0) compute_half(t, 0);
            if (more) stage_store(t + 1, (t + 1) & 1);
            if (n_valid > 0) compute_half(t, 1);
            __syncthreads();                                 // tile t + 1 is staged, tile t is free
        }
        // lane partials -> one candidate per lane (halving butterfly, as in k_brute)
#pragma unroll
        for (int o = 16, nn = kBfNC / 2; o > 0; o >>= 1, nn >>= 1) {
            const bool up = (lane & o) != 0;
#pragma unroll
            for (int j = 0; j < nn; ++j) {
                const float2 keep = up ? acc[j + nn] : acc[j];
                const float2 send = up ? acc[j] : acc[j + nn];
                acc[j].x = keep.x + __shfl_xor_sync(0xffffffffu, send.x, o);
                acc[j].y = keep.y + __shfl_xor_sync(0xffffffffu, send.y, o);
            }
        }
        if (lane < n_valid) {
            const int64_t j = ent_j[(size_t)g * kBfNC + lane];
            pair_v[(size_t)c * Gv + j] = acc[0];
        }
    }
}

__global__ void __launch_bounds__(kReduceBlock, 6)
k_score_vpairs(const double* __restrict__ vgrid, const EpochDev* __restrict__ ep, const float2* __restrict__ pair_v,
               int lpower, int64_t Gv, double* __restrict__ vscores, double* __restrict__ blk_partial,
               unsigned int* __restrict__ ticket, double* __restrict__ zval, double* __restrict__ rval,
               double* __restrict__ res) {
    const EpochDev& e = *ep;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = j < Gv;
    double score = 0.0;
    int oow = 0;
    if (active) {
        for (int c0 = 0; c0 < e.C; c0 += 4) {
            float2 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
                v[q] = (c0 + q < e.C) ? __ldcs(&pair_v[(size_t)(c0 + q) * Gv + j]) : make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (c0 + q >= e.C) break;
                if (v[q].x == v[q].x) score += mag_pow((double)v[q].x, (double)v[q].y, lpower);
                else ++oow;
            }
        }
        vscores[j] = score;
    }
    double v5[5] = {0, 0, 0, 0, active ? score : 0.0};
    block_reduce_store_vals<1>(v5, active ? score : -1.0, active ? (double)j : 9.0e18, (double)oow, blk_partial);
    if (take_last_ticket(ticket)) finish_velocity<2>(blk_partial, gridDim.x, e, vgrid, zval, rval, res);
}

int vel_brute_set_attributes(dpe_ctx* c) {   // per context (device); not while a stream capture is open
    DPE_CUDA(cudaFuncSetAttribute(k_brute_vel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kVelSmem));
    c->vel_attr_set = 1;
    return DPE_OK;
}

int launch_score_vel_brute(dpe_ctx* c, cudaStream_t s) {
    const int S = (int)c->S, C = c->epoch_C;
    const int NB = 2 * c->Wd + 1, nbuck = C * NB;
    prof_begin(c, DPE_STAGE_VELOCITY, s);
    dim3 gp((unsigned)((c->vS_pad + 255) / 256), C);
    k_vel_plane<<<gp, 256, 0, s>>>(c->xw, c->bb, c->dc_part, c->nchunk, c->rs, c->idx_next, c->no_flip, c->ep, S, c->vS_pad,
                                   c->vbb);
    c->launches++;
    const int nblk = (int)((c->Gv + kSortBlock - 1) / kSortBlock);
    k_vel_pair_bins<<<nblk, kSortBlock, sizeof(int32_t) * nbuck, s>>>(c->vgrid, c->ep, c->sat, c->cfg.fs, c->n_fft, c->Wd,
                                                                     c->NBd, c->T, c->Gv, c->vpair_k, c->vpair_a, c->vpair_v,
                                                                     c->vblk_hist);
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    SortLists L = {c->vpair_k, c->vpair_a, c->vblk_hist, c->vhist, c->vgroup_base, c->vbucket_base, c->vn_groups, c->vhdr,
                   c->vent_j, c->vent_a, c->vmax_groups};
    int rc = launch_sort_tail(c, L, c->Gv, c->Wd, C, c->ticket + 2, s);
    if (rc) return rc;
    if (!c->vel_attr_set) { int rc2 = vel_brute_set_attributes(c); if (rc2) return rc2; }
    k_brute_vel<<<c->sm_count, kBfWarps * 32, kVelSmem, s>>>(c->vbb, c->vS_pad, reinterpret_cast<const int4*>(c->vhdr), c->vent_j,
                                                      c->vent_a, c->vn_groups, c->vpair_v, c->Gv, c->Wd, c->n_fft);
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    const int nb2 = (int)((c->Gv + kReduceBlock - 1) / kReduceBlock);
    k_score_vpairs<<<nb2, kReduceBlock, 0, s>>>(c->vgrid, c->ep, c->vpair_v, c->cfg.lpower, c->Gv, c->vscores,
                                                c->vblk_partial, c->ticket + 3, c->zval, c->rval, c->result);
    c->launches++;
    prof_end(c, s);
    DPE_CUDA(cudaGetLastError());
    return DPE_OK;
}

int launch_score_vel(dpe_ctx* c, cudaStream_t s) {
    const int S = (int)c->S, C = c->epoch_C;
    prof_begin(c, DPE_STAGE_VELOCITY, s);
    dim3 g2(c->vnchunk, C);
    // block moments when the Doppler window is narrow enough for the 4-term Taylor bound (and the doubled phase stays
    // exact in a float), the direct evaluation otherwise
    const double x = 2.0 * 3.141592653589793 * (c->Wd + 1) * 15.5 / (double)c->n_fft;
    if (x * x * x * x / 24.0 < 2.0e-8 && c->n_fft <= (1 << 23) && !c->carr_direct)
        k_carr_partial<<<g2, 256, sizeof(float2) * 4 * c->NBd, s>>>(c->xw, c->bb, c->dc_part, c->rs, c->idx_next, c->no_flip,
                                                                   c->ep, S, c->Wd, c->NBd, c->n_fft, c->vnchunk, c->nchunk, c->vacc,
                                                                   c->vticket, c->carr);
    else
        k_carr_partial_direct<<<g2, 256, 0, s>>>(c->xw, c->bb, c->dc_part, c->rs, c->idx_next, c->no_flip, c->ep, S, c->Wd,
                                                 c->NBd, c->n_fft, c->vnchunk, c->nchunk, c->vacc, c->vticket, c->carr);
    const int nblk = (int)((c->Gv + kReduceBlock * kVelCand - 1) / (kReduceBlock * kVelCand));
#define DPE_VEL_ARGS c->vgrid, c->ep, c->sat, c->carr, c->cfg.fs, c->n_fft, c->Wd, c->NBd, c->T, c->cfg.lpower, c->Gv, c->vscores, \
                     c->vblk_partial, c->ticket + 3, c->vgrid, c->zval, c->rval, c->result
#define DPE_VEL_LAUNCH(LP, WS) launch_dep(k_score_vel<LP, WS>, nblk, kReduceBlock, 0, s, c->use_pdl != 0, DPE_VEL_ARGS)
    if (c->vel_weighted) {
        if (c->cfg.lpower == 1) DPE_VEL_LAUNCH(true, true); else DPE_VEL_LAUNCH(false, true);
    } else {
        if (c->cfg.lpower == 1) DPE_VEL_LAUNCH(true, false); else DPE_VEL_LAUNCH(false, false);
    }
#undef DPE_VEL_LAUNCH
#undef DPE_VEL_ARGS
    c->launches += 2;
    prof_end(c, s);
    DPE_CUDA(cudaGetLastError());
    return DPE_OK;
}

int kernel_attr_vel(const char* name, cudaFuncAttributes* a) {
    DPE_KATTR("k_carr_partial", k_carr_partial);
    DPE_KATTR("k_carr_partial_direct", k_carr_partial_direct);
    DPE_KATTR("k_score_vel", (k_score_vel<true, false>));
    DPE_KATTR("k_brute_vel", k_brute_vel);
    DPE_KATTR("k_vel_pair_bins", k_vel_pair_bins);
    DPE_KATTR("k_vel_plane", k_vel_plane);
    DPE_KATTR("k_score_vpairs", k_score_vpairs);
    return 0;
}

}  // namespace dpe
