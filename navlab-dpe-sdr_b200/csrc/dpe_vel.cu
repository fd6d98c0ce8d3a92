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
// reduced exactly in integers (n*m mod N_c) before sincospif; FP32 inside a 1024-sample chunk,
// FP64 across chunks.  The scoring kernel is the position one with a Doppler geometry.
#include "dpe_geom.cuh"

namespace dpe {

// integer sum of the block (exact; the reference reduces doubles with thrust, :1065)
__global__ void DPE_SIDE256 k_dc_sum(const int16_t* __restrict__ iq, int S, long long* __restrict__ out) {
    long long si = 0, sq = 0;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < S; n += gridDim.x * blockDim.x) {
        const short2 v = reinterpret_cast<const short2*>(iq)[n];
        si += v.x; sq += v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        si += __shfl_xor_sync(0xffffffffu, si, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(reinterpret_cast<unsigned long long*>(out), (unsigned long long)si);
        atomicAdd(reinterpret_cast<unsigned long long*>(out + 1), (unsigned long long)sq);
    }
}

// One CTA per (1024-sample chunk, channel); warp w owns bins w, w+8, ...; lanes stride the samples.
// bb[n] = zw[n] * chosen replica[n]  with zw = (x - mean) * conj(carrier) from k_prepare
// (BCS_SubtractDCOffset :470-485, BCS_ChoosyBatchMultiplyAndPad :422-452).
// Twiddles: lane handles samples n0 + lane + 32 i; exp(-j 2 pi n m / N_c) is evaluated exactly
// (integer n*m mod N_c -> sincospif) every 8th step and advanced by the exact 32-sample rotation in
// between (7 complex multiplies: < 5e-7 relative drift).
__global__ void DPE_SIDE256
k_carr_partial(const float2* __restrict__ zw, const int8_t* __restrict__ rs, const int32_t* __restrict__ idx_next,
               const int32_t* __restrict__ no_flip, const EpochDev* __restrict__ ep, int S, int Wd, int NBd,
               int n_fft, int nchunk, double2* __restrict__ vpart) {
    __shared__ float2 xs[kCorrChunk];
    const int c = blockIdx.y;
    if (c >= ep->C) return;
    const int chunk = blockIdx.x, n0 = chunk * kCorrChunk;
    const bool flip = !no_flip[c];
    const int edge = idx_next[c];
    for (int i = threadIdx.x; i < kCorrChunk; i += blockDim.x) {
        const int n = n0 + i;
        float2 v = make_float2(0.f, 0.f);
        if (n < S) {
            v = zw[(size_t)c * S + n];
            float r = (float)rs[(size_t)c * S + n];
            if (flip && n >= edge) r = -r;
            v.x *= r; v.y *= r;
        }
        xs[i] = v;
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
        for (int i0 = lane; i0 < kCorrChunk; i0 += 32 * 8) {
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
        if (lane == 0) vpart[((size_t)c * nchunk + chunk) * NBd + l] = make_double2((double)ar, (double)ai);
    }
}

// one warp per bin, lanes stride the chunks, xor-tree (fixed order)
__global__ void DPE_SIDE256
k_carr_finalize(const double2* __restrict__ vpart, const EpochDev* __restrict__ ep, int NBd,
                int nchunk, double2* __restrict__ carr) {
    const int c = blockIdx.x;
    if (c >= ep->C) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int l = blockIdx.y * 8 + warp;
    if (l >= NBd) return;
    double re = 0, im = 0;
    for (int ch = lane; ch < nchunk; ch += 32) {
        const double2 p = vpart[((size_t)c * nchunk + ch) * NBd + l];
        re += p.x; im += p.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        re += __shfl_xor_sync(0xffffffffu, re, o);
        im += __shfl_xor_sync(0xffffffffu, im, o);
    }
    if (lane == 0) carr[(size_t)c * NBd + l] = make_double2(re, im);
}

// BCM_VelMeasML with the arg-max fused (batchcorrmanifold.cu:1896-1962); partial layout as the
// position kernels' (sum s*v, sum s, max, argmax, out-of-window).
__global__ void __launch_bounds__(kReduceBlock, 6)
k_score_vel(const double* __restrict__ vgrid, const EpochDev* __restrict__ ep, const double* __restrict__ sat,
            const double2* __restrict__ carr, double fs, int n_fft, int Wd, int NBd, int T, int lpower, int64_t Gv,
            double* __restrict__ vscores, double* __restrict__ blk_partial, unsigned int* __restrict__ ticket,
            const double* __restrict__ vgrid_all, double* __restrict__ zval, double* __restrict__ rval,
            double* __restrict__ res) {
    __shared__ EpochDev e;
    __shared__ double los_s[DPE_MAX_CHAN][8];
    for (int i = threadIdx.x; i < (int)(sizeof(EpochDev) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t*>(&e)[i] = reinterpret_cast<const uint32_t*>(ep)[i];
    __syncthreads();
    // the line of sight goes to the grid CENTRE (batchcorrmanifold.cu:1917-1921): one per channel, not per candidate
    if (threadIdx.x < e.C) {
        const int c = threadIdx.x;
        const double* s = sat + ((size_t)c * T + T / 2) * 8;
        double los[3] = {s[0] - e.center[0], s[1] - e.center[1], s[2] - e.center[2]};
        const double range = norm(3, los);
        los_s[c][0] = los[0] / range; los_s[c][1] = los[1] / range; los_s[c][2] = los[2] / range;
        los_s[c][3] = s[4]; los_s[c][4] = s[5]; los_s[c][5] = s[6]; los_s[c][6] = s[7];
    }
    __syncthreads();
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = j < Gv;
    double score = 0.0;
    int oow = 0;
    Cand v = {0, 0, 0, 0};
    if (active) {
        const double* g = vgrid + 4 * j;
        v.px = e.R[0] * g[0] + e.R[1] * g[1] + e.R[2] * g[2] + e.center[4];
        v.py = e.R[3] * g[0] + e.R[4] * g[1] + e.R[5] * g[2] + e.center[5];
        v.pz = e.R[6] * g[0] + e.R[7] * g[1] + e.R[8] * g[2] + e.center[6];
        v.pt = g[3] + e.center[7];
        const double ex = v.px - K_OEDOT * e.center[1], ey = v.py + K_OEDOT * e.center[0], ez = v.pz;
        for (int c = 0; c < e.C; ++c) {
            const double* u = los_s[c];                          // unit LOS to the centre + satellite velocity / drift
            const double rate = (u[0] * (ex - u[3])) + (u[1] * (ey - u[4])) + (u[2] * (ez - u[5]));
            const double bc_fi = K_F_L1 * ((rate - v.pt) / K_C + u[6]) / e.doppler_sign;
            const double fi0 = bc_fi - e.fi[c];
            const double idx_base = (n_fft / fs) * fi0 + n_fft / 2.0;
            const bool valid = (idx_base < n_fft) && (idx_base > 0);
            const double idxo = idx_base + (double)((int64_t)n_fft * c);
            const double f = floor(idxo), gg = floor(idxo + 1.0);
            const int64_t l = (int64_t)f - (int64_t)n_fft * c - n_fft / 2 + Wd;
            if (valid && l >= 0 && l + 1 < NBd) {
                const double2 lo = carr[(size_t)c * NBd + l], hi = carr[(size_t)c * NBd + l + 1];
                const double wg = idxo - f, wf = gg - idxo;
                score += mag_pow(hi.x * wg + lo.x * wf, hi.y * wg + lo.y * wf, lpower);
            } else {
                ++oow;
            }
        }
        vscores[j] = score;
    }
    block_reduce_store(score, j, v, active, oow, blk_partial);
    // last CTA: arg-max over all candidates + BCM_MakeVelMeas (zVal[4:8], RVal rows 4-7, batchcorrmanifold.cu:2030-2068)
    if (take_last_ticket(ticket)) {
        double r[8];
        reduce_all_partials(blk_partial, gridDim.x, r);
        if (threadIdx.x == 0) {
            const int64_t jm = (int64_t)r[6];
            const double* g = vgrid_all + 4 * jm;
            const double z[4] = {e.R[0] * g[0] + e.R[1] * g[1] + e.R[2] * g[2] + e.center[4],
                                 e.R[3] * g[0] + e.R[4] * g[1] + e.R[5] * g[2] + e.center[5],
                                 e.R[6] * g[0] + e.R[7] * g[1] + e.R[8] * g[2] + e.center[6], g[3] + e.center[7]};
            for (int k = 0; k < 4; ++k) { zval[4 + k] = z[k]; res[4 + k] = z[k]; }
            for (int rr = 4; rr < 8; ++rr)
                for (int k = 0; k < 8; ++k) rval[rr * 8 + k] = (rr == k) ? 1.0 : 0.0;
            res[12] = r[5]; res[13] = (double)jm; res[14] = r[7];
        }
    }
}

// DC sum of the block; runs before k_prepare when a velocity grid exists (k_prepare then also
// emits zw = (x - mean) * conj(carrier))
int launch_dc_sum(dpe_ctx* c, cudaStream_t s) {
    DPE_CUDA(cudaMemsetAsync(c->dc_sum, 0, 2 * sizeof(long long), s));
    k_dc_sum<<<32, 256, 0, s>>>(c->iq, (int)c->S, c->dc_sum);
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    return DPE_OK;
}

int launch_score_vel(dpe_ctx* c, cudaStream_t s) {
    const int S = (int)c->S, C = c->epoch_C;
    prof_begin(c, DPE_STAGE_VELOCITY, s);
    dim3 g2(c->nchunk, C);
    k_carr_partial<<<g2, 256, 0, s>>>(c->bb, c->rs, c->idx_next, c->no_flip, c->ep, S, c->Wd, c->NBd, c->n_fft,
                                      c->nchunk, c->vpart);
    dim3 g3(C, (c->NBd + 7) / 8);
    k_carr_finalize<<<g3, 256, 0, s>>>(c->vpart, c->ep, c->NBd, c->nchunk, c->carr);
    const int nblk = (int)((c->Gv + kReduceBlock - 1) / kReduceBlock);
    k_score_vel<<<nblk, kReduceBlock, 0, s>>>(c->vgrid, c->ep, c->sat, c->carr, c->cfg.fs, c->n_fft, c->Wd, c->NBd,
                                              c->T, c->cfg.lpower, c->Gv, c->vscores, c->vblk_partial, c->ticket,
                                              c->vgrid, c->zval, c->rval, c->result);
    c->launches += 3;
    prof_end(c, s);
    DPE_CUDA(cudaGetLastError());
    return DPE_OK;
}

int kernel_attr_vel(const char* name, cudaFuncAttributes* a) {
    DPE_KATTR("k_dc_sum", k_dc_sum);
    DPE_KATTR("k_carr_partial", k_carr_partial);
    DPE_KATTR("k_carr_finalize", k_carr_finalize);
    DPE_KATTR("k_score_vel", k_score_vel);
    return 0;
}

}  // namespace dpe
