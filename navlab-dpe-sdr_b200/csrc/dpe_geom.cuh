// dpe_geom.cuh -- device helpers shared by the lookup and brute-force scoring
// kernels: candidate ENU->ECEF, geometric back-calculation of the code-phase
// bin (cudarecv/modules/src/batchcorrmanifold.cu:1760-1812) and the
// deterministic block reduction.
#pragma once
#include "dpe_internal.cuh"

namespace dpe {

struct Cand { double px, py, pz, pt; };
struct CandRel { double dx, dy, dz, d2; };      // candidate minus grid centre (ECEF), squared length

// batchcorrmanifold.cu:1760-1763 (same operation order: the rotated offset first, the centre added last)
__device__ __forceinline__ Cand cand_ecef(const EpochDev& e, const double* __restrict__ g4, CandRel* rel = nullptr) {
    const double2 a = *reinterpret_cast<const double2*>(g4);
    const double2 b = *reinterpret_cast<const double2*>(g4 + 2);
    const double dx = e.R[0] * a.x + e.R[1] * a.y + e.R[2] * b.x;
    const double dy = e.R[3] * a.x + e.R[4] * a.y + e.R[5] * b.x;
    const double dz = e.R[6] * a.x + e.R[7] * a.y + e.R[8] * b.x;
    Cand p;
    p.px = dx + e.center[0];
    p.py = dy + e.center[1];
    p.pz = dz + e.center[2];
    p.pt = b.y + e.center[3];
    if (rel) { rel->dx = dx; rel->dy = dy; rel->dz = dz; rel->d2 = dx * dx + dy * dy + dz * dz; }
    return p;
}

// Per-channel terms of the back-calculation that do not depend on the candidate, evaluated once per
// CTA with the reference's own operations (so the bins stay bit-identical): fs / fc (one FP64 division
// per pair otherwise), the two integer -> double conversions and (cpEnd - cpRef) * T_CA.
struct ChanConst { double ratio, tow, cpd; };

__device__ __forceinline__ void chan_consts(const EpochDev& e, double fs, ChanConst* __restrict__ cc) {
    for (int c = threadIdx.x; c < e.C; c += blockDim.x) {
        cc[c].ratio = fs / e.fc[c];
        cc[c].tow = (double)e.cp_ref_tow[c];
        cc[c].cpd = (e.cp_end[c] - e.cp_ref[c]) * K_T_CA;
    }
}

// batchcorrmanifold.cu:1779-1791: geometric back-calculation of the code phase
// and its (fractional) fft-shifted correlogram index for channel c.
__device__ __forceinline__ double code_index(const EpochDev& e, const ChanConst& k, const Cand& p,
                                             const double* __restrict__ sat, int c, double S) {
    double los[3] = {sat[0] - p.px, sat[1] - p.py, sat[2] - p.pz};
    const double range = norm(3, los);
    const double pr = range - K_C * sat[3] + p.pt;
    const double tx = e.rx_time - pr / K_C;
    const double frac = tx - k.tow - k.cpd;
    const double bc_rc = frac * K_F_CA;
    const double rc0 = bc_rc - e.rc_end[c];
    return k.ratio * (-rc0) + S / 2.0;
}

// ---------------------------------------------------------------------------------------------
// The same bin without the square root, the division and norm(): centre-relative geometry.
//
// With d0 = sat - centre, rho0 = |d0|, u = d0 / rho0 and the candidate offset D = p - centre,
//     |sat - p| = rho0 - a + (b - a^2) / (2 rho0) * (1 + a / rho0) + O(b^2 / rho0^3),   a = u.D,  b = D.D
// -- for |D| <= 1 km and rho0 >= 2e7 m the neglected terms are below 1e-11 m, so this range agrees with
// the reference's norm(3, los) to FP64 rounding (a few 1e-9 m).  That alone would not keep the bins
// bit-exact: the reference's  tx = rxTime - pr / c  rounds to the ulp of rxTime (5.8e-11 s = 1.7 cm of
// pseudorange = 1.5e-4 samples of index) -- every later operation of the chain is exact or far finer.
// So the fast path computes the SAME rounded tx whenever the exact difference is provably on the same
// side of every rounding boundary (|rounding error of the subtraction| < half an ulp minus a margin of
// 4e-16 s = 1.2e-7 m, 5x the error bound of the range above, the multiplication by 1/c included), and
// from tx on it repeats the reference's operations one by one: the index, hence the bin, the lerp weights
// and the score, are bit-identical to code_index().  When the margin is not met (probability ~1.4e-5 per
// pair) the pair goes through code_index() itself.  Checked exhaustively against the FP64 chain:
// tests/test_gpu_parity.py::test_code_phase_bins_bit_exact, test_fast_geometry_equals_the_reference_chain.
// ---------------------------------------------------------------------------------------------
struct SatGeo { double rho, ux, uy, uz, half_inv_rho, inv_rho, sat_dt; };

__device__ __forceinline__ SatGeo make_sat_geo(const EpochDev& e, const double* __restrict__ sat) {
    SatGeo g;
    const double dx = sat[0] - e.center[0], dy = sat[1] - e.center[1], dz = sat[2] - e.center[2];
    g.rho = sqrt(dx * dx + dy * dy + dz * dz);
    g.inv_rho = 1.0 / g.rho;
    g.ux = dx * g.inv_rho; g.uy = dy * g.inv_rho; g.uz = dz * g.inv_rho;
    g.half_inv_rho = 0.5 * g.inv_rho;
    g.sat_dt = sat[3];
    return g;
}

constexpr double kTxMargin = 4.0e-16;           // seconds; see above

// returns false when the pair must take the exact chain (too close to a rounding boundary of tx)
__device__ __forceinline__ bool code_index_fast_core(double rx_time, const ChanConst& k, const SatGeo& g, double pt,
                                                     const CandRel& r, double rc_end, double S, double* idx) {
    const double a = g.ux * r.dx + g.uy * r.dy + g.uz * r.dz;
    const double range = (g.rho - a) + (fma(-a, a, r.d2) * g.half_inv_rho) * fma(a, g.inv_rho, 1.0);
    const double pr = range - K_C * g.sat_dt + pt;
    const double q = pr * (1.0 / K_C);
    const double tx = rx_time - q;
    const double err = (rx_time - tx) - q;                         // exact rounding error of the subtraction above
    // half an ulp of tx: 2^(exponent - 53)
    const double half_ulp = __longlong_as_double((__double_as_longlong(tx) & 0x7ff0000000000000ll) - (53ll << 52));
    const double frac = tx - k.tow - k.cpd;
    const double bc_rc = frac * K_F_CA;
    const double rc0 = bc_rc - rc_end;
    *idx = k.ratio * (-rc0) + S / 2.0;
    return fabs(err) < half_ulp - kTxMargin;
}

__device__ __forceinline__ double code_index_fast(const EpochDev& e, const ChanConst& k, const SatGeo& g, const Cand& p,
                                                  const CandRel& r, const double* __restrict__ sat, int c, double S) {
    double idx;
    if (code_index_fast_core(e.rx_time, k, g, p.pt, r, e.rc_end[c], S, &idx)) return idx;
    return code_index(e, k, p, sat, c, S);
}

// (DPE_SAT_PER_TIME reads a per-(channel, time index) SatGeo table built by k_sat_geo, dpe_score.cu; the arg-max
// kernels build the C entries of the middle time index per CTA instead)

struct Bin { int64_t f; double wf, wg; int l; bool ok; };   // v = cs[l+1]*wg + cs[l]*wf

// The same bin for the scoring kernels: the row offset S * c and the first bin of the window are formed once per launch
// (all integers here are below 2^53, so the FP64 differences are exact and the window entry needs ONE conversion --
// make_bin's int64 chain was 8 of the ~100 instructions per pair).  Identical l / wg / wf / ok by construction;
// k_debug_bins keeps make_bin, and test_code_phase_bins_bit_exact compares the scores' bins with it.
struct BinFast { double wf, wg; int l; bool ok; };
__device__ __forceinline__ BinFast make_bin_fast(double idx_base, int c, int S, int W) {
    const double off = (double)((int64_t)S * c);                 // loop-invariant per channel: hoisted
    const double lbase = off + (double)(S / 2 - W);
    BinFast b;
    const bool valid = (idx_base < (double)S) && (idx_base > 0.0);
    const double idxo = idx_base + off;
    const double f = floor(idxo), g = floor(idxo + 1.0);
    b.wg = idxo - f;
    b.wf = g - idxo;
    b.l = (int)(f - lbase);                                      // saturating; NaN -> 0 with valid = false
    b.ok = valid && b.l >= 0 && b.l <= 2 * W;
    return b;
}

// batchcorrmanifold.cu:1795-1812 (row offset S*chan added before the floor, as there)
__device__ __forceinline__ Bin make_bin(double idx_base, int c, int S, int W) {
    Bin b;
    const bool valid = (idx_base < (double)S) && (idx_base > 0.0);
    const double idxo = idx_base + (double)((int64_t)S * c);
    const double f = floor(idxo), g = floor(idxo + 1.0);
    b.f = (int64_t)f;
    b.wg = idxo - f;
    b.wf = g - idxo;
    const int64_t l = b.f - (int64_t)S * c - S / 2 + W;
    b.l = (int)l;
    b.ok = valid && l >= 0 && l <= 2 * (int64_t)W;
    return b;
}

// sqrt(s) from the FP32 reciprocal square root (MUFU.RSQ) and ONE FP64 Newton step: the seed is good to 1.2e-7, the
// step squares that -- 2e-14 relative, against 1.1e-16 for the IEEE sqrt (7 more FP64 instructions and a conditional
// call per pair on a part whose FP64 pipe is what bounds the scoring kernels).  Branch-free, so that the chains of a
// thread's candidates interleave: s = 0 gives 0; below 1e-30 (a squared correlation of int16 samples is either 0 or
// far above) the clamped seed only bounds the result by 1e-15; above 1e37 is out of reach (|corr| < 1e11).
__device__ __forceinline__ double sqrt_newton(double s) {
    const double y = (double)rsqrtf(fmaxf((float)s, 1.0e-30f));
    const double m = s * y;
    return fma(0.5 * y, fma(-m, m, s), m);
}

// LP1 = true: the exponent is known to be 1 (the reference's default LPower): no test, no pow() call in the instruction
// stream -- the kernels that score several candidates per thread are instantiated for it, so that nothing but
// straight-line code separates the chains of the candidates
template <bool LP1>
__device__ __forceinline__ double mag_pow_t(double re, double im, int L);

__device__ __forceinline__ double mag_pow(double re, double im, int L) {
    const double m = sqrt_newton(re * re + im * im);
    if (L == 1) return m;
    if (L == 2) return m * m;
    return pow(m, (double)L);
}

template <bool LP1>
__device__ __forceinline__ double mag_pow_t(double re, double im, int L) {
    if (LP1) return sqrt_newton(re * re + im * im);
    return mag_pow(re, im, L);
}

// Block-level reduction of (score-weighted sums, sum, max/argmax, out-of-window)
// with warp shuffles + shared memory, fixed order => bit-reproducible.
// out[0..3]=sum s*p, [4]=sum s, [5]=max, [6]=global argmax (lowest on ties), [7]=oow
// NSUM = 5: score-weighted sums + sum of scores; NSUM = 1: the sum of scores only (arg-max estimate): v[4] alone is reduced
template <int NSUM>
__device__ __forceinline__ void block_reduce_store_vals(double (&v)[5], double mx, double mi, double oo,
                                                        double* __restrict__ out);

__device__ __forceinline__ void block_reduce_store(double score, int64_t gidx, const Cand& p, bool active,
                                                   int oow, double* __restrict__ out) {
    double v[5] = {0, 0, 0, 0, 0};
    double mx = -1.0, mi = 9.0e18;
    if (active) {
        v[0] = score * p.px; v[1] = score * p.py; v[2] = score * p.pz; v[3] = score * p.pt; v[4] = score;
        mx = score; mi = (double)gidx;
    }
    block_reduce_store_vals<5>(v, mx, mi, (double)oow, out);
}

template <int NSUM>
__device__ __forceinline__ void block_reduce_store_vals(double (&v)[5], double mx, double mi, double oo,
                                                        double* __restrict__ out) {
    __shared__ double sh[kReduceBlock / 32][8];
    constexpr int K0 = (NSUM == 5) ? 0 : 4;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = K0; k < 5; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        oo += __shfl_xor_sync(0xffffffffu, oo, o);
        const double omx = __shfl_xor_sync(0xffffffffu, mx, o);
        const double omi = __shfl_xor_sync(0xffffffffu, mi, o);
        if (omx > mx || (omx == mx && omi < mi)) { mx = omx; mi = omi; }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 5; ++k) sh[warp][k] = v[k];
        sh[warp][5] = mx; sh[warp][6] = mi; sh[warp][7] = oo;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = sh[0][k];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
#pragma unroll
            for (int k = 0; k < 5; ++k) r[k] += sh[w][k];
            r[7] += sh[w][7];
            if (sh[w][5] > r[5] || (sh[w][5] == r[5] && sh[w][6] < r[6])) { r[5] = sh[w][5]; r[6] = sh[w][6]; }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) out[(size_t)blockIdx.x * 8 + k] = r[k];
    }
}


// ---------------------------------------------------------------------------------------------
// Grid-level reduction without a second launch: the CTA that takes the last ticket reduces all
// block partials in a fixed order (so the result does not depend on which CTA happens to be last)
// and resets the ticket counter for the next launch.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool take_last_ticket(unsigned int* counter) {
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        // thread 0 wrote this CTA's partial (block_reduce_store / the bucket total): its fence makes it visible
        // device-wide before the ticket.  (A fence in EVERY thread, as in round 1, stalled each warp on its own
        // score store: "membar" was 0.7 stalled warps per issue in k_score_lookup.)
        __threadfence();
        const unsigned int t = atomicAdd(counter, 1u);
        s_last = (t == gridDim.x - 1);
        if (s_last) *counter = 0;
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last;
}

// all threads of one CTA (a power of two, <= kReduceBlock): r[0..4] sums, r[5]/r[6] max / arg-max (lowest index on ties), r[7] sum;
// valid in thread 0 on return
template <int kRound = 4>
__device__ __forceinline__ void reduce_all_partials(const double* __restrict__ blk, int n_blk, double (&r)[8]) {
    __shared__ double shr[kReduceBlock][8];
    r[0] = r[1] = r[2] = r[3] = r[4] = 0.0; r[5] = -1.0; r[6] = 9.0e18; r[7] = 0.0;
    // rounds of kRound blocks per thread (4; 2 in the 64-register pair kernels), all loads of a round issued before the first sum: the partials come from other SMs
    // (L2, ~0.7 us a round trip), and a loop that the compiler could not unroll (the trip count is dynamic) paid that once per
    // block -- 3 of the 5 us the last CTA of k_score_lookup spent here (profiles/r02ae_phase_stamps.txt).  Same order of
    // summation as before: block b0 + k * blockDim + tid for k = 0, 1, 2, ...
    for (int b0 = 0; b0 < n_blk; b0 += kRound * (int)blockDim.x) {
        double2 v[kRound][4];
        bool ok[kRound];
#pragma unroll
        for (int k = 0; k < kRound; ++k) {
            const int b = b0 + k * (int)blockDim.x + (int)threadIdx.x;
            ok[k] = b < n_blk;
            const double2* q = reinterpret_cast<const double2*>(blk + (size_t)(ok[k] ? b : 0) * 8);
#pragma unroll
            for (int j = 0; j < 4; ++j) v[k][j] = __ldcg(q + j);       // written by other CTAs: bypass L1
        }
#pragma unroll
        for (int k = 0; k < kRound; ++k) {
            if (!ok[k]) continue;
            r[0] += v[k][0].x; r[1] += v[k][0].y; r[2] += v[k][1].x; r[3] += v[k][1].y; r[4] += v[k][2].x;
            r[7] += v[k][3].y;
            const double m5 = v[k][2].y, m6 = v[k][3].x;
            if (m5 > r[5] || (m5 == r[5] && m6 < r[6])) { r[5] = m5; r[6] = m6; }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) shr[threadIdx.x][k] = r[k];
    __syncthreads();
    for (int s = (int)(blockDim.x >> 1); s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            double* a = shr[threadIdx.x];
            const double* b = shr[threadIdx.x + s];
#pragma unroll
            for (int k = 0; k < 5; ++k) a[k] += b[k];
            a[7] += b[7];
            if (b[5] > a[5] || (b[5] == a[5] && b[6] < a[6])) { a[5] = b[5]; a[6] = b[6]; }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = shr[0][k];
}

// Estimate from per-rank partials (rank order = ascending grid offset, so "first maximum" = lowest global index, like
// thrust::max_element / np.argmax).  partial[0..7] as block partials, [8..11] = ECEF / clock of the rank's arg-max
// candidate.  result layout mirrors dpe_result (doubles; indices exact below 2^53).  One thread.
// (not inlined: it runs once per launch, in one thread of the last CTA -- inlined it cost k_score_pairs 14 registers and two of its eight CTAs per SM)
static __device__ __noinline__ void finalize_estimate(const double* __restrict__ parts, int nranks, int est_mode,
                                                  double* __restrict__ zval, double* __restrict__ rval,
                                                  double* __restrict__ res) {
    double sum[5] = {0, 0, 0, 0, 0}, oow = 0, mx = -1.0, mi = 9.0e18;
    int best = -1;
    for (int r = 0; r < nranks; ++r) {
        const double* q = parts + (size_t)r * kPartialLen;
        for (int k = 0; k < 5; ++k) sum[k] += q[k];
        oow += q[7];
        if (q[5] > mx || (q[5] == mx && q[6] < mi)) { mx = q[5]; mi = q[6]; best = r; }
    }
    double z[4] = {0, 0, 0, 0};
    if (est_mode == DPE_EST_WEIGHTED) {                       // BCM_ReduceAndPosMeas :1497-1500
        for (int k = 0; k < 4; ++k) z[k] = sum[k] / sum[4];
    } else if (best >= 0) {                                   // BCM_MakePosMeas
        for (int k = 0; k < 4; ++k) z[k] = parts[(size_t)best * kPartialLen + 8 + k];
    }
    for (int k = 0; k < 4; ++k) zval[k] = z[k];
    for (int r = 0; r < 4; ++r)                               // RVal rows 0-3 <- identity (:2008-2014)
        for (int k = 0; k < 8; ++k) rval[r * 8 + k] = (r == k) ? 1.0 : 0.0;
    res[0] = z[0]; res[1] = z[1]; res[2] = z[2]; res[3] = z[3];
    res[8] = mx; res[9] = sum[4]; res[10] = mi; res[11] = oow;
}

// The single-rank estimate folded into the tail of a scoring kernel (est_mode < 0: not folded, k_finalize follows).
struct FoldEst { int est_mode; double* zval; double* rval; double* res; };

// per-rank partial of the position manifold: sums, max / arg-max, and the arg-max candidate's state
// (BCM_MakePosMeas, batchcorrmanifold.cu:1990-2000)
template <int kRound = 4>
__device__ __forceinline__ void finish_position_partial(const double* __restrict__ blk, int n_blk,
                                                        const double* __restrict__ grid, const EpochDev& e,
                                                        int64_t grid_offset, double* __restrict__ partial,
                                                        const FoldEst& fold) {
    double r[8];
    __shared__ double sp[kPartialLen];                    // the partial, also handed to the folded estimate from here (not read back from global)
    reduce_all_partials<kRound>(blk, n_blk, r);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) sp[k] = r[k];
        for (int k = 8; k < kPartialLen; ++k) sp[k] = 0.0;
        if (r[5] >= 0.0) {
            const Cand p = cand_ecef(e, grid + 4 * ((int64_t)r[6] - grid_offset));
            sp[8] = p.px; sp[9] = p.py; sp[10] = p.pz; sp[11] = p.pt;
        }
        for (int k = 0; k < kPartialLen; ++k) partial[k] = sp[k];
        if (fold.est_mode >= 0) finalize_estimate(sp, 1, fold.est_mode, fold.zval, fold.rval, fold.res);
    }
}

}  // namespace dpe
