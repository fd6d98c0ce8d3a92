// dpe_score.cu -- candidate scoring on the position-clock grid (sm_100a):
//   geometry -> correlogram bin (FP64, same expression order as the reference),
//   lookup + lerp + sum_prn |.|^L, fused block-level arg-max / weighted sums,
//   and the final estimate.
//
// Replaces (cudarecv/modules/src/batchcorrmanifold.cu):
//   BCM_PosMeasML :1710-1828, thrust::max_element :2589, BCM_MakePosMeas :1977-2016,
//   BCM_PosMeasReduction :816-1056, BCM_ReduceAndPosMeas :1365-1510.
// The reference runs <<<8,64>>> grid-stride + a separate device-wide max + a
// 1-block kernel with host syncs in between; here scoring, arg-max and weighted
// accumulation are one pass with a deterministic two-level reduction.
#include "dpe_geom.cuh"

namespace dpe {

// ---------------------------------------------------------------------------
// k_score_lookup: the reference's BCM_PosMeasML with the arg-max / weighted accumulation fused in.
// sat_mode: middle time-grid state (:1773-1775) or per-time-index state (:865-873).
// Every thread scores kLkCand candidates (coalesced: j = base + tid + 128 k): the per-PRN constants are read
// from shared memory once per 4 pairs and the four FP64 chains are independent (round 1, one candidate per
// thread: 169 issued instructions per pair, 13 of them LDS, "wait" + "long scoreboard" 5.5 stalled warps per
// issue -- a latency-bound kernel at 4 % of its HBM roofline).
// WITH_SUMS = 0 (arg-max estimate): only the sum of scores is reduced, not sum s*x.
// ---------------------------------------------------------------------------
// kLkCand candidates per thread: 3 (<= 64 registers, 8 CTAs per SM), 4 or 6 (<= 128 registers, 4 CTAs per SM).  The
// launcher picks the one whose CTA count fills whole waves best: at the demo size (390 625 candidates) 4 per thread is
// 763 CTAs on 592 slots -- 1.29 waves, the second one 29 % full -- while 6 per thread is 509 CTAs: one wave.
// (Tried: 64-thread CTAs, 8 per SM, so that the 0.86 wave spreads 6-7 CTAs instead of 3-4 over every SM: 41.6 us against
// 37.5 -- the finer spread does not pay for twice the prologues and block partials.  Tried: 6 candidates per thread at 96 /
// 80 registers, 5 / 6 CTAs per SM: 37.3 / 39.4 us against 33.5 -- the spills cost more than the extra warps hide.  Tried:
// 7 per thread at 168 registers, 3 CTAs per SM -- 437 CTAs on 444 slots, every SM holds 3, where 6 per thread leaves 65 SMs
// with 4 CTAs and 83 with 3 and the kernel ends with those that hold 4 (phase stamps, profiles/r02ae_phase_stamps.txt):
// 34.7 us against 33.8, and 58.9 against 54.3 us for the velocity branch with the same change in k_score_vel -- 12 warps per SM
// hide less of the FP64 / conversion latency than the even load gives back.)
template <int SAT_MODE, int WITH_SUMS, int kLkCand, bool LP1>
__global__ void __launch_bounds__(kReduceBlock, (kLkCand == 3) ? 8 : 4)
k_score_lookup(const double* __restrict__ grid, const EpochDev* __restrict__ ep, const double* __restrict__ sat,
               const double2* __restrict__ cs, double fs, int S, int W, int NL, int T, int lpower,
               int64_t G, int64_t grid_offset, double* __restrict__ scores, double* __restrict__ blk_partial,
               unsigned int* __restrict__ ticket, double* __restrict__ partial, const SatGeo* __restrict__ geo_tab,
               const FoldEst fold) {
    __shared__ EpochDev e;
    __shared__ ChanConst cc[DPE_MAX_CHAN];
    __shared__ SatGeo geo_mid[DPE_MAX_CHAN];
    DPE_PT_DECL;
    DPE_PT_MARK();                                    // [0] start
    for (int i = threadIdx.x; i < (int)(sizeof(EpochDev) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t*>(&e)[i] = reinterpret_cast<const uint32_t*>(ep)[i];
    __syncthreads();
    chan_consts(e, fs, cc);
    if (SAT_MODE == DPE_SAT_MIDDLE)
        for (int c = threadIdx.x; c < e.C; c += blockDim.x) geo_mid[c] = make_sat_geo(e, sat + ((size_t)c * T + T / 2) * 8);
    __syncthreads();
    DPE_PT_MARK();                                    // [1] per-channel constants in shared memory
    const int64_t base = (int64_t)blockIdx.x * (kReduceBlock * kLkCand) + threadIdx.x;
    CandRel rel[kLkCand];
    double pt[kLkCand], score[kLkCand];
    int it[kLkCand];
    bool act[kLkCand];
#pragma unroll
    for (int k = 0; k < kLkCand; ++k) {
        const int64_t j = base + (int64_t)k * kReduceBlock;
        act[k] = j < G;
        score[k] = 0.0;
        it[k] = T / 2;
        pt[k] = 0.0;
        rel[k].dx = rel[k].dy = rel[k].dz = rel[k].d2 = 0.0;
        if (act[k]) {
            const Cand p = cand_ecef(e, grid + 4 * j, &rel[k]);
            pt[k] = p.pt;
            if (SAT_MODE == DPE_SAT_PER_TIME) it[k] = (int)((j + grid_offset) % T);
        }
    }
    int oow = 0;
    const double rx_time = e.rx_time, Sd = (double)S;
    DPE_PT_MARK();                                    // [2] candidates read and turned
    grid_dep_wait();                                  // the correlogram (cs) and, per time index, the geometry table come from the kernels before
    for (int c = 0; c < e.C; ++c) {
        const ChanConst kc = cc[c];
        const double rc_end = e.rc_end[c];
        SatGeo sg;
        if (SAT_MODE == DPE_SAT_MIDDLE) sg = geo_mid[c];
        const double2* __restrict__ csc = cs + (size_t)c * NL;
        // straight-line code over the kLkCand candidates (no branch inside: the four FP64 chains interleave)
        double idx[kLkCand];
        bool fast = true;
#pragma unroll
        for (int k = 0; k < kLkCand; ++k) {
            if (SAT_MODE == DPE_SAT_PER_TIME) sg = geo_tab[(size_t)c * T + it[k]];
            fast &= code_index_fast_core(rx_time, kc, sg, pt[k], rel[k], rc_end, Sd, &idx[k]);
        }
        if (__builtin_expect(!fast, 0)) {
            // a pair of this thread sits next to a rounding boundary of tx (probability ~1e-5): redo the thread's pairs
            // of this channel through the reference's chain itself, on the same candidate states
#pragma unroll
            for (int k = 0; k < kLkCand; ++k) {
                const Cand p = {rel[k].dx + e.center[0], rel[k].dy + e.center[1], rel[k].dz + e.center[2], pt[k]};
                idx[k] = code_index(e, kc, p, sat + ((size_t)c * T + it[k]) * 8, c, Sd);
            }
        }
#pragma unroll
        for (int k = 0; k < kLkCand; ++k) {
            const BinFast b = make_bin_fast(idx[k], c, S, W);
            const int l = min(max(b.l, 0), NL - 2);              // clamped: the loads are unconditional
            const double2 lo = csc[l], hi = csc[l + 1];
            const double re = hi.x * b.wg + lo.x * b.wf;         // :1808-1812
            const double im = hi.y * b.wg + lo.y * b.wf;
            const double m = mag_pow_t<LP1>(re, im, lpower);     // :1816
            const bool use = b.ok && act[k];
            score[k] += use ? m : 0.0;
            oow += (act[k] && !b.ok) ? 1 : 0;
        }
    }
    DPE_PT_MARK();                                    // [3] all channels scored (warp 0)
    // this thread's candidates in increasing index order: strict > keeps the lowest index on ties
    double v[5] = {0, 0, 0, 0, 0}, mx = -1.0, mi = 9.0e18;
#pragma unroll
    for (int k = 0; k < kLkCand; ++k) {
        if (!act[k]) continue;
        const int64_t j = base + (int64_t)k * kReduceBlock;
        scores[j] = score[k];
        if (WITH_SUMS) {
            v[0] += score[k] * (rel[k].dx + e.center[0]);
            v[1] += score[k] * (rel[k].dy + e.center[1]);
            v[2] += score[k] * (rel[k].dz + e.center[2]);
            v[3] += score[k] * pt[k];
        }
        v[4] += score[k];
        if (score[k] > mx) { mx = score[k]; mi = (double)(j + grid_offset); }
    }
    block_reduce_store_vals<WITH_SUMS ? 5 : 1>(v, mx, mi, (double)oow, blk_partial);
    DPE_PT_MARK();                                    // [4] block partial stored
    const bool last = take_last_ticket(ticket);
    DPE_PT_MARK();                                    // [5] ticket taken
    if (last) finish_position_partial(blk_partial, gridDim.x, grid, e, grid_offset, partial, fold);
    DPE_PT_MARK();                                    // [6] (last CTA) partials reduced, estimate written
    DPE_PT_PRINT(last ? "look-last" : "look", last || (blockIdx.x % 60) == 0);
}

__global__ void __launch_bounds__(128) k_sat_geo(const EpochDev* __restrict__ ep, const double* __restrict__ sat, int T,
                                                 SatGeo* __restrict__ geo) {
    const EpochDev& e = *ep;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < e.C * T; i += gridDim.x * blockDim.x)
        geo[i] = make_sat_geo(e, sat + (size_t)i * 8);
}

// Bins only (parity tests: "code-phase bins bit-exact").  EXACT = 1: the reference's FP64 chain (code_index);
// EXACT = 0: the centre-relative fast path the scoring kernels use (code_index_fast) -- the two must agree bit for bit.
template <int SAT_MODE, int EXACT>
__global__ void __launch_bounds__(256) k_debug_bins(const double* __restrict__ grid, const EpochDev* __restrict__ ep,
                             const double* __restrict__ sat, double fs, int S, int W, int T,
                             int64_t i0, int64_t n, int64_t grid_offset, int64_t* __restrict__ f_idx,
                             double* __restrict__ alpha) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const EpochDev& e = *ep;
    const int64_t j = i0 + t;
    CandRel rel;
    const Cand p = cand_ecef(e, grid + 4 * j, &rel);
    const int it = (SAT_MODE == DPE_SAT_PER_TIME) ? (int)((j + grid_offset) % T) : T / 2;
    for (int c = 0; c < e.C; ++c) {
        const ChanConst k = {fs / e.fc[c], (double)e.cp_ref_tow[c], (e.cp_end[c] - e.cp_ref[c]) * K_T_CA};
        const double* s8 = sat + ((size_t)c * T + it) * 8;
        const double idx = EXACT ? code_index(e, k, p, s8, c, (double)S)
                                 : code_index_fast(e, k, make_sat_geo(e, s8), p, rel, s8, c, (double)S);
        const Bin b = make_bin(idx, c, S, W);
        f_idx[t * e.C + c] = b.f;
        alpha[t * e.C + c] = b.wg;
    }
}

// ---------------------------------------------------------------------------
// Estimate.  partial[0..7] as block partials, [8..11] = ECEF / clock of this rank's arg-max candidate
// (written by the last CTA of the scoring kernel, dpe_geom.cuh).
// ---------------------------------------------------------------------------
// One thread: combine the per-rank partials (rank order = ascending grid offset,
// so "first maximum" = lowest global index, like thrust::max_element / np.argmax).
// result layout mirrors dpe_result (doubles; indices exact below 2^53).
__global__ void k_finalize(const double* __restrict__ parts, int nranks, int est_mode,
                           double* __restrict__ zval, double* __restrict__ rval, double* __restrict__ res) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    finalize_estimate(parts, nranks, est_mode, zval, rval, res);
}

// ---------------------------------------------------------------------------
int launch_sat_geo(dpe_ctx* c, cudaStream_t s) {
    k_sat_geo<<<(c->epoch_C * c->T + 127) / 128, 128, 0, s>>>(c->ep, c->sat, c->T, reinterpret_cast<SatGeo*>(c->sat_geo));
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    return DPE_OK;
}

template <int SAT_MODE, int WITH_SUMS, int NC>
static void launch_lk(dpe_ctx* c, int nblk, const SatGeo* tab, const FoldEst& fold, cudaStream_t s) {
#define DPE_LK_ARGS c->grid, c->ep, c->sat, c->cs, c->cfg.fs, (int)c->S, c->W, c->NL, c->T, c->cfg.lpower, c->G, \
                    c->cfg.grid_offset, c->scores, c->blk_partial, c->ticket, c->partial, tab, fold
    const bool pdl = c->use_pdl != 0;
    if (c->cfg.lpower == 1) launch_dep(k_score_lookup<SAT_MODE, WITH_SUMS, NC, true>, nblk, kReduceBlock, 0, s, pdl, DPE_LK_ARGS);
    else launch_dep(k_score_lookup<SAT_MODE, WITH_SUMS, NC, false>, nblk, kReduceBlock, 0, s, pdl, DPE_LK_ARGS);
#undef DPE_LK_ARGS
}

int launch_score_lookup(dpe_ctx* c, int sat_mode, cudaStream_t s) {
    // candidates per thread: the choice that wastes the least of its last wave (ties: more per thread)
    static const int kNc[3] = {3, 4, 6}, kOcc[3] = {8, 4, 4};
    int nc = 4;
    if (c->lk_cand_forced) nc = c->lk_cand_forced;
    else {
        double best = -1.0;
        for (int i = 0; i < 3; ++i) {
            const double ctas = (double)((c->G + kReduceBlock * kNc[i] - 1) / (kReduceBlock * kNc[i]));
            const double waves = ctas / ((double)c->sm_count * kOcc[i]);
            const double eff = waves / ceil(waves) * (kNc[i] == 3 ? 0.9 : 1.0);      // 3 per thread: less ILP per warp
            if (eff >= best - 1e-9) { best = eff; nc = kNc[i]; }
        }
    }
    const int per = kReduceBlock * nc;
    const int nblk = (int)((c->G + per - 1) / per);
    prof_begin(c, DPE_STAGE_LOOKUP, s);
    const FoldEst fold = {c->fold_est_mode, c->zval, c->rval, c->result};
    const SatGeo* tab = nullptr;
    if (sat_mode == DPE_SAT_PER_TIME) {
        int rc = launch_sat_geo(c, s);
        if (rc) return rc;
        tab = reinterpret_cast<const SatGeo*>(c->sat_geo);
    }
#define DPE_LK_NC(SM, WS) do { if (nc == 3) launch_lk<SM, WS, 3>(c, nblk, tab, fold, s); else if (nc == 6) launch_lk<SM, WS, 6>(c, nblk, tab, fold, s); \
                               else launch_lk<SM, WS, 4>(c, nblk, tab, fold, s); } while (0)
    if (sat_mode == DPE_SAT_PER_TIME) { if (c->want_sums) DPE_LK_NC(DPE_SAT_PER_TIME, 1); else DPE_LK_NC(DPE_SAT_PER_TIME, 0); }
    else { if (c->want_sums) DPE_LK_NC(DPE_SAT_MIDDLE, 1); else DPE_LK_NC(DPE_SAT_MIDDLE, 0); }
#undef DPE_LK_NC
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    c->n_blk_partial = nblk;
    prof_end(c, s);
    return DPE_OK;
}

int launch_estimate(dpe_ctx* c, int est_mode, const double* gathered, int nranks, cudaStream_t s) {
    const double* parts = gathered ? gathered : c->partial;
    if (!gathered) nranks = 1;
    prof_begin(c, DPE_STAGE_ESTIMATE, s);
    k_finalize<<<1, 32, 0, s>>>(parts, nranks, est_mode, c->zval, c->rval, c->result);
    prof_end(c, s);
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    return DPE_OK;
}

int launch_debug_bins(dpe_ctx* c, int64_t i0, int64_t n, int sat_mode, cudaStream_t s) {
    const int nb = (int)((n + 255) / 256);
    const bool per_time = (sat_mode & 1) == DPE_SAT_PER_TIME, exact = (sat_mode & DPE_DEBUG_BINS_EXACT) != 0;
#define DPE_DBG(SM, EX) k_debug_bins<SM, EX><<<nb, 256, 0, s>>>(c->grid, c->ep, c->sat, c->cfg.fs, (int)c->S, c->W, \
                                                                 c->T, i0, n, c->cfg.grid_offset, c->dbg_f, c->dbg_alpha)
    if (per_time && exact) DPE_DBG(DPE_SAT_PER_TIME, 1);
    else if (per_time) DPE_DBG(DPE_SAT_PER_TIME, 0);
    else if (exact) DPE_DBG(DPE_SAT_MIDDLE, 1);
    else DPE_DBG(DPE_SAT_MIDDLE, 0);
#undef DPE_DBG
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    return DPE_OK;
}

int kernel_attr_score(const char* name, cudaFuncAttributes* a) {
    DPE_KATTR("k_score_lookup", (k_score_lookup<DPE_SAT_MIDDLE, 0, 4, true>));
    DPE_KATTR("k_score_lookup_3", (k_score_lookup<DPE_SAT_MIDDLE, 0, 3, true>));
    DPE_KATTR("k_score_lookup_6", (k_score_lookup<DPE_SAT_MIDDLE, 0, 6, true>));
    DPE_KATTR("k_finalize", k_finalize);
    return 0;
}

}  // namespace dpe
