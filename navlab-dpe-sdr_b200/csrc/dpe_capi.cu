// dpe_capi.cu -- the extern "C" boundary of libdpe_b200.so (include/dpe_b200.h).
// Context lifetime, parameter upload, stage sequencing.  No torch types, no
// exceptions across the boundary, no CPU fallback: every stage launches CUDA
// kernels or returns a negative error code.
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <new>
#include "dpe_internal.cuh"

namespace dpe {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace dpe

using namespace dpe;

namespace dpe {
void prof_begin(dpe_ctx* c, int stage, cudaStream_t s) {
    if (!c->prof_on || c->prof_n >= kProfMax) return;
    c->prof_stage[c->prof_n] = stage;
    cudaEventRecord(c->prof_ev[2 * c->prof_n], s);
}
void prof_end(dpe_ctx* c, cudaStream_t s) {
    if (!c->prof_on || c->prof_n >= kProfMax) return;
    cudaEventRecord(c->prof_ev[2 * c->prof_n + 1], s);
    c->prof_n++;
}
}  // namespace dpe

#define DPE_REQUIRE(cond, code, ...)            \
    do {                                        \
        if (!(cond)) {                          \
            dpe::set_error(__VA_ARGS__);        \
            return (code);                      \
        }                                       \
    } while (0)

template <typename Tp>
static int dev_alloc(Tp** p, size_t n, bool zero = true) {
    *p = nullptr;
    if (n == 0) n = 1;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(Tp));
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) -> %s", n * sizeof(Tp), cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? DPE_ENOMEM : DPE_ECUDA;
    }
    if (zero) {
        e = cudaMemset(*p, 0, n * sizeof(Tp));
        if (e != cudaSuccess) { set_error("cudaMemset -> %s", cudaGetErrorString(e)); return DPE_ECUDA; }
    }
    return DPE_OK;
}
#define DPE_ALLOC(ptr, n)                                 \
    do {                                                  \
        int rc__ = dev_alloc(&(ptr), (size_t)(n));        \
        if (rc__) { dpe_ctx_destroy(c); return rc__; }    \
    } while (0)

extern "C" {

const char* dpe_last_error(void) { return g_err; }
int dpe_abi_version(void) { return DPE_ABI_VERSION; }

// failure after `new dpe_ctx`: free what exists, then report
#define DPE_CREATE_REQUIRE(cond, code, ...)     \
    do {                                        \
        if (!(cond)) {                          \
            dpe::set_error(__VA_ARGS__);        \
            dpe_ctx_destroy(c);                 \
            return (code);                      \
        }                                       \
    } while (0)
#define DPE_CREATE_CUDA(call)                                                            \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            dpe::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                 \
                           cudaGetErrorString(e__));                                     \
            dpe_ctx_destroy(c);                                                          \
            return DPE_ECUDA;                                                            \
        }                                                                                \
    } while (0)

static size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int dpe_ctx_create(dpe_ctx** out, const dpe_cfg* cfg) {
    DPE_REQUIRE(out && cfg, DPE_EINVAL, "dpe_ctx_create: null argument");
    *out = nullptr;
    // ---- pure parameter checks: nothing is allocated before all of them pass ----
    DPE_REQUIRE(cfg->abi_version == DPE_ABI_VERSION, DPE_EINVAL, "ABI version %u != %d", cfg->abi_version,
                DPE_ABI_VERSION);
    DPE_REQUIRE(cfg->S >= 64 && (cfg->S % 2) == 0 && cfg->S <= (1 << 26), DPE_EINVAL,
                "S=%lld must be even, 64..2^26", (long long)cfg->S);
    DPE_REQUIRE(cfg->fs > 0, DPE_EINVAL, "fs must be positive");
    DPE_REQUIRE(cfg->max_chan >= 1 && cfg->max_chan <= DPE_MAX_CHAN, DPE_EINVAL, "max_chan out of range");
    DPE_REQUIRE(cfg->G >= 1 && cfg->G < (1ll << 31), DPE_EINVAL, "G out of range");
    DPE_REQUIRE(cfg->lpower >= 1, DPE_EINVAL, "lpower must be >= 1");
    DPE_REQUIRE(cfg->lag_halfwidth >= 0 && cfg->lag_halfwidth <= 160, DPE_EINVAL, "lag_halfwidth 0..160");
    DPE_REQUIRE(cfg->grid_offset >= 0 && cfg->G_total >= cfg->grid_offset + cfg->G, DPE_EINVAL,
                "grid shard [%lld,+%lld) outside G_total=%lld", (long long)cfg->grid_offset, (long long)cfg->G,
                (long long)cfg->G_total);
    const int W = cfg->lag_halfwidth > 0 ? cfg->lag_halfwidth : 32;
    DPE_REQUIRE(2 * W + 2 < cfg->S, DPE_EINVAL, "lag window wider than the block");
    const bool brute = (cfg->flags & DPE_FLAG_BRUTE_TILES) != 0;
    DPE_REQUIRE(!brute || sizeof(int32_t) * ((size_t)cfg->max_chan * (2 * W + 1) + 1 + 4 * (2 * W + 1)) <= 48 * 1024,
                DPE_EINVAL, "max_chan * (2W+1) too large for the pair sort");
    int64_t nf = 0;
    int Wd = 0;
    if (cfg->Gv > 0) {
        DPE_REQUIRE(cfg->Gv < (1ll << 31), DPE_EINVAL, "Gv out of range");
        nf = cfg->n_fft;
        if (nf <= 0) { nf = 1; while (nf < cfg->S) nf <<= 1; nf *= 8; }    // carrSTot, batchcorrscores.cu:761
        // the Doppler-bin twiddle index n*m mod N_c is turned into an angle in FP32: exact up to 2^24
        DPE_REQUIRE((nf & (nf - 1)) == 0 && nf >= cfg->S && nf <= (1 << 24), DPE_EINVAL,
                    "n_fft=%lld must be a power of two, S <= n_fft <= 2^24 (with a velocity grid the default "
                    "8*2^ceil(log2 S) limits S to 2^21)", (long long)nf);
        Wd = cfg->dopp_halfwidth > 0 ? cfg->dopp_halfwidth : 64;
        DPE_REQUIRE(2 * Wd + 2 < nf, DPE_EINVAL, "Doppler window wider than the spectrum");
        DPE_REQUIRE(!(cfg->flags & DPE_FLAG_BRUTE_VEL) ||
                    sizeof(int32_t) * ((size_t)cfg->max_chan * (2 * Wd + 1) + 1 + 4 * (2 * Wd + 1)) <= 48 * 1024, DPE_EINVAL,
                    "max_chan * (2 Wd + 1) too large for the velocity pair sort");
    }
    DevGuard guard(cfg->device >= 0 ? cfg->device : 0);
    int ndev = 0;
    DPE_CUDA(cudaGetDeviceCount(&ndev));
    DPE_REQUIRE(cfg->device >= 0 && cfg->device < ndev, DPE_EINVAL, "device %d of %d", cfg->device, ndev);
    DPE_CUDA(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    DPE_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    DPE_REQUIRE(prop.major == 10, DPE_EINVAL, "libdpe_b200 is built for sm_100a only; device is sm_%d%d",
                prop.major, prop.minor);

    dpe_ctx* c = new (std::nothrow) dpe_ctx();
    DPE_REQUIRE(c, DPE_ENOMEM, "out of host memory");
    memset(c, 0, sizeof(*c));
    c->cfg = *cfg;
    c->sm_count = prop.multiProcessorCount;
    c->S = cfg->S;
    c->S_pad = ((cfg->S + kBfTile - 1) / kBfTile) * kBfTile;
    c->G = cfg->G;
    c->Gv = cfg->Gv;
    c->W = W;
    c->NL = 2 * c->W + 2;
    c->NLp = ((c->NL + kLagTile - 1) / kLagTile) * kLagTile;
    c->H = ((c->W + 8 + 31) / 32) * 32;
    c->nchunk = (int)((c->S + kCorrChunk - 1) / kCorrChunk);
    c->vnchunk = (int)((c->S + kCarrChunk - 1) / kCarrChunk);
    c->maxC = cfg->max_chan;
    c->T = cfg->time_dim > 0 ? cfg->time_dim : 1;
    c->nranks = 1;
    c->want_sums = 1;
    c->fold_est_mode = -1;
    c->stage_fold_est = -1;
    const size_t C = c->maxC, S = c->S, G = c->G;

    // epoch packet {iq | EpochDev | sat}: device copy + page-locked staging copy
    c->sat_cap = C * c->T * 8;
    c->pkt_off_ep = round_up(sizeof(int16_t) * 2 * S + 64, 256);
    c->pkt_off_sat = c->pkt_off_ep + round_up(sizeof(EpochDev), 256);
    c->pkt_bytes = c->pkt_off_sat + round_up(sizeof(double) * c->sat_cap, 256);
    DPE_ALLOC(c->pkt, c->pkt_bytes);
    c->iq_own = reinterpret_cast<int16_t*>(c->pkt);
    c->ep = reinterpret_cast<EpochDev*>(c->pkt + c->pkt_off_ep);
    c->sat = reinterpret_cast<double*>(c->pkt + c->pkt_off_sat);
    DPE_ALLOC(c->ca, DPE_MAX_CHAN * 1024);
    DPE_ALLOC(c->sat_geo, C * c->T * 7);
    DPE_ALLOC(c->tidx, S);
    DPE_ALLOC(c->chan_ticket, DPE_MAX_CHAN);
    DPE_ALLOC(c->xw, C * S);
    DPE_ALLOC(c->rs, C * S);
    if (cfg->flags & DPE_FLAG_KEEP_CHIP_IDX) DPE_ALLOC(c->chip_idx, C * S);
    DPE_ALLOC(c->idx_next, C);
    DPE_ALLOC(c->no_flip, C);
    DPE_ALLOC(c->cacc, C * c->NLp * 4);
    DPE_ALLOC(c->cs, C * c->NL);
    DPE_ALLOC(c->grid, G * 4);
    DPE_ALLOC(c->scores, G);
    DPE_ALLOC(c->blk_partial, ((G + kReduceBlock - 1) / kReduceBlock) * 8);
    DPE_ALLOC(c->partial, kPartialLen);
    DPE_ALLOC(c->ticket, 4);
    DPE_ALLOC(c->zval, 16);     // the reference's EKF_PassMeas reads 16; keep the slack
    DPE_ALLOC(c->rval, 64);
    DPE_ALLOC(c->result, 16);
    if (brute) {
        c->bx_stride = skewX(c->S_pad + 2 * c->H);
        c->brd_stride = skewX(c->S_pad);
        DPE_ALLOC(c->bx, C * 8 * c->bx_stride);
        DPE_ALLOC(c->brd, C * c->brd_stride);
        const size_t NB = 2 * c->W + 1;
        c->max_groups = (int64_t)(C * ((G + kBfNC - 1) / kBfNC + NB * kBfWarps));   // every bucket padded to whole slots
        DPE_ALLOC(c->pair_k, C * G);
        DPE_ALLOC(c->pair_a, C * G);
        DPE_ALLOC(c->pair_v, C * G);
        DPE_ALLOC(c->hist, C * NB);
        DPE_ALLOC(c->blk_hist, C * NB * ((G + kSortBlock - 1) / kSortBlock));
        DPE_ALLOC(c->bucket_base, C * NB);
        DPE_ALLOC(c->group_base, C * NB + 1);
        DPE_ALLOC(c->hdr, c->max_groups * 4);
        DPE_ALLOC(c->ent_j, c->max_groups * kBfNC);
        DPE_ALLOC(c->ent_a, c->max_groups * kBfNC);
        DPE_ALLOC(c->n_groups, 1);
        DPE_ALLOC(c->tail_part, (size_t)c->sm_count * 2 * kBfWarps * kBfNC);
        DPE_ALLOC(c->tail_ticket, c->sm_count * kBfWarps);
    }
    if (cfg->Gv > 0) {
        c->n_fft = (int32_t)nf;
        c->Wd = Wd;
        c->NBd = 2 * c->Wd + 2;
        DPE_ALLOC(c->vgrid, (size_t)cfg->Gv * 4);
        DPE_ALLOC(c->vscores, (size_t)cfg->Gv);
        DPE_ALLOC(c->carr, C * c->NBd);
        DPE_ALLOC(c->dc_part, 2 * c->nchunk);
        DPE_ALLOC(c->bb, C * S);
        DPE_ALLOC(c->vacc, C * c->NBd * 2);
        DPE_ALLOC(c->vticket, DPE_MAX_CHAN);
        DPE_ALLOC(c->vblk_partial, ((cfg->Gv + kReduceBlock - 1) / kReduceBlock) * 8);
        if (cfg->flags & DPE_FLAG_BRUTE_VEL) {
            const size_t NBv = 2 * c->Wd + 1, Gv = (size_t)cfg->Gv;
            c->vS_pad = ((c->S + 1023) / 1024) * 1024;
            c->vmax_groups = (int64_t)(C * ((Gv + kBfNC - 1) / kBfNC + NBv * kBfWarps));
            DPE_ALLOC(c->vbb, C * c->vS_pad);
            DPE_ALLOC(c->vpair_k, C * Gv);
            DPE_ALLOC(c->vpair_a, C * Gv);
            DPE_ALLOC(c->vpair_v, C * Gv);
            DPE_ALLOC(c->vhist, C * NBv);
            DPE_ALLOC(c->vblk_hist, C * NBv * ((Gv + kSortBlock - 1) / kSortBlock));
            DPE_ALLOC(c->vbucket_base, C * NBv);
            DPE_ALLOC(c->vgroup_base, C * NBv + 1);
            DPE_ALLOC(c->vhdr, c->vmax_groups * 4);
            DPE_ALLOC(c->vent_j, c->vmax_groups * kBfNC);
            DPE_ALLOC(c->vent_a, c->vmax_groups * kBfNC);
            DPE_ALLOC(c->vn_groups, 1);
        }
    }
    DPE_CREATE_CUDA(cudaMallocHost(reinterpret_cast<void**>(&c->ep_pin), sizeof(EpochDev) * kPinSlots));
    DPE_CREATE_CUDA(cudaMallocHost(reinterpret_cast<void**>(&c->sat_pin), sizeof(double) * c->sat_cap * kPinSlots));
    DPE_CREATE_CUDA(cudaMallocHost(reinterpret_cast<void**>(&c->pkt_pin), c->pkt_bytes));
    DPE_CREATE_CUDA(cudaMallocHost(reinterpret_cast<void**>(&c->res_pin), sizeof(double) * 16));
    for (int i = 0; i < kPinSlots; ++i) DPE_CREATE_CUDA(cudaEventCreateWithFlags(&c->pin_ev[i], cudaEventDisableTiming));
    DPE_CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_epoch, cudaEventDisableTiming));
    DPE_CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_grid, cudaEventDisableTiming));
    DPE_CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_sort, cudaEventDisableTiming));
    DPE_CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming));
    DPE_CREATE_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    DPE_CREATE_CUDA(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
    if (cfg->Gv > 0) {
        DPE_CREATE_CUDA(cudaStreamCreateWithFlags(&c->vel_stream, cudaStreamNonBlocking));
        DPE_CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        DPE_CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_vel, cudaEventDisableTiming));
        const char* vf = getenv("DPE_VEL_FORK");
        c->vel_fork = !(vf && vf[0] == '0');
    }
    { const char* ng = getenv("DPE_NO_GRAPH"); c->use_graph = !(ng && ng[0] == '1'); }
    { const char* sp = getenv("DPE_BRUTE_SKIP_PAD"); c->brute_skip_pad = !(sp && sp[0] == '0'); }
    { const char* cd = getenv("DPE_CARR_DIRECT"); c->carr_direct = (cd && cd[0] == '1'); }
    { const char* lk = getenv("DPE_LK_CAND"); const int v = lk ? atoi(lk) : 0; c->lk_cand_forced = (v == 3 || v == 4 || v == 6) ? v : 0; }
    { const char* pd = getenv("DPE_PDL"); c->use_pdl = !(pd && pd[0] == '0'); }
    c->iq = c->iq_own;
    int rc = launch_gen_ca(c, 0);
    if (!rc) rc = launch_gen_time(c, 0);
    if (rc) { dpe_ctx_destroy(c); return rc; }
    DPE_CREATE_CUDA(cudaDeviceSynchronize());
    *out = c;
    return DPE_OK;
}

int dpe_ctx_destroy(dpe_ctx* c) {
    if (!c) return DPE_OK;
    DevGuard guard(c->cfg.device);
    cudaDeviceSynchronize();
    if (c->comm) dpe_comm_destroy(c);
    void* ptrs[] = {c->pkt, c->ca, c->sat_geo, c->tidx, c->chan_ticket, c->xw, c->rs, c->chip_idx, c->idx_next, c->no_flip,
                    c->cacc, c->cs, c->bx, c->brd, c->grid, c->scores, c->blk_partial, c->partial,
                    c->zval, c->rval, c->result, c->ticket, c->pair_k, c->pair_a, c->pair_v, c->hist, c->blk_hist,
                    c->bucket_base, c->group_base, c->hdr, c->ent_j, c->ent_a, c->n_groups, c->tail_part, c->tail_ticket, c->dbg_f, c->dbg_alpha,
                    c->vgrid, c->vscores, c->carr, c->dc_part, c->bb, c->vacc, c->vticket, c->vblk_partial, c->gathered,
                    c->vbb, c->vpair_k, c->vpair_a, c->vpair_v, c->vhist, c->vblk_hist, c->vbucket_base, c->vgroup_base,
                    c->vhdr, c->vent_j, c->vent_a, c->vn_groups};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (c->ep_pin) cudaFreeHost(c->ep_pin);
    if (c->sat_pin) cudaFreeHost(c->sat_pin);
    if (c->pkt_pin) cudaFreeHost(c->pkt_pin);
    if (c->res_pin) cudaFreeHost(c->res_pin);
    for (int i = 0; i < kPinSlots; ++i)
        if (c->pin_ev[i]) cudaEventDestroy(c->pin_ev[i]);
    if (c->ev_epoch) cudaEventDestroy(c->ev_epoch);
    if (c->ev_grid) cudaEventDestroy(c->ev_grid);
    if (c->ev_sort) cudaEventDestroy(c->ev_sort);
    if (c->ev_done) cudaEventDestroy(c->ev_done);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if (c->vel_stream) cudaStreamDestroy(c->vel_stream);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_vel) cudaEventDestroy(c->ev_vel);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)c->graph_exec);
    if (c->prof_ev) {
        for (int i = 0; i < 2 * kProfMax; ++i)
            if (c->prof_ev[i]) cudaEventDestroy(c->prof_ev[i]);
        delete[] c->prof_ev;
    }
    delete[] c->prof_stage;
    delete c;
    return DPE_OK;
}

int dpe_grid_set(dpe_ctx* c, const double* enu_dt, int64_t G, void* stream) {
    DPE_REQUIRE(c && enu_dt, DPE_EINVAL, "dpe_grid_set: null argument");
    DevGuard guard(c->cfg.device);
    DPE_REQUIRE(G == c->G, DPE_EINVAL, "dpe_grid_set: G=%lld, context holds %lld", (long long)G, (long long)c->G);
    if (c->sort_pending) {                         // a presort on another stream still reads the old grid
        DPE_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, c->ev_sort, 0));
        c->sort_pending = 0;
    }
    c->sort_valid = 0;
    c->have_scores = 0;
    DPE_CUDA(cudaMemcpyAsync(c->grid, enu_dt, sizeof(double) * 4 * G, cudaMemcpyDefault, (cudaStream_t)stream));
    DPE_CUDA(cudaEventRecord(c->ev_grid, (cudaStream_t)stream));    // a presort on another stream waits for the new grid
    if ((cudaStream_t)stream != c->own_stream) DPE_CUDA(cudaStreamWaitEvent(c->own_stream, c->ev_grid, 0));
    return DPE_OK;
}

int dpe_vel_grid_set(dpe_ctx* c, const double* venu_ddt, int64_t Gv, void* stream) {
    DPE_REQUIRE(c && venu_ddt, DPE_EINVAL, "dpe_vel_grid_set: null argument");
    DevGuard guard(c->cfg.device);
    c->fork_valid = 0;
    DPE_REQUIRE(c->Gv > 0, DPE_ESTATE, "context created without a velocity grid (cfg.Gv = 0)");
    DPE_REQUIRE(Gv == c->Gv, DPE_EINVAL, "dpe_vel_grid_set: Gv=%lld, context holds %lld", (long long)Gv,
                (long long)c->Gv);
    DPE_CUDA(cudaMemcpyAsync(c->vgrid, venu_ddt, sizeof(double) * 4 * Gv, cudaMemcpyDefault, (cudaStream_t)stream));
    if ((cudaStream_t)stream != c->own_stream) {
        DPE_CUDA(cudaEventRecord(c->ev_grid, (cudaStream_t)stream));
        DPE_CUDA(cudaStreamWaitEvent(c->own_stream, c->ev_grid, 0));
    }
    c->have_vgrid = 1;
    return DPE_OK;
}

int dpe_block_stage(dpe_ctx* c, const int16_t* iq, int64_t S, void* stream) {
    DPE_REQUIRE(c && iq, DPE_EINVAL, "dpe_block_stage: null argument");
    DevGuard guard(c->cfg.device);
    c->fork_valid = 0;
    DPE_REQUIRE(S == c->S, DPE_EINVAL, "block of %lld samples, context built for %lld", (long long)S,
                (long long)c->S);
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, iq);
    bool on_device = (e == cudaSuccess) && (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged);
    if (e != cudaSuccess) cudaGetLastError();
    if (on_device && at.device == c->cfg.device && (reinterpret_cast<uintptr_t>(iq) & 15) == 0) {
        c->iq = iq;                                   // zero copy
    } else {
        DPE_CUDA(cudaMemcpyAsync(c->iq_own, iq, sizeof(int16_t) * 2 * S, cudaMemcpyDefault,
                                 (cudaStream_t)stream));
        c->iq = c->iq_own;
    }
    c->have_block = 1;
    c->have_prepare = c->have_corr = c->have_scores = 0;
    return DPE_OK;
}

int dpe_epoch_set_part(dpe_ctx* c, const dpe_epoch* ep, const double* sat_states, unsigned parts,
                       void* stream) {
    DPE_REQUIRE(c && ep, DPE_EINVAL, "dpe_epoch_set: null argument");
    DevGuard guard(c->cfg.device);
    c->fork_valid = 0;
    DPE_REQUIRE(parts && !(parts & ~(DPE_PART_CHANNELS | DPE_PART_GEOMETRY)), DPE_EINVAL, "bad parts mask %u", parts);
    DPE_REQUIRE(ep->C >= 1 && ep->C <= c->maxC, DPE_EINVAL, "C=%d, context built for <= %d", ep->C, c->maxC);
    DPE_REQUIRE(!(parts & DPE_PART_GEOMETRY) || sat_states, DPE_EINVAL, "geometry part without sat_states");
    EpochDev& h = c->ep_host;
    if (c->have_epoch && h.C != ep->C) {           // channel count changed: both parts must be set again
        c->have_epoch = 0;
    }
    h.C = ep->C;
    for (int i = 0; i < ep->C; ++i)
        DPE_REQUIRE(ep->fc[i] > 0, DPE_EINVAL, "code frequency of channel %d not positive", i);
    if (parts & DPE_PART_CHANNELS) {
        h.doppler_sign = ep->doppler_sign;
        for (int i = 0; i < ep->C; ++i) {
            DPE_REQUIRE(ep->prn[i] >= 1 && ep->prn[i] <= DPE_MAX_CHAN, DPE_EINVAL, "PRN %d out of range",
                        ep->prn[i]);
            h.prn[i] = ep->prn[i];
            h.rc_start[i] = ep->rc_start[i]; h.ri_start[i] = ep->ri_start[i];
            h.fc[i] = ep->fc[i]; h.fi[i] = ep->fi[i];
            h.cp_start[i] = ep->cp_start[i]; h.cp_ref[i] = ep->cp_ref[i];
        }
    }
    if (parts & DPE_PART_GEOMETRY) {
        h.rx_time = ep->rx_time;
        memcpy(h.center, ep->center, sizeof(h.center));
        memcpy(h.R, ep->enu2ecef, sizeof(h.R));
        for (int i = 0; i < ep->C; ++i) {
            h.fc[i] = ep->fc[i];
            h.cp_ref[i] = ep->cp_ref[i];
            h.rc_end[i] = ep->rc_end[i]; h.cp_end[i] = ep->cp_end[i]; h.cp_ref_tow[i] = ep->cp_ref_tow[i];
        }
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (c->sort_pending) {                         // a presort on another stream still reads the old parameters
        DPE_CUDA(cudaStreamWaitEvent(s, c->ev_sort, 0));
        c->sort_pending = 0;
    }
    c->sort_valid = 0;
    // stage through a page-locked ring slot: truly asynchronous, and the caller may reuse its arrays at once
    const int slot = c->pin_next;
    c->pin_next = (slot + 1) % kPinSlots;
    DPE_CUDA(cudaEventSynchronize(c->pin_ev[slot]));          // the copy issued kPinSlots calls ago has long finished
    c->ep_pin[slot] = h;
    DPE_CUDA(cudaMemcpyAsync(c->ep, &c->ep_pin[slot], sizeof(h), cudaMemcpyHostToDevice, s));
    if (parts & DPE_PART_GEOMETRY) {
        const size_t n = 8 * (size_t)ep->C * c->T;
        cudaPointerAttributes at;
        const bool dev = cudaPointerGetAttributes(&at, sat_states) == cudaSuccess && at.type == cudaMemoryTypeDevice;
        if (!dev) cudaGetLastError();
        if (dev) {
            DPE_CUDA(cudaMemcpyAsync(c->sat, sat_states, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
        } else {
            double* stage = c->sat_pin + (size_t)slot * c->sat_cap;
            memcpy(stage, sat_states, sizeof(double) * n);
            DPE_CUDA(cudaMemcpyAsync(c->sat, stage, sizeof(double) * n, cudaMemcpyHostToDevice, s));
        }
    }
    DPE_CUDA(cudaEventRecord(c->pin_ev[slot], s));
    DPE_CUDA(cudaEventRecord(c->ev_epoch, s));
    c->epoch_C = ep->C;
    c->have_epoch |= (int)parts;
    if (parts & DPE_PART_CHANNELS) c->have_prepare = c->have_corr = 0;
    c->have_scores = 0;
    return DPE_OK;
}

// packs device-resident channel / geometry parameters into the context's EpochDev + satellite states
__global__ void DPE_SIDE128
k_pack_epoch(dpe_epoch_dev p, unsigned parts, int T, EpochDev* __restrict__ e, double* __restrict__ sat) {
    const int C = p.C;
    if (threadIdx.x == 0) e->C = C;
    if (parts & DPE_PART_CHANNELS) {
        if (threadIdx.x == 0) e->doppler_sign = p.doppler_sign ? p.doppler_sign[0] : 1;
        for (int i = threadIdx.x; i < C; i += blockDim.x) {
            e->prn[i] = p.prn[i];
            e->rc_start[i] = p.rc_start[i]; e->ri_start[i] = p.ri_start[i];
            e->fc[i] = p.fc[i]; e->fi[i] = p.fi[i];
            e->cp_start[i] = p.cp_start[i]; e->cp_ref[i] = p.cp_ref[i];
        }
    }
    if (parts & DPE_PART_GEOMETRY) {
        if (threadIdx.x == 0) e->rx_time = p.rx_time;
        if (threadIdx.x < 8) e->center[threadIdx.x] = p.center[threadIdx.x];
        if (threadIdx.x < 9) e->R[threadIdx.x] = p.enu2ecef[threadIdx.x];
        for (int i = threadIdx.x; i < C; i += blockDim.x) {
            e->fc[i] = p.fc[i];
            e->cp_ref[i] = p.cp_ref[i];
            e->rc_end[i] = p.rc_end[i]; e->cp_end[i] = p.cp_end[i]; e->cp_ref_tow[i] = p.cp_ref_tow[i];
        }
        for (int i = threadIdx.x; i < 8 * C * T; i += blockDim.x) sat[i] = p.sat_states[i];
    }
}

int dpe_epoch_set_device(dpe_ctx* c, const dpe_epoch_dev* ep, unsigned parts, void* stream) {
    DPE_REQUIRE(c && ep, DPE_EINVAL, "dpe_epoch_set_device: null argument");
    DevGuard guard(c->cfg.device);
    DPE_REQUIRE(parts && !(parts & ~(DPE_PART_CHANNELS | DPE_PART_GEOMETRY)), DPE_EINVAL, "bad parts mask %u", parts);
    DPE_REQUIRE(ep->C >= 1 && ep->C <= c->maxC, DPE_EINVAL, "C=%d, context built for <= %d", ep->C, c->maxC);
    DPE_REQUIRE(ep->fc && ep->cp_ref, DPE_EINVAL, "null fc / cp_ref");
    DPE_REQUIRE(!(parts & DPE_PART_CHANNELS) || (ep->prn && ep->rc_start && ep->ri_start && ep->fi && ep->cp_start),
                DPE_EINVAL, "channel part: null pointer");
    DPE_REQUIRE(!(parts & DPE_PART_GEOMETRY) || (ep->rc_end && ep->cp_end && ep->cp_ref_tow && ep->center &&
                                                 ep->enu2ecef && ep->sat_states), DPE_EINVAL, "geometry part: null pointer");
    if (c->have_epoch && c->epoch_C != ep->C) c->have_epoch = 0;      // channel count changed: both parts again
    cudaStream_t s = (cudaStream_t)stream;
    if (c->sort_pending) { DPE_CUDA(cudaStreamWaitEvent(s, c->ev_sort, 0)); c->sort_pending = 0; }
    c->sort_valid = 0;
    k_pack_epoch<<<1, 128, 0, s>>>(*ep, parts, c->T, c->ep, c->sat);
    c->launches++;
    DPE_CUDA(cudaGetLastError());
    DPE_CUDA(cudaEventRecord(c->ev_epoch, s));
    c->epoch_C = ep->C;
    c->ep_host.C = ep->C;
    c->have_epoch |= (int)parts;
    if (parts & DPE_PART_CHANNELS) c->have_prepare = c->have_corr = 0;
    c->have_scores = 0;
    return DPE_OK;
}

int dpe_epoch_set(dpe_ctx* c, const dpe_epoch* ep, const double* sat_states, void* stream) {
    DPE_REQUIRE(sat_states, DPE_EINVAL, "dpe_epoch_set: null argument");
    return dpe_epoch_set_part(c, ep, sat_states, DPE_PART_CHANNELS | DPE_PART_GEOMETRY, stream);
}

int dpe_replica_prepare(dpe_ctx* c, void* stream) {
    DPE_REQUIRE(c, DPE_EINVAL, "null context");
    DevGuard guard(c->cfg.device);
    c->fork_valid = 0;
    DPE_REQUIRE(c->have_block && (c->have_epoch & DPE_PART_CHANNELS), DPE_ESTATE,
                "replica_prepare before block_stage / the channel part of epoch_set");
    int rc = launch_prepare(c, (cudaStream_t)stream);
    if (rc) return rc;
    c->have_prepare = 1;
    c->have_planes = 0;
    c->have_corr = 0;
    return DPE_OK;
}

int dpe_correlogram(dpe_ctx* c, void* stream) {
    DPE_REQUIRE(c, DPE_EINVAL, "null context");
    DevGuard guard(c->cfg.device);
    c->fork_valid = 0;
    DPE_REQUIRE(c->have_prepare, DPE_ESTATE, "correlogram before replica_prepare");
    int rc = launch_correlogram(c, (cudaStream_t)stream);
    if (rc) return rc;
    c->have_corr = 1;
    c->have_planes = 0;
    return DPE_OK;
}

int dpe_code_scores_set(dpe_ctx* c, const double* cs, int C, void* stream) {
    DPE_REQUIRE(c && cs, DPE_EINVAL, "null argument");
    DevGuard guard(c->cfg.device);
    c->fork_valid = 0;
    DPE_REQUIRE(c->have_epoch, DPE_ESTATE, "code_scores_set before epoch_set");
    DPE_REQUIRE(C == c->epoch_C, DPE_EINVAL, "C=%d, epoch has %d channels", C, c->epoch_C);
    DPE_CUDA(cudaMemcpyAsync(c->cs, cs, sizeof(double2) * (size_t)C * c->NL, cudaMemcpyDefault,
                             (cudaStream_t)stream));
    c->have_corr = 1;
    c->have_prepare = 0;       // the brute-force planes do not belong to this correlogram
    return DPE_OK;
}

int dpe_score_pos(dpe_ctx* c, int score_mode, int sat_mode, void* stream) {
    DPE_REQUIRE(c, DPE_EINVAL, "null context");
    DevGuard guard(c->cfg.device);
    DPE_REQUIRE(c->have_corr, DPE_ESTATE, "score_pos before correlogram");
    DPE_REQUIRE(c->have_epoch & DPE_PART_GEOMETRY, DPE_ESTATE, "score_pos before the geometry part of epoch_set");
    DPE_REQUIRE(sat_mode == DPE_SAT_MIDDLE || sat_mode == DPE_SAT_PER_TIME, DPE_EINVAL, "bad sat_mode");
    int rc;
    if (c->vel_stream && c->vel_fork) {           // everything the velocity manifold reads is in place at this point of the stream
        DPE_CUDA(cudaEventRecord(c->ev_fork, (cudaStream_t)stream));
        c->fork_valid = 1;
    }
    const bool fold_here = c->fold_est_mode < 0 && c->stage_fold_est >= 0;     // stage-by-stage caller asked for it (dpe_fold_estimate)
    if (fold_here) c->fold_est_mode = c->stage_fold_est;
    c->folded_est = c->fold_est_mode >= 0 ? c->fold_est_mode + 1 : 0;
    if (score_mode == DPE_SCORE_LOOKUP) {
        rc = launch_score_lookup(c, sat_mode, (cudaStream_t)stream);
    } else if (score_mode == DPE_SCORE_BRUTE) {
        DPE_REQUIRE(c->cfg.flags & DPE_FLAG_BRUTE_TILES, DPE_ESTATE,
                    "context created without DPE_FLAG_BRUTE_TILES");
        DPE_REQUIRE(c->have_prepare, DPE_ESTATE,
                    "brute-force scoring needs this context's own replica_prepare + correlogram");
        rc = launch_score_brute(c, sat_mode, (cudaStream_t)stream);
    } else {
        set_error("bad score_mode %d", score_mode);
        rc = DPE_EINVAL;
    }
    if (fold_here) c->fold_est_mode = -1;
    if (rc) { c->folded_est = 0; return rc; }
    c->have_scores = 1;
    return DPE_OK;
}

int dpe_brute_presort(dpe_ctx* c, int sat_mode, void* stream) {
    DPE_REQUIRE(c, DPE_EINVAL, "null context");
    DevGuard guard(c->cfg.device);
    DPE_REQUIRE(c->cfg.flags & DPE_FLAG_BRUTE_TILES, DPE_ESTATE, "context created without DPE_FLAG_BRUTE_TILES");
    DPE_REQUIRE((c->have_epoch & (DPE_PART_CHANNELS | DPE_PART_GEOMETRY)) == (DPE_PART_CHANNELS | DPE_PART_GEOMETRY),
                DPE_ESTATE, "brute_presort before both parts of epoch_set");
    DPE_REQUIRE(sat_mode == DPE_SAT_MIDDLE || sat_mode == DPE_SAT_PER_TIME, DPE_EINVAL, "bad sat_mode");
    cudaStream_t s = (cudaStream_t)stream;
    DPE_CUDA(cudaStreamWaitEvent(s, c->ev_epoch, 0));          // the parameters this epoch's upload put in place
    if (!c->capturing) {
        DPE_CUDA(cudaStreamWaitEvent(s, c->ev_grid, 0));       // ... and the grid, should it have been replaced since
        if (c->sort_pending) DPE_CUDA(cudaStreamWaitEvent(s, c->ev_sort, 0));
    }
    int rc = launch_brute_sort(c, sat_mode, s);
    if (rc) return rc;
    DPE_CUDA(cudaEventRecord(c->ev_sort, s));
    c->sort_valid = 1 + sat_mode;
    c->sort_pending = 1;
    return DPE_OK;
}

int dpe_estimate(dpe_ctx* c, int est_mode, const double* gathered, int nranks, void* stream) {
    DPE_REQUIRE(c, DPE_EINVAL, "null context");
    DevGuard guard(c->cfg.device);
    DPE_REQUIRE(c->have_scores, DPE_ESTATE, "estimate before score_pos");
    DPE_REQUIRE(est_mode == DPE_EST_ARGMAX || est_mode == DPE_EST_WEIGHTED, DPE_EINVAL, "bad est_mode");
    DPE_REQUIRE(!gathered || nranks >= 1, DPE_EINVAL, "nranks must be >= 1");
    if (!gathered && c->folded_est == est_mode + 1) return DPE_OK;     // dpe_score_pos already wrote this estimate (dpe_fold_estimate)
    return launch_estimate(c, est_mode, gathered, nranks, (cudaStream_t)stream);
}

int dpe_fold_estimate(dpe_ctx* c, int est_mode) {
    DPE_REQUIRE(c, DPE_EINVAL, "null context");
    DPE_REQUIRE(est_mode == -1 || est_mode == DPE_EST_ARGMAX || est_mode == DPE_EST_WEIGHTED, DPE_EINVAL, "bad est_mode");
    c->stage_fold_est = est_mode;
    c->folded_est = 0;
    return DPE_OK;
}

int dpe_score_vel(dpe_ctx* c, void* stream) { return dpe_score_vel_est(c, DPE_EST_ARGMAX, stream); }

int dpe_score_vel_est(dpe_ctx* c, int est_mode, void* stream) {
    DPE_REQUIRE(c, DPE_EINVAL, "null context");
    DPE_REQUIRE(est_mode == DPE_EST_ARGMAX || est_mode == DPE_EST_WEIGHTED, DPE_EINVAL, "bad est_mode");
    DevGuard guard(c->cfg.device);
    c->vel_weighted = (est_mode == DPE_EST_WEIGHTED);
    DPE_REQUIRE(c->Gv > 0 && c->have_vgrid, DPE_ESTATE, "score_vel without a velocity grid");
    DPE_REQUIRE(c->have_prepare && c->have_corr, DPE_ESTATE, "score_vel before replica_prepare / correlogram");
    DPE_REQUIRE(c->have_epoch & DPE_PART_GEOMETRY, DPE_ESTATE, "score_vel before the geometry part of epoch_set");
    cudaStream_t s = (cudaStream_t)stream;
    if (!c->fork_valid) return launch_score_vel(c, s);
    // beside the position scoring that dpe_score_pos put on `s`: fork behind the point it recorded, join `s` again
    c->fork_valid = 0;
    DPE_CUDA(cudaStreamWaitEvent(c->vel_stream, c->ev_fork, 0));
    int rc = launch_score_vel(c, c->vel_stream);
    DPE_CUDA(cudaEventRecord(c->ev_vel, c->vel_stream));
    DPE_CUDA(cudaStreamWaitEvent(s, c->ev_vel, 0));
    return rc;
}

int dpe_score_vel_brute(dpe_ctx* c, void* stream) {
    DPE_REQUIRE(c, DPE_EINVAL, "null context");
    DevGuard guard(c->cfg.device);
    DPE_REQUIRE(c->Gv > 0 && c->have_vgrid, DPE_ESTATE, "score_vel_brute without a velocity grid");
    DPE_REQUIRE(c->vbb, DPE_ESTATE, "context created without DPE_FLAG_BRUTE_VEL");
    DPE_REQUIRE(c->have_prepare && c->have_corr, DPE_ESTATE, "score_vel_brute before replica_prepare / correlogram");
    DPE_REQUIRE(c->have_epoch & DPE_PART_GEOMETRY, DPE_ESTATE, "score_vel_brute before the geometry part of epoch_set");
    return launch_score_vel_brute(c, (cudaStream_t)stream);
}

int dpe_result_fetch(dpe_ctx* c, dpe_result* out, void* stream) {
    DPE_REQUIRE(c && out, DPE_EINVAL, "null argument");
    DevGuard guard(c->cfg.device);
    double r[16];
    DPE_CUDA(cudaMemcpyAsync(r, c->result, sizeof(r), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    DPE_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    memset(out, 0, sizeof(*out));
    for (int i = 0; i < 8; ++i) out->z[i] = r[i];
    out->max_score = r[8];
    out->sum_score = r[9];
    out->argmax = (int64_t)r[10];
    out->out_of_window = (int64_t)r[11];
    out->vel_max_score = r[12];
    out->vel_argmax = (int64_t)r[13];
    out->vel_out_of_window = (int64_t)r[14];
    return DPE_OK;
}

// ---------------------------------------------------------------------------------------------
// One whole epoch enqueued on stream `s`: packet upload (rank 0) -> [ncclBroadcast] -> pre-pass ->
// correlogram -> scoring of this context's shard -> [ncclAllGather of the partials] -> estimate ->
// [velocity manifold].  No host synchronisation; the pair sort of a brute-force epoch runs on the
// context's second stream beside the sample pre-pass.
// ---------------------------------------------------------------------------------------------
static int upload_epoch(dpe_ctx* c, const int16_t* iq, const dpe_epoch* ep, const double* sat_states,
                        int score_mode, int est_mode, int with_vel, cudaStream_t s, bool own_copy) {
    DPE_REQUIRE(ep, DPE_EINVAL, "epoch: null parameters");
    DPE_REQUIRE(ep->C >= 1 && ep->C <= c->maxC, DPE_EINVAL, "C=%d, context built for <= %d", ep->C, c->maxC);
    DPE_REQUIRE(score_mode == DPE_SCORE_LOOKUP || score_mode == DPE_SCORE_BRUTE, DPE_EINVAL, "bad score_mode %d", score_mode);
    DPE_REQUIRE(est_mode == DPE_EST_ARGMAX || est_mode == DPE_EST_WEIGHTED, DPE_EINVAL, "bad est_mode");
    DPE_REQUIRE(score_mode != DPE_SCORE_BRUTE || (c->cfg.flags & DPE_FLAG_BRUTE_TILES), DPE_ESTATE,
                "context created without DPE_FLAG_BRUTE_TILES");
    DPE_REQUIRE(!with_vel || (c->Gv > 0 && c->have_vgrid), DPE_ESTATE, "velocity manifold requested without a velocity grid");
    DPE_REQUIRE(with_vel >= 0 && with_vel <= 3, DPE_EINVAL,
                "with_vel: 0 none, 1 lookup (arg-max), 2 brute force (arg-max), 3 lookup (score-weighted mean)");
    const bool root = !c->comm || c->rank == 0;
    const int C = ep->C;
    const size_t sat_bytes = sizeof(double) * 8 * (size_t)C * c->T;
    const size_t used = c->pkt_off_sat + sat_bytes;
    if (c->sort_pending) { DPE_CUDA(cudaStreamWaitEvent(s, c->ev_sort, 0)); c->sort_pending = 0; }
    if (root) {
        DPE_REQUIRE(iq && sat_states, DPE_EINVAL, "epoch: null block / satellite states on the root rank");
        for (int i = 0; i < C; ++i) {
            DPE_REQUIRE(ep->prn[i] >= 1 && ep->prn[i] <= DPE_MAX_CHAN, DPE_EINVAL, "PRN %d out of range", ep->prn[i]);
            DPE_REQUIRE(ep->fc[i] > 0, DPE_EINVAL, "code frequency of channel %d not positive", i);
        }
        EpochDev& h = *reinterpret_cast<EpochDev*>(c->pkt_pin + c->pkt_off_ep);
        h.C = C;
        h.doppler_sign = ep->doppler_sign;
        h.rx_time = ep->rx_time;
        memcpy(h.center, ep->center, sizeof(h.center));
        memcpy(h.R, ep->enu2ecef, sizeof(h.R));
        for (int i = 0; i < C; ++i) {
            h.prn[i] = ep->prn[i];
            h.rc_start[i] = ep->rc_start[i]; h.ri_start[i] = ep->ri_start[i];
            h.fc[i] = ep->fc[i]; h.fi[i] = ep->fi[i];
            h.cp_start[i] = ep->cp_start[i]; h.cp_ref[i] = ep->cp_ref[i];
            h.rc_end[i] = ep->rc_end[i]; h.cp_end[i] = ep->cp_end[i]; h.cp_ref_tow[i] = ep->cp_ref_tow[i];
        }
        c->ep_host = h;
        cudaPointerAttributes at;
        bool sat_dev = cudaPointerGetAttributes(&at, sat_states) == cudaSuccess && at.type == cudaMemoryTypeDevice;
        if (!sat_dev) { cudaGetLastError(); memcpy(c->pkt_pin + c->pkt_off_sat, sat_states, sat_bytes); }
        cudaError_t e = cudaPointerGetAttributes(&at, iq);
        const bool iq_dev = (e == cudaSuccess) && (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged);
        if (e != cudaSuccess) cudaGetLastError();
        if (!iq_dev) {
            // host block: one page-locked packet, ONE H2D for block + parameters + satellite states
            memcpy(c->pkt_pin, iq, sizeof(int16_t) * 2 * c->S);
            DPE_CUDA(cudaMemcpyAsync(c->pkt, c->pkt_pin, used, cudaMemcpyHostToDevice, s));
            c->iq = c->iq_own;
        } else {
            DPE_CUDA(cudaMemcpyAsync(c->pkt + c->pkt_off_ep, c->pkt_pin + c->pkt_off_ep, used - c->pkt_off_ep,
                                     cudaMemcpyHostToDevice, s));
            const bool in_place = !own_copy && !c->comm && at.device == c->cfg.device && (reinterpret_cast<uintptr_t>(iq) & 15) == 0;
            if (in_place) {
                c->iq = iq;                                   // zero copy
            } else {
                DPE_CUDA(cudaMemcpyAsync(c->iq_own, iq, sizeof(int16_t) * 2 * c->S, cudaMemcpyDefault, s));
                c->iq = c->iq_own;
            }
        }
        if (sat_dev) DPE_CUDA(cudaMemcpyAsync(c->sat, sat_states, sat_bytes, cudaMemcpyDeviceToDevice, s));
    } else {
        c->iq = c->iq_own;
        c->ep_host.C = C;
    }
    c->pkt_used = used;
    c->epoch_C = C;
    return DPE_OK;
}

// the device work of one epoch after the upload: [broadcast] -> pre-pass + correlogram (|| pair sort on the second
// stream) -> scoring -> [all-gather] -> estimate -> [velocity manifold].  Pure enqueue: this is what gets captured.
static int compute_epoch(dpe_ctx* c, int score_mode, int est_mode, int with_vel, cudaStream_t s) {
    int rc;
    if (c->comm && (rc = comm_broadcast(c, c->pkt, c->pkt_used, s))) return rc;
    DPE_CUDA(cudaEventRecord(c->ev_epoch, s));
    c->have_block = 1;
    c->have_epoch = DPE_PART_CHANNELS | DPE_PART_GEOMETRY;
    c->have_prepare = c->have_corr = c->have_scores = 0;
    c->sort_valid = 0;
    c->fork_valid = 0;
    const int sat_mode = (est_mode == DPE_EST_WEIGHTED) ? DPE_SAT_PER_TIME : DPE_SAT_MIDDLE;
    if (score_mode == DPE_SCORE_BRUTE && (rc = dpe_brute_presort(c, sat_mode, c->aux_stream))) return rc;
    if ((rc = dpe_replica_prepare(c, s))) return rc;
    if ((rc = dpe_correlogram(c, s))) return rc;
    c->want_sums = (est_mode == DPE_EST_WEIGHTED);     // an arg-max epoch needs no per-candidate sum s*x
    c->fold_est_mode = c->comm ? -1 : est_mode;        // one GPU: the estimate is the tail of the scoring kernel
    const int stage_fold = c->stage_fold_est;
    c->stage_fold_est = -1;                            // (a sharded epoch never folds, whatever dpe_fold_estimate set)
    rc = dpe_score_pos(c, score_mode, sat_mode, s);
    c->stage_fold_est = stage_fold;
    c->want_sums = 1;
    c->fold_est_mode = -1;
    if (rc) return rc;
    if (c->comm) {
        if ((rc = comm_allgather(c, c->partial, c->gathered, kPartialLen, s))) return rc;
        if ((rc = launch_estimate(c, est_mode, c->gathered, c->nranks, s))) return rc;
    }
    if (with_vel == 2) { if ((rc = dpe_score_vel_brute(c, s))) return rc; }
    else if (with_vel && (rc = dpe_score_vel_est(c, with_vel == 3 ? DPE_EST_WEIGHTED : DPE_EST_ARGMAX, s))) return rc;
    return DPE_OK;
}

static int enqueue_epoch(dpe_ctx* c, const int16_t* iq, const dpe_epoch* ep, const double* sat_states,
                         int score_mode, int est_mode, int with_vel, cudaStream_t s) {
    int rc = upload_epoch(c, iq, ep, sat_states, score_mode, est_mode, with_vel, s, false);
    if (rc) return rc;
    return compute_epoch(c, score_mode, est_mode, with_vel, s);
}

// dpe_epoch_submit's device work as ONE cudaGraphLaunch: the kernel chain of compute_epoch is captured once per
// (scoring path, estimator, velocity, channel count, communicator) and replayed; every kernel argument of the chain is a
// context-owned buffer (the block is always copied into the packet in this mode), so a replay needs no update.
static int launch_epoch_graph(dpe_ctx* c, int score_mode, int est_mode, int with_vel, cudaStream_t s) {
    // (pkt_used < 2^29 for S <= 2^26: bits 13..41; the velocity mode sits above it)
    const uint64_t key = 1u | (uint64_t)score_mode << 1 | (uint64_t)est_mode << 2 | (uint64_t)(with_vel & 3) << 60 |
                         (uint64_t)c->epoch_C << 4 | (uint64_t)(c->comm != nullptr) << 12 | (uint64_t)c->pkt_used << 13;
    if (!c->graph_exec || c->graph_key != key) {
        if (c->graph_exec) { cudaGraphExecDestroy((cudaGraphExec_t)c->graph_exec); c->graph_exec = nullptr; }
        if (!c->brute_attr_set && score_mode == DPE_SCORE_BRUTE) {      // function attributes cannot be set while capturing
            int rc = brute_set_attributes(c);
            if (rc) return rc;
        }
        if (!c->vel_attr_set && with_vel == 2) {
            int rc = vel_brute_set_attributes(c);
            if (rc) return rc;
        }
        // a grid upload on another stream must be visible before the graph; inside the capture only captured events may be waited on
        DPE_CUDA(cudaStreamWaitEvent(s, c->ev_grid, 0));
        const int64_t l0 = c->launches;
        DPE_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        c->capturing = 1;
        int rc = compute_epoch(c, score_mode, est_mode, with_vel, s);
        c->capturing = 0;
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamEndCapture(s, &g);
        if (rc || e != cudaSuccess || !g) {
            if (g) cudaGraphDestroy(g);
            if (!rc) { set_error("stream capture of the epoch failed: %s", cudaGetErrorString(e)); rc = DPE_ECUDA; }
            cudaGetLastError();
            return rc;
        }
        c->graph_launches = c->launches - l0;
        c->launches = l0;
        cudaGraphExec_t ex = nullptr;
        e = cudaGraphInstantiate(&ex, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { set_error("cudaGraphInstantiate -> %s", cudaGetErrorString(e)); return DPE_ECUDA; }
        c->graph_exec = ex;
        c->graph_key = key;
    }
    DPE_CUDA(cudaGraphLaunch((cudaGraphExec_t)c->graph_exec, s));
    c->launches += c->graph_launches;
    c->have_block = 1;
    c->have_epoch = DPE_PART_CHANNELS | DPE_PART_GEOMETRY;
    c->have_prepare = c->have_corr = c->have_scores = 1;
    c->have_planes = (score_mode == DPE_SCORE_BRUTE);
    c->sort_valid = 0;
    c->sort_pending = 0;
    return DPE_OK;
}

static void unpack_result(const double* r, dpe_result* out) {
    memset(out, 0, sizeof(*out));
    for (int i = 0; i < 8; ++i) out->z[i] = r[i];
    out->max_score = r[8];
    out->sum_score = r[9];
    out->argmax = (int64_t)r[10];
    out->out_of_window = (int64_t)r[11];
    out->vel_max_score = r[12];
    out->vel_argmax = (int64_t)r[13];
    out->vel_out_of_window = (int64_t)r[14];
}

int dpe_epoch_run(dpe_ctx* c, const int16_t* iq_host, const dpe_epoch* ep, const double* sat_states,
                  int score_mode, int est_mode, int with_vel, dpe_result* out, void* stream) {
    DPE_REQUIRE(c && out, DPE_EINVAL, "dpe_epoch_run: null argument");
    DPE_REQUIRE(!c->inflight, DPE_ESTATE, "dpe_epoch_run while a submitted epoch is in flight");
    DevGuard guard(c->cfg.device);
    int rc = enqueue_epoch(c, iq_host, ep, sat_states, score_mode, est_mode, with_vel, (cudaStream_t)stream);
    if (rc) return rc;
    return dpe_result_fetch(c, out, stream);
}

int dpe_epoch_submit(dpe_ctx* c, const int16_t* iq, const dpe_epoch* ep, const double* sat_states,
                     int score_mode, int est_mode, int with_vel) {
    DPE_REQUIRE(c, DPE_EINVAL, "null context");
    DPE_REQUIRE(!c->inflight, DPE_ESTATE, "dpe_epoch_submit: collect the previous epoch first");
    DevGuard guard(c->cfg.device);
    cudaStream_t s = c->own_stream;
    int rc;
    // (a context with a communicator enqueues kernel by kernel: NCCL refuses thread-local stream capture)
    if (c->use_graph && !c->prof_on && !c->comm) {
        if ((rc = upload_epoch(c, iq, ep, sat_states, score_mode, est_mode, with_vel, s, true))) return rc;
        if ((rc = launch_epoch_graph(c, score_mode, est_mode, with_vel, s))) return rc;
    } else if ((rc = enqueue_epoch(c, iq, ep, sat_states, score_mode, est_mode, with_vel, s))) {
        return rc;
    }
    DPE_CUDA(cudaMemcpyAsync(c->res_pin, c->result, sizeof(double) * 16, cudaMemcpyDeviceToHost, s));
    DPE_CUDA(cudaEventRecord(c->ev_done, s));
    c->inflight = 1;
    return DPE_OK;
}

int dpe_epoch_collect(dpe_ctx* c, dpe_result* out) {
    DPE_REQUIRE(c && out, DPE_EINVAL, "null argument");
    DPE_REQUIRE(c->inflight, DPE_ESTATE, "dpe_epoch_collect without a submitted epoch");
    DevGuard guard(c->cfg.device);
    c->inflight = 0;
    DPE_CUDA(cudaEventSynchronize(c->ev_done));
    unpack_result(c->res_pin, out);
    return DPE_OK;
}

int dpe_epoch_pending(dpe_ctx* c) { return c ? c->inflight : 0; }

int dpe_epoch_run_dist(dpe_ctx* c, const int16_t* iq, const dpe_epoch* ep, const double* sat_states,
                       int score_mode, int est_mode, int with_vel, dpe_result* out) {
    int rc = dpe_epoch_submit(c, iq, ep, sat_states, score_mode, est_mode, with_vel);
    if (rc) return rc;
    return dpe_epoch_collect(c, out);
}

void* dpe_ctx_stream(dpe_ctx* c) { return c ? (void*)c->own_stream : nullptr; }

int dpe_kernel_attr(const char* kernel, int* regs, int* smem_bytes, int* max_threads) {
    DPE_REQUIRE(kernel, DPE_EINVAL, "null argument");
    cudaFuncAttributes a;
    memset(&a, 0, sizeof(a));
    if (!(kernel_attr_prepare(kernel, &a) || kernel_attr_score(kernel, &a) || kernel_attr_brute(kernel, &a) ||
          kernel_attr_vel(kernel, &a))) {
        cudaGetLastError();
        set_error("dpe_kernel_attr: no kernel '%s' (or no device)", kernel);
        return DPE_EINVAL;
    }
    if (regs) *regs = a.numRegs;
    if (smem_bytes) *smem_bytes = (int)a.sharedSizeBytes;
    if (max_threads) *max_threads = a.maxThreadsPerBlock;
    return DPE_OK;
}

const void* dpe_dev_ptr(dpe_ctx* c, int which) {
    if (!c) return nullptr;
    switch (which) {
        case DPE_PTR_SAMPLES: return c->iq;
        case DPE_PTR_CODE_SCORES: return c->cs;
        case DPE_PTR_POS_SCORES: return c->scores;
        case DPE_PTR_ZVAL: return c->zval;
        case DPE_PTR_RVAL: return c->rval;
        case DPE_PTR_GRID: return c->grid;
        case DPE_PTR_PARTIAL: return c->partial;
        case DPE_PTR_CHIP_IDX: return c->chip_idx;
        case DPE_PTR_XW: return c->xw;
        case DPE_PTR_CARR_SCORES: return c->carr;
        case DPE_PTR_VEL_SCORES: return c->vscores;
        case DPE_PTR_VEL_GRID: return c->vgrid;
        case DPE_PTR_REPLICA_SIGN: return c->rs;
        case DPE_PTR_CA_TABLE: return c->ca;
        default: return nullptr;
    }
}

int dpe_debug_channel_flags(dpe_ctx* c, int32_t* idx_next, int32_t* no_flip, int C) {
    DPE_REQUIRE(c && idx_next && no_flip, DPE_EINVAL, "null argument");
    DevGuard guard(c->cfg.device);
    DPE_REQUIRE(C >= 1 && C <= c->maxC, DPE_EINVAL, "bad C");
    DPE_CUDA(cudaDeviceSynchronize());
    DPE_CUDA(cudaMemcpy(idx_next, c->idx_next, sizeof(int32_t) * C, cudaMemcpyDeviceToHost));
    DPE_CUDA(cudaMemcpy(no_flip, c->no_flip, sizeof(int32_t) * C, cudaMemcpyDeviceToHost));
    return DPE_OK;
}

int dpe_debug_bins(dpe_ctx* c, int64_t i0, int64_t n, int sat_mode, int64_t* f_idx, double* alpha,
                   void* stream) {
    DPE_REQUIRE(c && f_idx && alpha, DPE_EINVAL, "null argument");
    DevGuard guard(c->cfg.device);
    DPE_REQUIRE(c->have_epoch & DPE_PART_GEOMETRY, DPE_ESTATE, "debug_bins before the geometry part of epoch_set");
    DPE_REQUIRE(i0 >= 0 && n >= 1 && i0 + n <= c->G, DPE_EINVAL, "candidate range outside the grid");
    const size_t cnt = (size_t)n * c->epoch_C;
    if (c->dbg_f) { cudaFree(c->dbg_f); cudaFree(c->dbg_alpha); c->dbg_f = nullptr; c->dbg_alpha = nullptr; }
    int rc;
    if ((rc = dev_alloc(&c->dbg_f, cnt))) return rc;
    if ((rc = dev_alloc(&c->dbg_alpha, cnt))) return rc;
    if ((rc = launch_debug_bins(c, i0, n, sat_mode, (cudaStream_t)stream))) return rc;
    DPE_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    DPE_CUDA(cudaMemcpy(f_idx, c->dbg_f, sizeof(int64_t) * cnt, cudaMemcpyDeviceToHost));
    DPE_CUDA(cudaMemcpy(alpha, c->dbg_alpha, sizeof(double) * cnt, cudaMemcpyDeviceToHost));
    return DPE_OK;
}

int dpe_debug_read(dpe_ctx* c, int which, size_t offset, void* dst, size_t nbytes) {
    DPE_REQUIRE(c && dst, DPE_EINVAL, "null argument");
    DevGuard guard(c->cfg.device);
    const char* p = static_cast<const char*>(dpe_dev_ptr(c, which));
    DPE_REQUIRE(p, DPE_ESTATE, "buffer %d not allocated", which);
    DPE_CUDA(cudaDeviceSynchronize());
    DPE_CUDA(cudaMemcpy(dst, p + offset, nbytes, cudaMemcpyDeviceToHost));
    return DPE_OK;
}

int64_t dpe_launch_count(dpe_ctx* c) { return c ? c->launches : -1; }

int dpe_stream_create(void** stream) {
    DPE_REQUIRE(stream, DPE_EINVAL, "null argument");
    cudaStream_t s;
    DPE_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (void*)s;
    return DPE_OK;
}
int dpe_stream_destroy(void* stream) {
    DPE_CUDA(cudaStreamDestroy((cudaStream_t)stream));
    return DPE_OK;
}
int dpe_stream_sync(void* stream) {
    DPE_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return DPE_OK;
}
int dpe_host_alloc(void** ptr, size_t bytes) {
    DPE_REQUIRE(ptr && bytes, DPE_EINVAL, "bad argument");
    DPE_CUDA(cudaMallocHost(ptr, bytes));
    return DPE_OK;
}
int dpe_host_free(void* ptr) {
    if (ptr) DPE_CUDA(cudaFreeHost(ptr));
    return DPE_OK;
}
int dpe_stream_create_on(void** stream, int device) {
    DPE_REQUIRE(stream, DPE_EINVAL, "null argument");
    DevGuard guard(device);
    cudaStream_t s;
    DPE_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (void*)s;
    return DPE_OK;
}
int dpe_device_alloc(void** ptr, size_t bytes, int device) {
    DPE_REQUIRE(ptr && bytes, DPE_EINVAL, "bad argument");
    DevGuard guard(device);
    DPE_CUDA(cudaMalloc(ptr, bytes));
    return DPE_OK;
}
int dpe_device_free(void* ptr, int device) {
    DevGuard guard(device);
    if (ptr) DPE_CUDA(cudaFree(ptr));
    return DPE_OK;
}
int dpe_copy_h2d(void* dst_device, const void* src_host, size_t bytes, void* stream, int device) {
    DPE_REQUIRE(dst_device && src_host, DPE_EINVAL, "null argument");
    DevGuard guard(device);
    DPE_CUDA(cudaMemcpyAsync(dst_device, src_host, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return DPE_OK;
}
int dpe_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int dpe_profile_enable(dpe_ctx* c, int on) {
    DPE_REQUIRE(c, DPE_EINVAL, "null context");
    DevGuard guard(c->cfg.device);
    if (on) {
        if (!c->prof_ev) c->prof_ev = new (std::nothrow) cudaEvent_t[2 * kProfMax]();      // null handles: destroy skips them
        if (!c->prof_stage) c->prof_stage = new (std::nothrow) int[kProfMax]();
        DPE_REQUIRE(c->prof_ev && c->prof_stage, DPE_ENOMEM, "out of host memory");
        for (int i = 0; i < 2 * kProfMax; ++i)
            if (!c->prof_ev[i]) DPE_CUDA(cudaEventCreate(&c->prof_ev[i]));
    }
    c->prof_on = on ? 1 : 0;
    c->prof_n = 0;
    return DPE_OK;
}

int dpe_profile_read(dpe_ctx* c, double* ms, int64_t* count) {
    DPE_REQUIRE(c && ms && count, DPE_EINVAL, "null argument");
    DevGuard guard(c->cfg.device);
    DPE_CUDA(cudaDeviceSynchronize());
    for (int i = 0; i < c->prof_n; ++i) {
        float t = 0.f;
        DPE_CUDA(cudaEventElapsedTime(&t, c->prof_ev[2 * i], c->prof_ev[2 * i + 1]));
        ms[c->prof_stage[i]] += t;
        count[c->prof_stage[i]] += 1;
    }
    c->prof_n = 0;
    return DPE_OK;
}

int64_t dpe_brute_pairs(dpe_ctx* c) {
    if (!c || !c->hist) return -1;
    DevGuard guard(c->cfg.device);
    const int n = c->epoch_C * (2 * c->W + 1);
    int32_t* h = new (std::nothrow) int32_t[n];
    if (!h) return -1;
    int64_t tot = -1;
    if (cudaDeviceSynchronize() == cudaSuccess &&
        cudaMemcpy(h, c->hist, sizeof(int32_t) * n, cudaMemcpyDeviceToHost) == cudaSuccess) {
        tot = 0;
        for (int i = 0; i < n; ++i) tot += h[i];
    }
    delete[] h;
    return tot;
}

}  // extern "C"
