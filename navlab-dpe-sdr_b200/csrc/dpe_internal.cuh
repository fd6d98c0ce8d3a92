// dpe_internal.cuh -- shared declarations of libdpe_b200 (sm_100a only).
//
// Data layout in HBM (one context = one flow = one GPU):
//   iq        int16  [2*S]              staged 20 ms block (interleaved I,Q)
//   ca        int8   [37][1024]         C/A chips, +/-1  (BCS chipsCACode_d)
//   xw        float2 [C][S]             wiped samples x[n]*conj(carrier)      (natural order)
//   rs        int8   [C][S]             no-flip replica sign r[n]
//   bx        float  [C][8][skewX(S_pad+2H)] wiped samples for the brute-force kernel: copy s holds
//                                       x[(p + s) mod S] at element p + H (circular halo H), (re,im)
//                                       interleaved, float4-skewed (conflict-free LDS.128).  Lag
//                                       k = 8 q + s reads copy s at element offset 8 q: every tile
//                                       a slot streams is aligned on the replica position p
//   brd       float  [C][skewX(S_pad)]  chosen replica (flip applied) by position pair:
//                                       (d[p], d[p+1], r[p], r[p+1]), d[p] = r[p-1] - r[p] (circular),
//                                       zero beyond S; same float4 skew
//   cacc      int64  [C][NLp][4]        correlogram totals in 2^-19 fixed point: (A.re, A.im, B.re, B.im) per lag, A: n<idxNext, B: rest
//                                       (integer atomics from every chunk's CTA: order-independent; cleared by the reader)
//   cs        double2[C][NL]            fft-shifted "CodeScores" window, lag k = l - W
//   grid      double [G][4]             candidates (ENU metres + clock metres)
//   scores    double [G]                "PosScores"
//   pair_*    per (channel, candidate) lag (int16) / alpha (float) / correlation v (float2) for the brute-force path
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/dpe_b200.h"

// utils/inc/consthelper.h:5-27 of the reference
#define K_C      (299792458.0)
#define K_PI     (3.1415926535898)
#define K_F_L1   (1.57542e9)
#define K_F_CA   (1.023e6)
#define K_L_CA   (1023)
#define K_T_CA   (0.001)
#define K_OEDOT  (7.2921151467e-5)

namespace dpe {

#ifndef DPE_CORR_CHUNK
#define DPE_CORR_CHUNK 1024
#endif
constexpr int kCorrChunk = DPE_CORR_CHUNK;   // samples per partial-correlogram block (k_prep_corr)
constexpr int kCarrChunk = 1024;       // samples per partial carrier-spectrum block (k_carr_partial)
constexpr int kLagTile = 8;            // lags per thread in the correlogram kernel
constexpr int kPartialLen = DPE_PARTIAL_LEN;
constexpr int kReduceBlock = 128;
constexpr int kSortBlock = 128;        // candidates per CTA of the pair sort (k_pair_bins / k_scatter)
constexpr int kProfMax = 8192;
constexpr int kPinSlots = 8;
constexpr double kFixScale = 524288.0;  // 2^19: fixed-point scale of the chunk-partial accumulators (k_prep_corr, k_carr_partial)

// brute-force kernel geometry
constexpr int kBfNC = 32;              // candidates per warp (one group)
constexpr int kBfNS = 8;               // contiguous samples per lane per chunk
constexpr int kBfChunk = 32 * kBfNS;   // samples per warp-chunk (256)
constexpr int kBfTile = 1024;          // samples per TMA stage
constexpr int kBfWarps = 8;            // consumer warps per CTA
constexpr int kBfStages = 4;            // TMA stages of 20 KB (sample tile + replica tile)
constexpr int kBfLag = 1;               // a stage is refilled this many tiles after its release (8 / 4: no gain)

// Every kernel except k_brute is a "side" kernel: 128 threads x <= 80 registers or 256 threads x <= 40
// registers (10 240 registers), so that one CTA of it fits on an SM beside the persistent k_brute CTA
// (256 threads x 216 registers = 55 296 of 65 536).  With two contexts in flight the pre-pass, the pair
// sort and the reductions of one epoch then run UNDER the k_brute of the other.
#define DPE_SIDE128 __launch_bounds__(128, 6)
#define DPE_SIDE256 __launch_bounds__(256, 6)

// Phase stamps of the single-wave kernels (probe builds only: make VARIANT=phase adds -DDPE_PHASE_TIMING; the product
// build compiles these to nothing).  Thread 0 of a CTA keeps %globaltimer (ns) at every mark and prints the differences.
#ifdef DPE_PHASE_TIMING
#define DPE_PT_DECL unsigned long long pt_[10]; int pt_n_ = 0
#define DPE_PT_MARK() do { if (threadIdx.x == 0 && pt_n_ < 10) { unsigned long long t_; \
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); pt_[pt_n_++] = t_; } } while (0)
#define DPE_PT_PRINT(tag, sel) do { if (threadIdx.x == 0 && (sel)) { unsigned int sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_)); \
        printf("PT %s cta %d,%d sm %u t0 %llu :", tag, (int)blockIdx.x, (int)blockIdx.y, sm_, pt_[0]); \
        for (int q_ = 1; q_ < pt_n_; ++q_) printf(" %llu", pt_[q_] - pt_[0]); printf("\n"); } } while (0)
#else
#define DPE_PT_DECL do { } while (0)
#define DPE_PT_MARK() do { } while (0)
#define DPE_PT_PRINT(tag, sel) do { } while (0)
#endif

// Programmatic dependent launch (sm_90+): a scoring kernel launched with launch_dep(..., pdl = true) may start while the
// kernel before it on the stream is still running -- its prologue (per-channel constants, candidate states) overlaps the
// producer's tail -- and calls grid_dep_wait() before its first read of what the producer writes.  The producer calls
// grid_dep_trigger() at its start (all of its CTAs are resident by then: nothing it needs is taken away).  Without the
// launch attribute both calls are no-ops.
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_dep(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl,
                                     Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Device copy of the per-epoch parameters (+ values derived on the device).
struct EpochDev {
    int32_t C, doppler_sign;
    double rx_time;
    double center[8];
    double R[9];
    uint8_t prn[DPE_MAX_CHAN + 3];
    double rc_start[DPE_MAX_CHAN], ri_start[DPE_MAX_CHAN], fc[DPE_MAX_CHAN], fi[DPE_MAX_CHAN];
    int32_t cp_start[DPE_MAX_CHAN], cp_ref[DPE_MAX_CHAN];
    double rc_end[DPE_MAX_CHAN];
    int32_t cp_end[DPE_MAX_CHAN], cp_ref_tow[DPE_MAX_CHAN];
};

// float4-skew of the brute-force planes: float offset of element n (a sample (re,im) or half of a
// replica position pair).  One float4 holds two elements; one pad float4 after every 4 (lane l
// reads float4 4l..4l+3 -> physical 5l..5l+3, conflict-free LDS.128 because 5 is odd).
__host__ __device__ constexpr inline int64_t skewX(int64_t n) { return 4 * ((n >> 1) + (n >> 3)) + 2 * (n & 1); }

}  // namespace dpe

struct dpe_ctx {
    dpe_cfg cfg;
    int64_t S, S_pad, G, Gv;
    int32_t W, NL, NLp, H;             // lag half width, lags, padded lags, replica halo
    int32_t nchunk;                    // correlogram chunks per channel
    int32_t vnchunk;                   // carrier-spectrum chunks per channel
    int32_t maxC, T;
    int sm_count;
    // device buffers
    // epoch packet: {iq int16[2S] | EpochDev | sat double[maxC][T][8]} -- one H2D / one ncclBroadcast per epoch
    unsigned char* pkt; unsigned char* pkt_pin;
    size_t pkt_off_ep, pkt_off_sat, pkt_bytes;
    int16_t* iq_own; const int16_t* iq;    // iq_own = pkt
    int8_t* ca;
    double* tidx;                      // [S] time index t_n (BCS_GenTimeIdcs), built once
    unsigned int* chan_ticket;         // [maxC] arrivals per channel in k_prep_corr (self-resetting)
    dpe::EpochDev* ep;                 // = pkt + pkt_off_ep
    double* sat;                       // = pkt + pkt_off_sat, [C][T][8]
    double* sat_geo;                   // dpe::SatGeo [C][T]: centre-relative geometry per satellite state (DPE_SAT_PER_TIME)
    float2* xw; int8_t* rs; int16_t* chip_idx;
    int32_t* idx_next; int32_t* no_flip;
    long long* cacc; double2* cs;      // cacc [maxC][NLp][4]
    float *bx, *brd;
    int64_t bx_stride, brd_stride;     // floats per sample-plane copy (8 copies per channel) / per replica plane
    double* grid; double* scores;
    double* blk_partial; int32_t n_blk_partial;
    unsigned int* ticket;              // last-CTA ticket counters (self-resetting): [0] position scoring, [1] pair sort, [2] velocity pair sort, [3] velocity scoring
    double* partial; double* zval; double* rval; double* result;  // result: device mirror of dpe_result
    // brute-force work lists
    int16_t* pair_k; float* pair_a; float2* pair_v;    // [C][G]
    int32_t* hist;                     // [C][NB] counts, NB = 2W+1
    int32_t* blk_hist;                 // [C*NB][ceil(G/256)] per-block counts, then exclusive block prefixes
    int64_t* bucket_base; int32_t* group_base;
    int32_t* hdr;                      // group headers {c, krel, n, pad}
    int32_t* ent_j; float* ent_a;      // bucketed entries
    int32_t* n_groups;                 // device scalar: total groups (multiple of kBfWarps)
    int64_t max_groups;
    float2* tail_part;                 // [sm_count][2][kBfWarps * kBfNC] partial sums of the slots a CTA boundary cuts
    unsigned int* tail_ticket;         // [sm_count][kBfWarps] arrivals per cut slot and warp (self-resetting)
    // debug
    int64_t* dbg_f; double* dbg_alpha;
    // velocity (section 8 f-1)
    double* vgrid; double* vscores; double2* carr;       // [Gv][4], [Gv], [C][NBd]
    long long* dc_part; float2* bb; long long* vacc; unsigned int* vticket; double* vblk_partial;   // vacc [maxC][NBd][2]: carrier-spectrum totals, 2^-19 fixed point; vticket [maxC];   // [nchunk][2] DC sums per chunk; bb = conj(carrier) plane [C][S]
    // brute-force velocity manifold (DPE_FLAG_BRUTE_VEL): baseband plane with the chosen replica applied + pair lists
    float2* vbb; int64_t vS_pad;
    int16_t* vpair_k; float* vpair_a; float2* vpair_v;
    int32_t *vhist, *vblk_hist, *vgroup_base, *vhdr, *vent_j, *vn_groups; int64_t* vbucket_base; float* vent_a;
    int64_t vmax_groups; int vel_attr_set;
    int stage_fold_est, folded_est;    // dpe_fold_estimate's mode (-1 = off); est_mode + 1 of the estimate the last dpe_score_pos wrote, 0 = none
    int fold_est_mode;                 // >= 0: the scoring kernel's last CTA also writes the single-rank estimate (dpe_epoch_*)
    int vel_weighted;                  // the next launch_score_vel forms the score-weighted estimate (dpe_score_vel_est)
    int carr_direct;                   // debug switch: evaluate the carrier spectrum directly (k_carr_partial_direct)
    int32_t Wd, NBd, n_fft; int have_vgrid;
    // state
    int have_block, have_epoch, have_prepare, have_corr, have_scores;
    int epoch_C;
    int brute_attr_set;
    int brute_skip_pad;                // padding groups skip their FFMA work (DPE_BRUTE_SKIP_PAD, default 1)
    int have_planes;                   // brute-force planes match the current prepare + correlogram
    int sort_valid;                    // 0, or 1 + sat_mode the brute-force work lists were built for (this epoch)
    int sort_pending;                  // a presort on another stream has not been waited for yet
    cudaEvent_t ev_epoch, ev_grid, ev_sort;   // epoch upload done / grid upload done / presort done
    cudaStream_t aux_stream;           // dpe_epoch_run's own presort stream (created on first use)
    // the velocity manifold (lookup formulation) runs on its own stream beside the position scoring: dpe_score_pos records
    // ev_fork before its first launch, dpe_score_vel enqueues on vel_stream behind it and joins the caller's stream again
    cudaStream_t vel_stream; cudaEvent_t ev_fork, ev_vel; int fork_valid, vel_fork;
    int64_t launches;
    int lk_cand_forced;                // DPE_LK_CAND = 3 | 4 | 6: candidates per thread of k_score_lookup (0 = chosen per launch)
    int use_pdl;                       // DPE_PDL (default 1): the scoring kernels of the lookup path are launched as programmatic dependents
    int want_sums;                     // k_score_pairs accumulates sum s*x (0 only inside an arg-max dpe_epoch_submit)
    // asynchronous epochs (dpe_epoch_submit / dpe_epoch_collect)
    cudaStream_t own_stream;
    cudaEvent_t ev_done;
    double* res_pin;                   // page-locked mirror of `result` [16]
    int inflight;
    size_t pkt_used;                   // bytes of the packet the current epoch uses (broadcast size)
    // the device work of dpe_epoch_submit as one CUDA graph
    int use_graph, capturing;
    void* graph_exec;                  // cudaGraphExec_t
    uint64_t graph_key;
    int64_t graph_launches;
    // multi-GPU
    void* comm;                        // ncclComm_t
    int nranks, rank;
    int comm_reserve_sms;              // SMs k_brute leaves free for the NCCL kernels (DPE_COMM_RESERVE_SMS, default 1)
    double* gathered;                  // [nranks][kPartialLen]
    dpe::EpochDev ep_host;
    // page-locked staging ring for the per-epoch uploads (no implicit stream sync, safe reuse)
    dpe::EpochDev* ep_pin; double* sat_pin; cudaEvent_t pin_ev[dpe::kPinSlots]; int pin_next; size_t sat_cap;
    // per-stage event brackets (dpe_profile_*)
    int prof_on; int prof_n;                 // brackets recorded since the last read
    cudaEvent_t* prof_ev;                    // [2 * kProfMax]
    int* prof_stage;                         // [kProfMax]
};

namespace dpe {
void set_error(const char* fmt, ...);
#define DPE_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            dpe::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                 \
                           cudaGetErrorString(e__));                                     \
            return DPE_ECUDA;                                                            \
        }                                                                                \
    } while (0)

// stage brackets: no-ops unless dpe_profile_enable(ctx, 1)
void prof_begin(dpe_ctx* c, int stage, cudaStream_t s);
void prof_end(dpe_ctx* c, cudaStream_t s);

// launchers (each returns DPE_OK / DPE_ECUDA and bumps ctx->launches)
int launch_gen_ca(dpe_ctx* c, cudaStream_t s);
int launch_prepare(dpe_ctx* c, cudaStream_t s);
int launch_gen_time(dpe_ctx* c, cudaStream_t s);
int launch_correlogram(dpe_ctx* c, cudaStream_t s);
int launch_score_lookup(dpe_ctx* c, int sat_mode, cudaStream_t s);
int launch_sat_geo(dpe_ctx* c, cudaStream_t s);
int launch_score_brute(dpe_ctx* c, int sat_mode, cudaStream_t s);
struct SortLists {                 // one set of pair-sort work lists (position pairs, or velocity pairs)
    int16_t* pair_k; float* pair_a; int32_t* blk_hist; int32_t* hist; int32_t* group_base; int64_t* bucket_base;
    int32_t* n_groups; int32_t* hdr; int32_t* ent_j; float* ent_a; int64_t max_groups;
};
int launch_sort_tail(dpe_ctx* c, const SortLists& L, int64_t G, int W, int C, unsigned int* ticket, cudaStream_t s);
int launch_score_vel_brute(dpe_ctx* c, cudaStream_t s);
int vel_brute_set_attributes(dpe_ctx* c);
int launch_brute_corr(dpe_ctx* c, cudaStream_t s);
int brute_set_attributes(dpe_ctx* c);
int launch_brute_score(dpe_ctx* c, cudaStream_t s);
// cudaFuncGetAttributes of a kernel by name, one lookup per translation unit (1 = found)
int kernel_attr_prepare(const char* name, cudaFuncAttributes* a);
int kernel_attr_score(const char* name, cudaFuncAttributes* a);
int kernel_attr_brute(const char* name, cudaFuncAttributes* a);
int kernel_attr_vel(const char* name, cudaFuncAttributes* a);
#define DPE_KATTR(kname, sym) if (!strcmp(name, kname)) return cudaFuncGetAttributes(a, sym) == cudaSuccess
// NCCL (dlopen'ed, dpe_comm.cu)
int comm_broadcast(dpe_ctx* c, void* buf, size_t bytes, cudaStream_t s);
int comm_allgather(dpe_ctx* c, const double* send, double* recv, size_t count, cudaStream_t s);
// RAII: bind the context's device for the duration of an extern "C" entry, restore the caller's on exit
struct DevGuard {
    int prev, dev;
    explicit DevGuard(int d) : prev(-1), dev(d) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DevGuard() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
};
int launch_brute_planes(dpe_ctx* c, cudaStream_t s);
int launch_brute_sort(dpe_ctx* c, int sat_mode, cudaStream_t s);
int launch_score_vel(dpe_ctx* c, cudaStream_t s);
size_t brute_smem_bytes(int H);
int launch_estimate(dpe_ctx* c, int est_mode, const double* gathered, int nranks, cudaStream_t s);
int launch_debug_bins(dpe_ctx* c, int64_t i0, int64_t n, int sat_mode, cudaStream_t s);
}  // namespace dpe
