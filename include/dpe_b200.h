/*
 * dpe_b200.h -- C ABI of the B200-native DPE batch-correlation-manifold hot path.
 *
 * The reference (Stanford-NavLab/NavLab-DPE-SDR, CUDARecv) has no C ABI of its
 * own: its operator interface is the C++ class dsp::Module with named Ports and
 * Params (cudarecv/modules/inc/module.h:13-144, cudarecv/dsp/inc/dsp.h:70-146).
 * The entry points below are what the three hot-path modules' Start/Update
 * bodies bind to; each one names the reference code it replaces.  The C++
 * module mirror in navlab-dpe-sdr_b200/host/ and the Python ctypes binding in
 * navlab-dpe-sdr_b200/capi.py call nothing else.
 *
 * Conventions: every function returns 0 on success and a negative DPE_E* code
 * otherwise (maps onto the reference's "-1 = fatal, stop the flow",
 * cudarecv/dsp/src/flow.cu:125-131); dpe_last_error() gives the text.  Plain
 * pointers and sizes only.  The context owns all device memory; the caller
 * owns every host array.  All work is asynchronous on the caller's stream
 * (a cudaStream_t passed as void*, like Module::Update(void* cuFlowStream),
 * module.h:23) unless a function is documented as synchronising.  One context
 * per flow / per GPU; no internal locking (the reference calls every module
 * from one flow thread, flow.cu:105-137).
 */
#ifndef DPE_B200_H_
#define DPE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPE_MAX_CHAN 37          /* CONST_PRN_MAX, utils/inc/consthelper.h:14 */
#define DPE_ABI_VERSION 2
#define DPE_PARTIAL_LEN 16          /* doubles per per-rank partial estimate         */

enum {
    DPE_OK = 0,
    DPE_EINVAL = -1,             /* bad argument / size / state                */
    DPE_ECUDA = -2,              /* CUDA runtime error (see dpe_last_error)    */
    DPE_ENOMEM = -3,
    DPE_ESTATE = -4,             /* call order violated (e.g. score before prep) */
    DPE_EWINDOW = -5,            /* candidates fell outside the lag window     */
    DPE_ECOMM = -6               /* NCCL missing / failed (see dpe_last_error) */
};

/* dpe_score_pos scoring modes */
enum {
    DPE_SCORE_LOOKUP = 0,        /* correlogram bin lookup + lerp: BCM_PosMeasML,
                                    batchcorrmanifold.cu:1710-1828               */
    DPE_SCORE_BRUTE = 1          /* north-star kernel: every (candidate, PRN) pair
                                    correlates the whole block against its own
                                    blended replica (SURVEY.md section 8 a')     */
};

/* dpe_estimate modes */
enum {
    DPE_EST_ARGMAX = 0,          /* thrust::max_element + BCM_MakePosMeas,
                                    batchcorrmanifold.cu:2589,1977-2016 (first max) */
    DPE_EST_WEIGHTED = 1         /* BCM_PosMeasReduction + BCM_ReduceAndPosMeas,
                                    batchcorrmanifold.cu:816-1056,1365-1510       */
};

/* which satellite state of the [C][T] batch a candidate uses */
enum {
    DPE_SAT_MIDDLE = 0,          /* batchcorrmanifold.cu:1773-1775 (ML kernel)   */
    DPE_SAT_PER_TIME = 1         /* batchcorrmanifold.cu:865-873 (reduction kernel) */
};

/* dpe_dev_ptr selectors */
enum {
    DPE_PTR_SAMPLES = 0,         /* int16 I,Q [2*S]        (SampleBlock "Samples")   */
    DPE_PTR_CODE_SCORES = 1,     /* double2 [C][2W+2]: the window of BCS "CodeScores"
                                    a position grid can reach; entry l = fft-shifted
                                    bin S/2 - W + l = circular lag l - W            */
    DPE_PTR_POS_SCORES = 2,      /* double [G]             (BCM "PosScores")         */
    DPE_PTR_ZVAL = 3,            /* double [8]             (BCM "zVal")              */
    DPE_PTR_RVAL = 4,            /* double [64]            (BCM "RVal")              */
    DPE_PTR_GRID = 5,            /* double [G][4]          (gridPosLocs_d)           */
    DPE_PTR_PARTIAL = 6,         /* double [DPE_PARTIAL_LEN] per-rank partial (below)*/
    DPE_PTR_CHIP_IDX = 7,        /* int16 [C][S] C/A chip index (debug flag only)    */
    DPE_PTR_XW = 8,              /* float2 [C][S] wiped samples                      */
    DPE_PTR_CARR_SCORES = 9,     /* double2 [C][NB] windowed carrier spectrum        */
    DPE_PTR_VEL_SCORES = 10,     /* double [Gv]                                      */
    DPE_PTR_VEL_GRID = 11,
    DPE_PTR_REPLICA_SIGN = 12,   /* int8 [C][S] no-flip replica chips (+1/-1)         */
    DPE_PTR_CA_TABLE = 13        /* int8 [37][1024] C/A code table (chipsCACode_d)    */
};

/* cfg.flags */
#define DPE_FLAG_KEEP_CHIP_IDX 1u   /* store chip indices (parity tests)          */
#define DPE_FLAG_BRUTE_TILES   2u   /* allocate + emit the brute-force tiles      */
#define DPE_FLAG_KEEP_BINS     4u   /* store per-(candidate,PRN) bins (parity)    */
#define DPE_FLAG_BRUTE_VEL     8u   /* allocate the work lists of the brute-force velocity manifold (needs Gv > 0) */

typedef struct dpe_ctx dpe_ctx;

typedef struct dpe_cfg {
    uint32_t abi_version;        /* DPE_ABI_VERSION                               */
    int32_t  device;             /* CUDA device ordinal                           */
    double   fs;                 /* sampling frequency, Hz (SampleBlock param)    */
    int64_t  S;                  /* samples per block = round(fs*T), even.  The
                                    reference keeps this in an unsigned short
                                    (sampleblock.h:81); 64-bit here so 10 MHz works */
    int32_t  max_chan;           /* <= DPE_MAX_CHAN                               */
    int32_t  time_dim;           /* T: time-grid points = sat states per channel  */
    int64_t  G;                  /* position-clock candidates held by THIS context
                                    (its shard of the grid)                       */
    int64_t  grid_offset;        /* global index of local candidate 0             */
    int64_t  G_total;            /* global grid size (for index % time_dim)       */
    int32_t  lpower;             /* L in sum |.|^L ("LPower", bcm.cu:2290)        */
    int32_t  lag_halfwidth;      /* W: correlogram lags -W..W+1 are produced      */
    uint32_t flags;
    int64_t  Gv;                 /* velocity-drift candidates (0 = no vel grid)   */
    int32_t  n_fft;              /* carrier spectrum length N_c (bcs.cu:761); 0 =
                                    8*2^ceil(log2 S)                              */
    int32_t  dopp_halfwidth;     /* carrier bins -Wd..Wd+1 around N_c/2           */
} dpe_cfg;

/* Per-epoch channel parameters.  "start" = referenced to the first sample of the
 * block (BatchCorrScores inputs, batchcorrscores.cu:681-692); "end" = referenced
 * to the end of the block (BatchCorrManifold inputs, dpeflow.cpp:187-191).      */
typedef struct dpe_epoch {
    int32_t C;                              /* tracked channels                  */
    int32_t doppler_sign;                   /* +1 / -1                           */
    uint8_t prn[DPE_MAX_CHAN + 3];          /* "ValidPRNs"                       */
    double  rc_start[DPE_MAX_CHAN];         /* "CodePhaseStart"    (chips)       */
    double  ri_start[DPE_MAX_CHAN];         /* "CarrierPhaseStart" (cycles)      */
    double  fc[DPE_MAX_CHAN];               /* "CodeFrequency"     (Hz)          */
    double  fi[DPE_MAX_CHAN];               /* "CarrierFrequency"  (Hz)          */
    int32_t cp_start[DPE_MAX_CHAN];         /* "cpElapsedStart"                  */
    int32_t cp_ref[DPE_MAX_CHAN];           /* "cpReference" / "cpRef"           */
    double  rc_end[DPE_MAX_CHAN];           /* "CodePhase" (= CodePhaseEnd)      */
    int32_t cp_end[DPE_MAX_CHAN];           /* "cpElapsedEnd"                    */
    int32_t cp_ref_tow[DPE_MAX_CHAN];       /* "cpRefTOW"                        */
    double  rx_time;                        /* "rxTime" (host double, already += T) */
    double  center[8];                      /* "xCurrkk1"                        */
    double  enu2ecef[9];                    /* "ENU2ECEFMat" row-major           */
} dpe_epoch;

/* Result of one epoch, host-visible. */
typedef struct dpe_result {
    double  z[8];                /* zVal: ECEF x,y,z, c*dt, then velocity part    */
    double  max_score;           /* score of the arg-max candidate                */
    double  sum_score;           /* sum of all candidate scores                   */
    int64_t argmax;              /* GLOBAL candidate index (lowest index on ties) */
    int64_t out_of_window;       /* (candidate, PRN) pairs that could not be scored */
    double  vel_max_score;
    int64_t vel_argmax;          /* index into the velocity grid                  */
    int64_t vel_out_of_window;
} dpe_result;

/* ---- lifetime ---------------------------------------------------------------
 * Replaces the cudaMalloc blocks of BatchCorrScores::Start
 * (batchcorrscores.cu:748-865) and BatchCorrManifold::Start
 * (batchcorrmanifold.cu:2363-2405).  Synchronises.                              */
int dpe_ctx_create(dpe_ctx** out, const dpe_cfg* cfg);
int dpe_ctx_destroy(dpe_ctx* ctx);
const char* dpe_last_error(void);
int dpe_abi_version(void);

/* ---- grid --------------------------------------------------------------------
 * dpe_grid_set: upload this context's G candidates, host double [G][4] =
 * {x, y, z, delta_t} ENU metres, flat order x slowest / t fastest.  Replaces
 * BCM_InitPosGrid + the CSV upload, batchcorrmanifold.cu:2421-2448 (the CSV
 * itself stays host code).  dpe_vel_grid_set: same for BCM_InitVelGrid (:2449). */
int dpe_grid_set(dpe_ctx* ctx, const double* enu_dt, int64_t G, void* stream);
int dpe_vel_grid_set(dpe_ctx* ctx, const double* venu_ddt, int64_t Gv, void* stream);

/* ---- per-epoch stages ---------------------------------------------------------
 * dpe_block_stage: make one 20 ms block of interleaved little-endian int16 I,Q
 * resident.  `iq` may be a host pointer (pinned or pageable; copied with
 * cudaMemcpyAsync) or a device pointer (used in place, zero copy).  Replaces
 * the H2D of SampleBlock::GetSamplesThread (sampleblock.cu:403) + the pointer
 * hand-over of SampleBlock::Update (:508).                                      */
int dpe_block_stage(dpe_ctx* ctx, const int16_t* iq, int64_t S, void* stream);

/* dpe_epoch_set: upload the channel / geometry parameters of this epoch and the
 * satellite states, host double [C][T][8] (state_t<double>, statehelper.h:11-21,
 * as produced by CHM_GridPrep, cuchanmgr.cu:853-923).                            */
int dpe_epoch_set(dpe_ctx* ctx, const dpe_epoch* ep, const double* sat_states, void* stream);

/* dpe_epoch_set_part: the same upload split the way the reference splits its modules.
 *   DPE_PART_CHANNELS  what BatchCorrScores reads (batchcorrscores.cu:681-692): C, prn,
 *                      rc_start, ri_start, fc, fi, cp_start, cp_ref, doppler_sign
 *   DPE_PART_GEOMETRY  what BatchCorrManifold reads (batchcorrmanifold.cu:2261-2279): fc,
 *                      rc_end, cp_end, cp_ref, cp_ref_tow, rx_time, center, enu2ecef and
 *                      sat_states ([C][T][8]; may be NULL when the part is not selected)
 * Fields of the other part are ignored.  C must agree between the two parts.        */
#define DPE_PART_CHANNELS 1u
#define DPE_PART_GEOMETRY 2u
int dpe_epoch_set_part(dpe_ctx* ctx, const dpe_epoch* ep, const double* sat_states, unsigned parts,
                       void* stream);

/* dpe_epoch_set_device: the same upload when the parameters already live on the DEVICE -- the
 * reference's cuChanMgr and cuEKF publish every one of them as a CUDA_DEVICE port
 * (cuchanmgr.cu:973-990,1136-1171; cuekf.cu xCurrkk1) and BatchCorrScores / BatchCorrManifold read them
 * there (batchcorrscores.cu:991-1005, batchcorrmanifold.cu:2512-2533).  A one-block kernel on `stream`
 * packs them; nothing is copied through the host.  Pointers of a part that is not selected may be NULL.
 * rx_time is a host value (the "rxTime" port is HOST, cuchanmgr.cu:973).                          */
typedef struct dpe_epoch_dev {
    int32_t C;                       /* host: VectorLength of the channel ports                      */
    const uint8_t* prn;              /* "ValidPRNs"          char   [C]                               */
    const double*  rc_start;         /* "CodePhaseStart"     double [C]                               */
    const double*  ri_start;         /* "CarrierPhaseStart"                                           */
    const double*  fc;               /* "CodeFrequency"                                               */
    const double*  fi;               /* "CarrierFrequency"                                            */
    const int32_t* cp_start;         /* "cpElapsedStart"     int    [C]                               */
    const int32_t* cp_ref;           /* "cpReference" / "cpRef"                                       */
    const int32_t* doppler_sign;     /* "DopplerSign"        int    [1]                               */
    const double*  rc_end;           /* "CodePhase" (= CodePhaseEnd)                                  */
    const int32_t* cp_end;           /* "cpElapsedEnd"                                                */
    const int32_t* cp_ref_tow;       /* "cpRefTOW"                                                    */
    const double*  center;           /* "xCurrkk1"           double [8]                               */
    const double*  enu2ecef;         /* "ENU2ECEFMat"        double [9]                               */
    const double*  sat_states;       /* "SatStates"          double [C][T][8]                         */
    double rx_time;                  /* "rxTime" (host double, already += T)                          */
} dpe_epoch_dev;
int dpe_epoch_set_device(dpe_ctx* ctx, const dpe_epoch_dev* ep, unsigned parts, void* stream);

/* dpe_replica_prepare: int16 unpack, carrier NCO wipe-off, C/A chip index and
 * replica sign, nav-bit edge, per-lag partial correlations.  Replaces BCS_Load,
 * BCS_NavBitBoundary, BCS_ComputeDopplerWipeoff, BCS_ComputeCodeReplica,
 * BCS_BatchMultiply (batchcorrscores.cu:209-407, launches :1048-1113).          */
int dpe_replica_prepare(dpe_ctx* ctx, void* stream);

/* dpe_correlogram: finish the circular code correlogram on lags -W..W+1, choose
 * flip / no-flip per channel, write the fft-shifted "CodeScores" rows.  Replaces
 * the cuFFT chain + BCS_ChooseCodeCorr + BCS_cufftBatchShift
 * (batchcorrscores.cu:1099-1153).  (The planes the brute-force kernel streams are
 * built by the first dpe_score_pos(DPE_SCORE_BRUTE) of the epoch.)                */
int dpe_correlogram(dpe_ctx* ctx, void* stream);

/* dpe_code_scores_set: use a correlogram produced elsewhere (e.g. the reference's own
 * BatchCorrScores feeding this BatchCorrManifold through the "CodeScores" port,
 * batchcorrmanifold.cu:2261): `cs` = [C][2W+2] complex doubles (host or device), entry
 * l of channel c = fft-shifted bin S/2 - W + l of that channel's row.  Marks the
 * correlogram stage done; needs dpe_epoch_set first.  Lookup scoring only.       */
int dpe_code_scores_set(dpe_ctx* ctx, const double* cs, int C, void* stream);

/* dpe_score_pos: score every candidate of this context (BCM_PosMeasML /
 * BCM_PosMeasReduction scoring part) into "PosScores" and reduce the block-level
 * arg-max / weighted sums into the per-rank partial (DPE_PTR_PARTIAL):
 *   partial[0..3] = sum_i s_i * (x, y, z, c*dt)_i   partial[4] = sum_i s_i
 *   partial[5] = max score   partial[6] = global arg-max index (exact in double)
 *   partial[7] = out-of-window pair count
 *   partial[8..11] = ECEF x,y,z and clock (m) of this rank's arg-max candidate     */
int dpe_score_pos(dpe_ctx* ctx, int score_mode, int sat_mode, void* stream);

/* dpe_brute_presort (optional): bin and sort the (candidate, PRN) pairs of the current
 * epoch for DPE_SCORE_BRUTE ahead of time -- typically on a second stream, while
 * dpe_replica_prepare / dpe_correlogram run on the first: the sort needs dpe_epoch_set
 * only, not the samples.  Ordering is handled inside: the sort waits for the epoch
 * upload, dpe_score_pos(DPE_SCORE_BRUTE) with the same sat_mode waits for the sort and
 * skips its own, and the next dpe_epoch_set waits for it before it overwrites the
 * parameters.  dpe_epoch_run does this by itself on an internal stream.           */
int dpe_brute_presort(dpe_ctx* ctx, int sat_mode, void* stream);

/* dpe_estimate: turn partial(s) into zVal[0:4] / RVal rows 0-3.  `gathered` is
 * either NULL (single GPU: use this context's own partial) or a DEVICE pointer
 * to nranks*DPE_PARTIAL_LEN doubles (the all-gathered partials; rank order =
 * ascending grid_offset).  Replaces thrust::max_element + BCM_MakePosMeas or
 * BCM_ReduceAndPosMeas (batchcorrmanifold.cu:2589-2596, 1365-1510).             */
int dpe_estimate(dpe_ctx* ctx, int est_mode, const double* gathered, int nranks, void* stream);
/* dpe_fold_estimate: a single-GPU stage-by-stage caller (one context = the whole grid) lets dpe_score_pos write the
 * estimate itself, in the tail of the scoring kernel -- what dpe_epoch_run does; the following
 * dpe_estimate(ctx, est_mode, NULL, ...) with the same mode then has nothing left to launch.  est_mode = -1 turns
 * it off again (sharded grids: the partials must be gathered first).                                            */
int dpe_fold_estimate(dpe_ctx* ctx, int est_mode);

/* Velocity-drift manifold: DC-removed carrier branch on
 * carrier bins -Wd..Wd+1 (direct DFT of the zero-padded spectrum's bins) +
 * BCM_VelMeasML + BCM_MakeVelMeas (batchcorrscores.cu:1158-1180,
 * batchcorrmanifold.cu:1861-1963,2030-2068).  Fills zVal[4:8].                   */
int dpe_score_vel(dpe_ctx* ctx, void* stream);
/* dpe_score_vel_est: the same manifold with the estimator chosen -- DPE_EST_ARGMAX (what dpe_score_vel does) or
 * DPE_EST_WEIGHTED, the reference's dormant BCM_VelMeasReduction + BCM_ReduceAndVelMeas (batchcorrmanifold.cu:1090-1347,
 * 1525-1667): zVal[4:8] = sum_i score_i v_i / sum_i score_i over the ECEF velocity candidates.  In dpe_epoch_run /
 * dpe_epoch_submit: with_vel = 3.                                                                                */
int dpe_score_vel_est(dpe_ctx* ctx, int est_mode, void* stream);
/* dpe_score_vel_brute: the same fix with every (velocity candidate, PRN) pair correlating the whole block against its
 * own BLENDED carrier (1 - a) exp(-j 2 pi n m / N_c) + a exp(-j 2 pi n (m+1) / N_c) -- SURVEY.md section 8 a', last
 * sentence; by linearity equal to the lerp of two CarrScores bins (batchcorrmanifold.cu:1950-1958) to rounding.
 * Needs DPE_FLAG_BRUTE_VEL.  In dpe_epoch_run / dpe_epoch_submit: with_vel = 2.                                   */
int dpe_score_vel_brute(dpe_ctx* ctx, void* stream);

/* dpe_result_fetch: D2H of the result block; synchronises the stream.           */
int dpe_result_fetch(dpe_ctx* ctx, dpe_result* out, void* stream);

/* dpe_epoch_run: the reference-facing one-call epoch with HOST buffers -- block
 * H2D, parameter H2D, prepare, correlogram, score, estimate, result D2H.  This is
 * what BatchCorrScores::Update + BatchCorrManifold::Update do back to back
 * (batchcorrscores.cu:975-1208, batchcorrmanifold.cu:2501-2635).  Synchronises. */
int dpe_epoch_run(dpe_ctx* ctx, const int16_t* iq_host, const dpe_epoch* ep,
                  const double* sat_states, int score_mode, int est_mode, int with_vel,
                  dpe_result* out, void* stream);

/* ---- asynchronous epochs ---------------------------------------------------------
 * dpe_epoch_submit enqueues one whole epoch (parameter + block upload, pre-pass, correlogram,
 * scoring, estimate, optional velocity manifold, result D2H into page-locked memory) on the
 * context's OWN stream and returns at once; dpe_epoch_collect waits for that epoch and hands the
 * result over.  One epoch may be in flight per context.  A caller that holds two contexts and
 * alternates between them gets cross-epoch overlap: every kernel except k_brute is sized to share
 * an SM with a running k_brute CTA, so the pre-pass / pair sort / reductions of one epoch run
 * under the k_brute of the other (the reference overlaps its BCS streams the same way,
 * batchcorrscores.cu:727-745,1046-1180).  `iq` may be a host or a device pointer; on ranks != 0
 * of a communicator (below) it is ignored and may be NULL.                                   */
int dpe_epoch_submit(dpe_ctx* ctx, const int16_t* iq, const dpe_epoch* ep, const double* sat_states,
                     int score_mode, int est_mode, int with_vel);
int dpe_epoch_collect(dpe_ctx* ctx, dpe_result* out);
/* 1 while a submitted epoch has not been collected                                            */
int dpe_epoch_pending(dpe_ctx* ctx);

/* ---- multi-GPU ---------------------------------------------
 * The candidate grid is sharded in contiguous index ranges (cfg.grid_offset / G / G_total);
 * one context per GPU, one NCCL communicator per context, every collective on the context's
 * own stream between its kernels -- no host code in between:
 *     rank 0: H2D of one packet {block, channel + geometry parameters, satellite states}
 *     ncclBroadcast(packet)  ->  pre-pass, scoring of the shard  ->  ncclAllGather(16-double
 *     partial)  ->  k_finalize on every rank (lowest global index wins arg-max ties)
 * The reference is single-GPU (no cudaSetDevice anywhere); this replaces nothing of it and
 * extends BatchCorrManifold::Update (batchcorrmanifold.cu:2501-2635) over `nranks` GPUs.
 * NCCL is dlopen()ed (libnccl.so.2) on the first dpe_comm_* call; single-GPU use needs none.
 *
 * dpe_comm_get_unique_id: rank 0 obtains DPE_COMM_ID_BYTES bytes and ships them to the other
 * ranks by any means (torch.distributed store, a file, a pipe, shared memory of one process).
 * dpe_comm_init: collective over all ranks (one process per GPU, or one thread per GPU inside
 * one process -- dpe_console's NumGPUs); blocks until every rank has joined.
 * dpe_epoch_run_dist = dpe_epoch_submit + dpe_epoch_collect; every rank passes the same `ep`
 * (the launch geometry needs C on the host); the device-side parameters, the satellite states
 * and the block are rank 0's.                                                                 */
#define DPE_COMM_ID_BYTES 128
int dpe_comm_get_unique_id(void* id);
int dpe_comm_init(dpe_ctx* ctx, int nranks, int rank, const void* id);
int dpe_comm_destroy(dpe_ctx* ctx);
int dpe_comm_info(dpe_ctx* ctx, int* nranks, int* rank, int* nccl_version);
int dpe_epoch_run_dist(dpe_ctx* ctx, const int16_t* iq, const dpe_epoch* ep, const double* sat_states,
                       int score_mode, int est_mode, int with_vel, dpe_result* out);

/* ---- access / introspection ---------------------------------------------------*/
const void* dpe_dev_ptr(dpe_ctx* ctx, int which);
/* debug copies (synchronise): chip indices int16 [C][S]; per-channel flags     */
int dpe_debug_channel_flags(dpe_ctx* ctx, int32_t* idx_next, int32_t* no_flip, int C);
/* bins of candidate range [i0, i0+n): f_idx int64 [n][C], alpha double [n][C].  sat_mode may be OR-ed with
 * DPE_DEBUG_BINS_EXACT: evaluate the reference's FP64 chain (batchcorrmanifold.cu:1779-1791) literally instead
 * of the centre-relative form the scoring kernels use -- the two must agree bit for bit.   */
#define DPE_DEBUG_BINS_EXACT 16
int dpe_debug_bins(dpe_ctx* ctx, int64_t i0, int64_t n, int sat_mode, int64_t* f_idx,
                   double* alpha, void* stream);
/* synchronous D2H read of `nbytes` at byte `offset` of a dpe_dev_ptr buffer       */
int dpe_debug_read(dpe_ctx* ctx, int which, size_t offset, void* dst, size_t nbytes);
/* number of kernels launched by this context since creation                     */
int64_t dpe_launch_count(dpe_ctx* ctx);
/* the context's own stream (cudaStream_t as void*): what dpe_epoch_submit launches on          */
void* dpe_ctx_stream(dpe_ctx* ctx);
/* registers per thread / dynamic+static shared bytes of a kernel of this library, by name
 * ("k_brute", "k_prep_corr", ...): lets a bench report whether the side kernels fit beside k_brute */
int dpe_kernel_attr(const char* kernel, int* regs, int* smem_bytes, int* max_threads);

/* ---- runtime helpers for host code that does not link the CUDA runtime ------------
 * (the C++ flow mirror in navlab-dpe-sdr_b200/host/ links libdpe_b200.so only).
 * dpe_stream_*: the flow's stream (Flow::Start creates one, flow.cu:33).  dpe_host_alloc:
 * page-locked host memory for the sample ring (cudaMallocHost, sampleblock.cu:205).   */
int dpe_stream_create(void** stream);
int dpe_stream_destroy(void* stream);
int dpe_stream_sync(void* stream);
int dpe_host_alloc(void** ptr, size_t bytes);
int dpe_host_free(void* ptr);
/* the device side of SampleBlock's ring (cudaMalloc + the reader thread's cudaMemcpyAsync on its own stream,
 * sampleblock.cu:221,403): dpe_stream_create_on / dpe_device_alloc bind `device` for the call only           */
int dpe_stream_create_on(void** stream, int device);
int dpe_device_alloc(void** ptr, size_t bytes, int device);
int dpe_device_free(void* ptr, int device);
int dpe_copy_h2d(void* dst_device, const void* src_host, size_t bytes, void* stream, int device);
int dpe_device_count(void);

/* ---- per-stage device timing (bench.py's roofline.achieved) --------------------
 * When enabled, every stage is bracketed by cudaEvents on the stream it is
 * launched on.  dpe_profile_read synchronises, adds the elapsed milliseconds of
 * every bracket recorded since the last read to ms[stage] (DPE_N_STAGES entries),
 * the number of brackets to count[stage], and clears the record.                 */
enum {
    DPE_STAGE_PREPARE = 0,       /* k_prep_corr: unpack, wipe-off, replica, correlogram window, flip choice */
    DPE_STAGE_CORRELOGRAM = 1,   /* (nothing since round 2: the correlogram is finished inside k_prep_corr) */
    DPE_STAGE_LOOKUP = 2,        /* k_score_lookup (grid reduction fused, last CTA)  */
    DPE_STAGE_BRUTE_BINS = 3,    /* k_pair_bins + k_block_scan + k_scatter           */
    DPE_STAGE_BRUTE_CORR = 4,    /* k_brute (the north-star kernel) alone            */
    DPE_STAGE_BRUTE_SCORE = 5,   /* k_score_pairs (grid reduction fused, last CTA)   */
    DPE_STAGE_ESTIMATE = 6,      /* k_finalize                                       */
    DPE_STAGE_VELOCITY = 7,      /* velocity manifold: k_carr_partial + k_score_vel, or the brute-force chain */
    DPE_N_STAGES = 8
};
int dpe_profile_enable(dpe_ctx* ctx, int on);
int dpe_profile_read(dpe_ctx* ctx, double* ms, int64_t* count);
/* valid (candidate, PRN) pairs the last brute-force pass correlated (synchronises) */
int64_t dpe_brute_pairs(dpe_ctx* ctx);

/* ---- micro-benchmarks (roofline denominators measured in the same run) -------
 * fp32: dependent-free FFMA2 streams on every SM; returns achieved TFLOP/s.
 * hbm : device copy of `bytes`; returns GB/s (read+write).                       */
int dpe_microbench_fp32(int device, int use_ffma2, double* tflops);
int dpe_microbench_fp64(int device, double* tflops);      /* DFMA stream: the pipe that bounds the lookup-path scoring kernels */
int dpe_microbench_hbm(int device, size_t bytes, double* gbs);

#ifdef __cplusplus
}
#endif
#endif /* DPE_B200_H_ */
