// modules_dpe.cpp -- BatchCorrScores and BatchCorrManifold: the two hot-path modules.  Their
// Update() bodies gather the same input ports as the reference's modules and make C-ABI calls;
// no device code, no CUDA runtime here.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <iostream>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <condition_variable>
#include "modules.h"

namespace dsp {

static std::mutex g_mu;
static std::map<void*, SharedCtx*> g_shared;

// ------------------------------------------------------------------------------------------
// RankPool: ranks 1..N-1 of a multi-GPU flow, one host thread per GPU (NCCL collectives of one
// communicator must be issued by concurrent callers; the flow thread is rank 0).  Each worker owns a
// context holding its contiguous shard of the grid and runs dpe_epoch_run_dist once per epoch.
// ------------------------------------------------------------------------------------------
struct RankPool {
    struct Job { const dpe_epoch* ep = nullptr; const double* sat = nullptr; int score = 0, est = 0; };
    std::vector<dpe_ctx*> ctx;            // [r], r = 0 is owned by SharedCtx
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv;
    long epoch = 0;
    int pending = 0, failed = 0;
    bool quit = false;
    Job job;
    std::string error;

    void Worker(int r, dpe_cfg cfg, std::vector<double> shard, int nranks, std::vector<unsigned char> id) {
        bool ok = dpe_ctx_create(&ctx[r], &cfg) == DPE_OK &&
                  dpe_grid_set(ctx[r], shard.data(), cfg.G, dpe_ctx_stream(ctx[r])) == DPE_OK &&
                  dpe_comm_init(ctx[r], nranks, r, id.data()) == DPE_OK;
        std::vector<double>().swap(shard);
        {
            std::lock_guard<std::mutex> lk(mu);
            if (!ok) { ++failed; error = dpe_last_error(); }
            --pending;
        }
        cv.notify_all();
        long seen = 0;
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return quit || epoch != seen; });
                if (quit) break;
                seen = epoch;
                j = job;
            }
            dpe_result res;
            const int rc = ok ? dpe_epoch_run_dist(ctx[r], nullptr, j.ep, j.sat, j.score, j.est, 0, &res) : DPE_ESTATE;
            {
                std::lock_guard<std::mutex> lk(mu);
                if (rc) { ++failed; error = dpe_last_error(); }
                --pending;
            }
            cv.notify_all();
        }
        if (ctx[r]) { dpe_ctx_destroy(ctx[r]); ctx[r] = nullptr; }
    }
    // rank 0's side of one epoch: wake the workers, run the root's call, wait for everybody
    int Run(dpe_ctx* root, const int16_t* iq, const dpe_epoch* ep, const double* sat, int score, int est, int with_vel,
            dpe_result* out) {
        {
            std::lock_guard<std::mutex> lk(mu);
            job.ep = ep; job.sat = sat; job.score = score; job.est = est;
            pending = (int)workers.size();
            ++epoch;
        }
        cv.notify_all();
        const int rc = dpe_epoch_run_dist(root, iq, ep, sat, score, est, with_vel, out);
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return pending == 0; });
        return (rc || failed) ? -1 : 0;
    }
    void Shutdown() {
        { std::lock_guard<std::mutex> lk(mu); quit = true; }
        cv.notify_all();
        for (size_t i = 0; i < workers.size(); ++i) if (workers[i].joinable()) workers[i].join();
        workers.clear();
    }
};

SharedCtx* SharedFor(void* key) {
    std::lock_guard<std::mutex> lk(g_mu);
    SharedCtx*& p = g_shared[key];
    if (!p) { p = new SharedCtx(); std::memset(&p->ep, 0, sizeof(p->ep)); }
    return p;
}

void SharedRelease(void* key) {
    std::lock_guard<std::mutex> lk(g_mu);
    std::map<void*, SharedCtx*>::iterator it = g_shared.find(key);
    if (it == g_shared.end()) return;
    if (it->second->pool) { it->second->pool->Shutdown(); delete it->second->pool; }
    if (it->second->ctx) dpe_ctx_destroy(it->second->ctx);
    delete it->second;
    g_shared.erase(it);
}

#define DPE_CALL(stmt)                                                                          \
    do {                                                                                        \
        if ((stmt) != DPE_OK) {                                                                 \
            std::cerr << "[" << ModuleName << "] " #stmt " failed: " << dpe_last_error() << std::endl; \
            return -1;                                                                          \
        }                                                                                       \
    } while (0)

// ------------------------------------------------------------------------------------------
// BatchCorrScores (input table: cudarecv/modules/src/batchcorrscores.cu:681-698)
// ------------------------------------------------------------------------------------------
BatchCorrScores::BatchCorrScores() {
    ModuleName = "BatchCorrScores";
    AllocateInputs(11);
    ConfigExpectedInput(0, "Samples", UNDEFINED_t, VALUE_CMPX, VECTORLENGTH_ANY);
    ConfigExpectedInput(1, "ValidPRNs", CHAR_t, VALUE, VECTORLENGTH_ANY);
    ConfigExpectedInput(2, "CodePhaseStart", DOUBLE_t, VALUE, VECTORLENGTH_ANY);
    ConfigExpectedInput(3, "CarrierPhaseStart", DOUBLE_t, VALUE, VECTORLENGTH_ANY);
    ConfigExpectedInput(4, "CodeFrequency", DOUBLE_t, FREQUENCY_HZ, VECTORLENGTH_ANY);
    ConfigExpectedInput(5, "CarrierFrequency", DOUBLE_t, FREQUENCY_HZ, VECTORLENGTH_ANY);
    ConfigExpectedInput(6, "cpElapsedStart", INT_t, VALUE, VECTORLENGTH_ANY);
    ConfigExpectedInput(7, "cpReference", INT_t, VALUE, VECTORLENGTH_ANY);
    ConfigExpectedInput(8, "DopplerSign", INT_t, VALUE, VECTORLENGTH_ANY);
    ConfigExpectedInput(9, "SamplingFrequency", DOUBLE_t, FREQUENCY_HZ, 1);
    ConfigExpectedInput(10, "SampleLength", DOUBLE_t, VALUE, 1);
    AllocateOutputs(3);
    ConfigOutput(0, "CodeScores", UNDEFINED_t, VALUE_CMPX, CUDA_DEVICE, VECTORLENGTH_ANY, nullptr, 0);
    ConfigOutput(1, "CarrScores", UNDEFINED_t, VALUE_CMPX, CUDA_DEVICE, VECTORLENGTH_ANY, nullptr, 0);
    ConfigOutput(2, "NumFFTPoints", INT_t, VALUE, HOST, 1, &numFFTPoints, 0);
}

int BatchCorrScores::Start(void*) {
    if (Started) return 0;
    if (!InputsConnected()) return -1;
    const double fs = *In<double>(9), T = *In<double>(10);
    const int64_t S = (int64_t)(fs * T + 0.5);
    int64_t p2 = 1;
    while (p2 < S) p2 <<= 1;
    numFFTPoints = (int)(p2 * 8);                  // carrSTot, batchcorrscores.cu:761
    Started = true;
    return 0;
}

int BatchCorrScores::Update(void* cuFlowStream) {
    if (!Started) return -1;
    SharedCtx* sh = SharedFor(cuFlowStream);
    if (!sh->ctx) { std::cerr << "[" << ModuleName << "] no context (BatchCorrManifold not started)" << std::endl; return -1; }
    void* stream = StreamOf(cuFlowStream);
    const int C = (int)InLen(1);
    dpe_epoch& ep = sh->ep;
    ep.C = C;
    ep.doppler_sign = In<int>(8)[0];
    for (int i = 0; i < C; ++i) {
        ep.prn[i] = (uint8_t)In<char>(1)[i];
        ep.rc_start[i] = In<double>(2)[i];
        ep.ri_start[i] = In<double>(3)[i];
        ep.fc[i] = In<double>(4)[i];
        ep.fi[i] = In<double>(5)[i];
        ep.cp_start[i] = In<int>(6)[i];
        ep.cp_ref[i] = In<int>(7)[i];
    }
    if (sh->pool) {                                // multi-GPU: the epoch runs as one call in BatchCorrManifold::Update
        sh->iq = In<int16_t>(0);
        sh->iq_len = InLen(0);
        UpdateOutput(0, C, const_cast<void*>(dpe_dev_ptr(sh->ctx, DPE_PTR_CODE_SCORES)), 0);
        return 0;
    }
    DPE_CALL(dpe_block_stage(sh->ctx, In<int16_t>(0), InLen(0), stream));
    DPE_CALL(dpe_epoch_set_part(sh->ctx, &ep, nullptr, DPE_PART_CHANNELS, stream));
    DPE_CALL(dpe_replica_prepare(sh->ctx, stream));
    DPE_CALL(dpe_correlogram(sh->ctx, stream));
    // the window of CodeScores a position grid can reach: [C][2W+2] complex doubles on the device
    UpdateOutput(0, C, const_cast<void*>(dpe_dev_ptr(sh->ctx, DPE_PTR_CODE_SCORES)), 0);
    return 0;
}

// ------------------------------------------------------------------------------------------
// BatchCorrManifold (input table: cudarecv/modules/src/batchcorrmanifold.cu:2261-2300)
// ------------------------------------------------------------------------------------------
BatchCorrManifold::BatchCorrManifold() {
    ModuleName = "BatchCorrManifold";
    AllocateInputs(19);
    ConfigExpectedInput(0, "CodeScores", UNDEFINED_t, VALUE_CMPX, VECTORLENGTH_ANY);
    ConfigExpectedInput(1, "CarrScores", UNDEFINED_t, VALUE_CMPX, VECTORLENGTH_ANY);
    ConfigExpectedInput(2, "xCurrkk1", DOUBLE_t, STATE, VECTORLENGTH_ANY);
    ConfigExpectedInput(3, "txTime", DOUBLE_t, VALUE, VECTORLENGTH_ANY);
    ConfigExpectedInput(4, "SatStates", DOUBLE_t, STATE, VECTORLENGTH_ANY);
    ConfigExpectedInput(5, "rxTime", DOUBLE_t, VALUE, 1);
    ConfigExpectedInput(6, "SampleLength", DOUBLE_t, VALUE, 1);
    ConfigExpectedInput(7, "SamplingFrequency", DOUBLE_t, FREQUENCY_HZ, 1);
    ConfigExpectedInput(8, "CodeFrequency", DOUBLE_t, FREQUENCY_HZ, VECTORLENGTH_ANY);
    ConfigExpectedInput(9, "CarrierFrequency", DOUBLE_t, FREQUENCY_HZ, VECTORLENGTH_ANY);
    ConfigExpectedInput(10, "DopplerSign", INT_t, VALUE, VECTORLENGTH_ANY);
    ConfigExpectedInput(11, "NumFFTPoints", INT_t, VALUE, 1);
    ConfigExpectedInput(12, "ENU2ECEFMat", DOUBLE_t, VALUE, 9);
    ConfigExpectedInput(13, "SatStatesOld", DOUBLE_t, STATE, VECTORLENGTH_ANY);
    ConfigExpectedInput(14, "CodePhase", DOUBLE_t, VALUE, VECTORLENGTH_ANY);
    ConfigExpectedInput(15, "CarrierPhase", DOUBLE_t, VALUE, VECTORLENGTH_ANY);
    ConfigExpectedInput(16, "cpRefTOW", INT_t, VALUE, VECTORLENGTH_ANY);
    ConfigExpectedInput(17, "cpElapsedEnd", INT_t, VALUE, VECTORLENGTH_ANY);
    ConfigExpectedInput(18, "cpRef", INT_t, VALUE, VECTORLENGTH_ANY);
    InsertParam("PosGridDimSize", &posGridDimSize, INT_t, sizeof(int), sizeof(int));
    InsertParam("VelGridDimSize", &velGridDimSize, INT_t, sizeof(int), sizeof(int));
    InsertParam("GridDimSpacing", &gridDimSpacing, FLOAT_t, sizeof(float), sizeof(float));
    InsertParam("GridType", &gridType, INT_t, sizeof(int), sizeof(int));
    InsertParam("LPower", &LPower, INT_t, sizeof(int), sizeof(int));
    InsertParam("GridLogFileName", Filename, CHAR_t, kNameCap, 0);
    InsertParam("LoadPosGrid", &loadPosGrid, BOOL_t, sizeof(bool), sizeof(bool));
    InsertParam("LoadPosGridFilename", loadPosGridFilename, CHAR_t, kNameCap, 0);
    InsertParam("BruteForce", &bruteForce, BOOL_t, sizeof(bool), sizeof(bool));          // north-star kernel
    InsertParam("WeightedMean", &weightedMean, BOOL_t, sizeof(bool), sizeof(bool));      // dormant Method 1
    InsertParam("LagHalfwidth", &lagHalfwidth, INT_t, sizeof(int), sizeof(int));
    InsertParam("DopplerHalfwidth", &doppHalfwidth, INT_t, sizeof(int), sizeof(int));
    InsertParam("NumGPUs", &numGPUs, INT_t, sizeof(int), sizeof(int));                    // grid sharded over GPUs
    InsertParam("Device", &device, INT_t, sizeof(int), sizeof(int));                      // first CUDA ordinal
    AllocateOutputs(4);
    ConfigOutput(0, "zVal", DOUBLE_t, STATE, HOST, 8, zVal, 0);
    ConfigOutput(1, "RVal", DOUBLE_t, COVARIANCE, HOST, 64, RVal, 0);
    ConfigOutput(2, "TimeGrid", DOUBLE_t, VALUE, HOST, VECTORLENGTH_ANY, nullptr, 0);
    ConfigOutput(3, "PosScores", DOUBLE_t, GRID, CUDA_DEVICE, VECTORLENGTH_ANY, nullptr, 0);
    std::memset(zVal, 0, sizeof(zVal));
    for (int i = 0; i < 64; ++i) RVal[i] = (i % 9 == 0) ? 1.0 : 0.0;
}

int BatchCorrManifold::Start(void* cuFlowStream) {
    if (Started) return 0;
    if (!InputsConnected()) return -1;
    const int dims[4] = {posGridDimSize, posGridDimSize, posGridDimSize, posGridDimSize};
    const double sp[4] = {gridDimSpacing, gridDimSpacing, gridDimSpacing, gridDimSpacing};
    if (gridType != 0 && gridType != 2) { std::clog << "[" << ModuleName << "] unsupported manifold type" << std::endl; return -1; }
    gnss::MakeGrid(dims, sp, gridType, &grid, &timeGrid);     // timeGrid always comes from the generated grid
    if (loadPosGrid) {                                        // (batchcorrmanifold.cu:2416-2448)
        std::vector<double> g;
        if (gnss::ReadGridCsv(loadPosGridFilename, &g)) {
            std::clog << "[" << ModuleName << "] Open loadGridFile failed: " << loadPosGridFilename << std::endl;
            return -1;
        }
        if (g.size() != grid.size()) {
            std::clog << "[" << ModuleName << "] grid file holds " << g.size() / 4 << " points, PosGridDimSize^4 = "
                      << grid.size() / 4 << std::endl;
            return -1;
        }
        grid.swap(g);
    }
    const double fs = *In<double>(7), T = *In<double>(6);
    dpe_cfg cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.abi_version = DPE_ABI_VERSION;
    const int ndev = dpe_device_count();
    if (numGPUs < 1 || device < 0 || device + numGPUs > ndev) {
        std::cerr << "[" << ModuleName << "] Device " << device << " + NumGPUs " << numGPUs << " exceeds the "
                  << ndev << " visible GPU(s)" << std::endl;
        return -1;
    }
    cfg.device = device;
    cfg.fs = fs;
    cfg.S = (int64_t)(fs * T + 0.5);
    cfg.max_chan = DPE_MAX_CHAN;
    cfg.time_dim = posGridDimSize;
    cfg.G = cfg.G_total = (int64_t)grid.size() / 4;
    cfg.lpower = LPower;
    // lag window: a candidate at ENU offset p and clock offset t moves the code phase by at most
    // (|p| + |t|) / c seconds = (|p| + |t|) fs / c samples (+3 for the lerp neighbour and tracking residuals)
    int W = lagHalfwidth;
    if (W <= 0) {
        double ext = 0;
        for (size_t i = 0; i + 3 < grid.size(); i += 4)
            ext = std::max(ext, std::sqrt(grid[i] * grid[i] + grid[i + 1] * grid[i + 1] + grid[i + 2] * grid[i + 2]) +
                                    std::fabs(grid[i + 3]));
        W = std::min(160, std::max(4, (int)std::ceil(ext * fs / gnss::kC) + 3));
    }
    cfg.lag_halfwidth = W;
    cfg.flags = bruteForce ? DPE_FLAG_BRUTE_TILES : 0;
    // velocity / drift manifold: BCM_InitVelGrid (batchcorrmanifold.cu:265-316) is uniform for every grid type
    std::vector<double> vgrid;
    if (velGridDimSize > 0) {
        const int vd[4] = {velGridDimSize, velGridDimSize, velGridDimSize, velGridDimSize};
        gnss::MakeGrid(vd, sp, 0, &vgrid, nullptr);
        cfg.Gv = (int64_t)vgrid.size() / 4;
        // Doppler window: velocity v and drift d move the carrier by at most (|v| + |d|) F_L1 / c Hz;
        // bins of the zero-padded spectrum are fs / N_c wide (+3 bins for the lerp neighbour and residuals)
        int Wd = doppHalfwidth;
        if (Wd <= 0) {
            double ext = 0;
            for (size_t i = 0; i + 3 < vgrid.size(); i += 4)
                ext = std::max(ext, std::sqrt(vgrid[i] * vgrid[i] + vgrid[i + 1] * vgrid[i + 1] + vgrid[i + 2] * vgrid[i + 2]) +
                                        std::fabs(vgrid[i + 3]));
            int64_t nfft = 1;
            while (nfft < (int64_t)(fs * T + 0.5)) nfft <<= 1;
            nfft *= 8;
            Wd = std::min(4096, std::max(4, (int)std::ceil(ext * gnss::kFL1 / gnss::kC * (double)nfft / fs) + 3));
        }
        cfg.dopp_halfwidth = Wd;
    }
    SharedCtx* sh = SharedFor(cuFlowStream);
    flowStream = cuFlowStream;                                      // from here on Stop() has something to release
    if (sh->pool) { sh->pool->Shutdown(); delete sh->pool; sh->pool = nullptr; }
    if (sh->ctx) { dpe_ctx_destroy(sh->ctx); sh->ctx = nullptr; }
    const int64_t G_total = cfg.G_total;
    const int64_t per = (G_total + numGPUs - 1) / numGPUs;          // contiguous index ranges (sharding.py::shard_range)
    if ((int64_t)(numGPUs - 1) * per >= G_total && numGPUs > 1) {   // checked before any rank thread exists: a rank that
        std::cerr << "[" << ModuleName << "] more GPUs than grid points" << std::endl;   // never joins would leave the others in the communicator's rendezvous
        return -1;
    }
    cfg.G = std::min(per, G_total);
    cfg.grid_offset = 0;
    DPE_CALL(dpe_ctx_create(&sh->ctx, &cfg));
    sh->score_mode = bruteForce ? DPE_SCORE_BRUTE : DPE_SCORE_LOOKUP;
    sh->est_mode = weightedMean ? DPE_EST_WEIGHTED : DPE_EST_ARGMAX;
    void* stream = StreamOf(cuFlowStream);
    if (numGPUs > 1) {
        unsigned char id[DPE_COMM_ID_BYTES];
        DPE_CALL(dpe_comm_get_unique_id(id));
        RankPool* pool = new RankPool();
        sh->pool = pool;
        pool->ctx.assign(numGPUs, nullptr);
        pool->ctx[0] = sh->ctx;
        pool->pending = numGPUs - 1;
        for (int r = 1; r < numGPUs; ++r) {
            dpe_cfg rc = cfg;
            rc.device = device + r;
            rc.grid_offset = std::min((int64_t)r * per, G_total);
            rc.G = std::min(per, G_total - rc.grid_offset);
            rc.Gv = 0;                                              // the velocity manifold runs on rank 0 only
            std::vector<double> shard(grid.begin() + 4 * rc.grid_offset, grid.begin() + 4 * (rc.grid_offset + rc.G));
            pool->workers.emplace_back(&RankPool::Worker, pool, r, rc, std::move(shard), numGPUs,
                                       std::vector<unsigned char>(id, id + DPE_COMM_ID_BYTES));
        }
        DPE_CALL(dpe_comm_init(sh->ctx, numGPUs, 0, id));           // collective: returns once every rank has joined
        std::unique_lock<std::mutex> lk(pool->mu);
        pool->cv.wait(lk, [&] { return pool->pending == 0; });
        if (pool->failed) {
            std::cerr << "[" << ModuleName << "] a GPU rank failed to start: " << pool->error << std::endl;
            return -1;
        }
        std::clog << "[" << ModuleName << "] grid sharded over " << numGPUs << " GPUs, " << per << " candidates each" << std::endl;
    }
    DPE_CALL(dpe_fold_estimate(sh->ctx, numGPUs > 1 ? -1 : sh->est_mode));   // one GPU: the estimate is the scoring kernel's tail
    DPE_CALL(dpe_grid_set(sh->ctx, grid.data(), cfg.G, stream));
    if (cfg.Gv > 0) DPE_CALL(dpe_vel_grid_set(sh->ctx, vgrid.data(), cfg.Gv, stream));
    haveVel = cfg.Gv > 0;
    DPE_CALL(dpe_stream_sync(stream));
    UpdateOutput(2, (int64_t)timeGrid.size(), timeGrid.data(), 0);
    UpdateOutput(3, cfg.G, const_cast<void*>(dpe_dev_ptr(sh->ctx, DPE_PTR_POS_SCORES)), 0);   // rank 0's shard
    Started = true;
    return 0;
}

int BatchCorrManifold::Update(void* cuFlowStream) {
    if (!Started) return -1;
    SharedCtx* sh = SharedFor(cuFlowStream);
    void* stream = StreamOf(cuFlowStream);
    dpe_epoch& ep = sh->ep;
    const int C = (int)InLen(8);
    if (C != ep.C) { std::cerr << "[" << ModuleName << "] channel count differs from BatchCorrScores" << std::endl; return -1; }
    for (int i = 0; i < C; ++i) {
        ep.fc[i] = In<double>(8)[i];
        ep.rc_end[i] = In<double>(14)[i];
        ep.cp_ref_tow[i] = In<int>(16)[i];
        ep.cp_end[i] = In<int>(17)[i];
        ep.cp_ref[i] = In<int>(18)[i];
    }
    ep.rx_time = *In<double>(5);
    for (int i = 0; i < 8; ++i) ep.center[i] = In<double>(2)[i];
    for (int i = 0; i < 9; ++i) ep.enu2ecef[i] = In<double>(12)[i];
    if (sh->pool) {
        // one dpe_epoch_run_dist per GPU: rank 0 uploads the packet, NCCL broadcasts it, every rank scores its
        // shard, the partials are all-gathered and reduced on every rank (include/dpe_b200.h, "multi-GPU")
        if (sh->pool->Run(sh->ctx, sh->iq, &ep, In<double>(4), sh->score_mode, sh->est_mode, haveVel ? 1 : 0, &last)) {
            std::cerr << "[" << ModuleName << "] multi-GPU epoch failed: " << dpe_last_error() << " " << sh->pool->error << std::endl;
            return -1;
        }
    } else {
        DPE_CALL(dpe_epoch_set_part(sh->ctx, &ep, In<double>(4), DPE_PART_GEOMETRY, stream));
        const int sat_mode = (sh->est_mode == DPE_EST_WEIGHTED) ? DPE_SAT_PER_TIME : DPE_SAT_MIDDLE;
        DPE_CALL(dpe_score_pos(sh->ctx, sh->score_mode, sat_mode, stream));
        DPE_CALL(dpe_estimate(sh->ctx, sh->est_mode, nullptr, 1, stream));
        if (haveVel) DPE_CALL(dpe_score_vel(sh->ctx, stream));
        DPE_CALL(dpe_result_fetch(sh->ctx, &last, stream));
    }
    if (last.out_of_window)
        std::clog << "[" << ModuleName << "] " << last.out_of_window << " candidate-PRN pairs outside the lag window"
                  << std::endl;
    for (int i = 0; i < 4; ++i) zVal[i] = last.z[i];
    if (last.vel_out_of_window)
        std::clog << "[" << ModuleName << "] " << last.vel_out_of_window
                  << " velocity candidate-PRN pairs outside the Doppler window" << std::endl;
    // velocity / drift half: the velocity manifold's arg-max, or (VelGridDimSize = 0) the prediction passed through
    for (int i = 4; i < 8; ++i) zVal[i] = haveVel ? last.z[i] : ep.center[i];
    return 0;
}

int BatchCorrManifold::Stop() {
    if (flowStream) SharedRelease(flowStream);                      // also after a Start() that failed half way
    flowStream = nullptr;
    Started = false;
    return 0;
}

}  // namespace dsp
