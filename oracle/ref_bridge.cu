// oracle/ref_bridge.cu -- TEST INFRASTRUCTURE (never linked into the product).
//
// The reference-side binding of INTEGRATION.md section 2, compiled: this translation unit takes the
// place of the reference's batchcorrscores.cu and batchcorrmanifold.cu in the link of CUDARecv.  It
// #includes the reference's own class declarations (modules/inc/batchcorrscores.h,
// batchcorrmanifold.h -- where they lie, nothing copied) and defines the member functions of
// dsp::BatchCorrScores and dsp::BatchCorrManifold with bodies that contain nothing but calls into
// libdpe_b200.so's C ABI (include/dpe_b200.h).  Everything else -- DPInit, SampleBlock, cuChanMgr,
// cuEKF, DataLogger, Flow, DPEFlow -- is the reference's unmodified object code, so the channel
// parameters, satellite states and grid centre arrive exactly the way the reference publishes them:
// as CUDA_DEVICE ports (cuchanmgr.cu:973-990, 1136-1171).  They are packed on the device
// (dpe_epoch_set_device); zVal / RVal / PosScores / TimeGrid go back as CUDA_DEVICE ports
// (batchcorrmanifold.cu:2297-2300, 2398-2405) and the reference's cuEKF / cuChanMgr consume them.
//
// oracle/Makefile links this with oracle/ref_driver.cu (-DREF_BRIDGE) into oracle/_ref/ref_dpe_bridge;
// tests/test_bridge.py runs it on the golden files and compares with the pure-reference run.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <vector>
#include "batchcorrscores.h"
#include "batchcorrmanifold.h"
#include "../include/dpe_b200.h"
#include "../include/dpe_flow.h"

namespace {
// one flow per process in this harness: the context both modules share
dpe_ctx* g_ctx = NULL;
int g_num_fft = 0;
int g_score_mode = DPE_SCORE_LOOKUP;
double* g_time_grid_d = NULL;
bool g_have_vel = false;

#define BR_CALL(stmt)                                                                              \
    do {                                                                                           \
        if ((stmt) != DPE_OK) {                                                                    \
            std::cerr << "[" << ModuleName << "] " #stmt " failed: " << dpe_last_error() << std::endl; \
            return -1;                                                                             \
        }                                                                                          \
    } while (0)
}  // namespace

// ------------------------------------------------------------------------------------------------
// dsp::BatchCorrScores  (ports: batchcorrscores.cu:672-698)
// ------------------------------------------------------------------------------------------------
dsp::BatchCorrScores::BatchCorrScores() {
    ModuleName = "BatchCorrScores";
    AllocateInputs(12);
    AllocateOutputs(3);
    Started = 0;
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
    ConfigOutput(0, "CodeScores", UNDEFINED_t, VALUE_CMPX, CUDA_DEVICE, VECTORLENGTH_ANY, NULL, 0);
    ConfigOutput(1, "CarrScores", UNDEFINED_t, VALUE_CMPX, CUDA_DEVICE, VECTORLENGTH_ANY, NULL, 0);
    ConfigOutput(2, "NumFFTPoints", INT_t, VALUE, HOST, 1, NULL, 0);
}

dsp::BatchCorrScores::~BatchCorrScores() {
    if (Started) Stop();
    delete[] expectedInputs;
    delete[] inputs;
    delete[] outputs;
}

int dsp::BatchCorrScores::Start(void*) {
    if (Started) return 0;
    const double fs = *(double*)inputs[9]->Data, T = *(double*)inputs[10]->Data;
    const long long S = (long long)(fs * T + 0.5);
    long long p2 = 1;
    while (p2 < S) p2 <<= 1;
    g_num_fft = (int)(p2 * 8);                         // carrSTot, batchcorrscores.cu:761
    outputs[2].Data = &g_num_fft;
    outputs[2].VectorLength = 1;
    Started = 1;
    return 0;
}

int dsp::BatchCorrScores::Stop(void) {
    Started = 0;
    return 0;
}

int dsp::BatchCorrScores::Update(void* cuFlowStream) {
    if (!Started || !g_ctx) return -1;
    void* stream = (void*)*(cudaStream_t*)cuFlowStream;
    const int C = inputs[1]->VectorLength;
    dpe_epoch_dev ep;
    memset(&ep, 0, sizeof(ep));
    ep.C = C;
    ep.prn = (const uint8_t*)inputs[1]->Data;
    ep.rc_start = (const double*)inputs[2]->Data;
    ep.ri_start = (const double*)inputs[3]->Data;
    ep.fc = (const double*)inputs[4]->Data;
    ep.fi = (const double*)inputs[5]->Data;
    ep.cp_start = (const int32_t*)inputs[6]->Data;
    ep.cp_ref = (const int32_t*)inputs[7]->Data;
    ep.doppler_sign = (const int32_t*)inputs[8]->Data;
    const double fs = *(double*)inputs[9]->Data, T = *(double*)inputs[10]->Data;
    // SampleBlock repoints "Samples" at the next DEVICE block every epoch (sampleblock.cu:508): used in place
    BR_CALL(dpe_block_stage(g_ctx, (const int16_t*)inputs[0]->Data, (int64_t)(fs * T + 0.5), stream));
    BR_CALL(dpe_epoch_set_device(g_ctx, &ep, DPE_PART_CHANNELS, stream));
    BR_CALL(dpe_replica_prepare(g_ctx, stream));
    BR_CALL(dpe_correlogram(g_ctx, stream));
    outputs[0].Data = const_cast<void*>(dpe_dev_ptr(g_ctx, DPE_PTR_CODE_SCORES));
    outputs[1].Data = const_cast<void*>(dpe_dev_ptr(g_ctx, DPE_PTR_CARR_SCORES));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// dsp::BatchCorrManifold  (ports / params: batchcorrmanifold.cu:2247-2303)
// ------------------------------------------------------------------------------------------------
dsp::BatchCorrManifold::BatchCorrManifold() {
    ModuleName = "BatchCorrManifold";
    AllocateInputs(19);
    AllocateOutputs(4);
    Started = 0;
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
    ConfigExpectedInput(10, "DopplerSign", INT_t, VALUE, 1);
    ConfigExpectedInput(11, "NumFFTPoints", INT_t, VALUE, 1);
    ConfigExpectedInput(12, "ENU2ECEFMat", DOUBLE_t, VALUE, 9);
    ConfigExpectedInput(13, "SatStatesOld", DOUBLE_t, STATE, VECTORLENGTH_ANY);
    ConfigExpectedInput(14, "CodePhase", DOUBLE_t, VALUE, VECTORLENGTH_ANY);
    ConfigExpectedInput(15, "CarrierPhase", DOUBLE_t, VALUE, VECTORLENGTH_ANY);
    ConfigExpectedInput(16, "cpRefTOW", INT_t, VALUE, VECTORLENGTH_ANY);
    ConfigExpectedInput(17, "cpElapsedEnd", INT_t, VALUE, VECTORLENGTH_ANY);
    ConfigExpectedInput(18, "cpRef", INT_t, VALUE, VECTORLENGTH_ANY);
    InsertParam("PosGridDimSize", (void*)&posGridDimSizeParam, INT_t, sizeof(int), sizeof(int));
    InsertParam("VelGridDimSize", (void*)&velGridDimSizeParam, INT_t, sizeof(int), sizeof(int));
    InsertParam("GridDimSpacing", (void*)&gridDimSpacingParam, FLOAT_t, sizeof(float), sizeof(float));
    InsertParam("GridType", (void*)&gridTypeParam, INT_t, sizeof(dsp::utils::ManifoldGridTypes),
                sizeof(dsp::utils::ManifoldGridTypes));
    InsertParam("LPower", (void*)&LPower, INT_t, sizeof(int), sizeof(int));
    InsertParam("GridLogFileName", (void*)&Filename, CHAR_t, FilenameCapacity, 0);
    InsertParam("LoadPosGrid", (void*)&loadPosGrid, BOOL_t, sizeof(bool), sizeof(bool));
    InsertParam("LoadPosGridFilename", (void*)&loadPosGridFilename, CHAR_t, FilenameCapacity, 0);
    ConfigOutput(0, "zVal", DOUBLE_t, STATE, CUDA_DEVICE, VECTORLENGTH_ANY, NULL, 0);
    ConfigOutput(1, "RVal", DOUBLE_t, COVARIANCE, CUDA_DEVICE, VECTORLENGTH_ANY, NULL, 0);
    ConfigOutput(2, "TimeGrid", DOUBLE_t, VALUE, CUDA_DEVICE, VECTORLENGTH_ANY, NULL, 0);
    ConfigOutput(3, "PosScores", DOUBLE_t, GRID, CUDA_DEVICE, VECTORLENGTH_ANY, NULL, 0);
}

dsp::BatchCorrManifold::~BatchCorrManifold() {
    if (Started) Stop();
    delete[] inputs;
    delete[] outputs;
    delete[] expectedInputs;
}

int dsp::BatchCorrManifold::Start(void* cuFlowStream) {
    if (Started) return 0;
    void* stream = (void*)*(cudaStream_t*)cuFlowStream;
    const int n = posGridDimSizeParam;
    const int dims[4] = {n, n, n, n};
    const double sp[4] = {gridDimSpacingParam, gridDimSpacingParam, gridDimSpacingParam, gridDimSpacingParam};
    const long G = (long)n * n * n * n;
    std::vector<double> grid((size_t)G * 4), tgrid(n);
    // host-side grid generator / CSV reader of this repository (include/dpe_flow.h)
    if (dpe_host_make_grid(dims, sp, (int)gridTypeParam, grid.data(), G * 4) != G) return -1;
    for (int i = 0; i < n; ++i) tgrid[i] = grid[(size_t)i * 4 + 3];            // t is the fastest axis
    if (loadPosGrid && dpe_host_read_grid(loadPosGridFilename, grid.data(), G * 4) != G) {
        std::clog << "[" << ModuleName << "] Open loadGridFile failed" << std::endl;
        return -1;
    }
    const double fs = *(double*)inputs[7]->Data, T = *(double*)inputs[6]->Data;
    dpe_cfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.abi_version = DPE_ABI_VERSION;
    cfg.device = 0;
    cfg.fs = fs;
    cfg.S = (int64_t)(fs * T + 0.5);
    cfg.max_chan = DPE_MAX_CHAN;
    cfg.time_dim = n;
    cfg.G = cfg.G_total = G;
    cfg.lpower = LPower;
    double ext = 0;
    for (long i = 0; i < G; ++i)
        ext = std::max(ext, std::sqrt(grid[4 * i] * grid[4 * i] + grid[4 * i + 1] * grid[4 * i + 1] +
                                      grid[4 * i + 2] * grid[4 * i + 2]) + std::fabs(grid[4 * i + 3]));
    cfg.lag_halfwidth = std::min(160, std::max(4, (int)std::ceil(ext * fs / 299792458.0) + 3));
    const char* brute = getenv("DPE_BRIDGE_BRUTE");
    g_score_mode = (brute && brute[0] == '1') ? DPE_SCORE_BRUTE : DPE_SCORE_LOOKUP;
    cfg.flags = g_score_mode == DPE_SCORE_BRUTE ? DPE_FLAG_BRUTE_TILES : 0;
    std::vector<double> vgrid;
    const int nv = velGridDimSizeParam;
    if (nv > 0) {
        const int vd[4] = {nv, nv, nv, nv};
        const long Gv = (long)nv * nv * nv * nv;
        vgrid.resize((size_t)Gv * 4);
        if (dpe_host_make_grid(vd, sp, 0, vgrid.data(), Gv * 4) != Gv) return -1;   // BCM_InitVelGrid is uniform (:265-316)
        cfg.Gv = Gv;
        double vext = 0;
        for (long i = 0; i < Gv; ++i)
            vext = std::max(vext, std::sqrt(vgrid[4 * i] * vgrid[4 * i] + vgrid[4 * i + 1] * vgrid[4 * i + 1] +
                                            vgrid[4 * i + 2] * vgrid[4 * i + 2]) + std::fabs(vgrid[4 * i + 3]));
        long long nfft = 1;
        while (nfft < cfg.S) nfft <<= 1;
        nfft *= 8;
        cfg.dopp_halfwidth = std::min(4096, std::max(4, (int)std::ceil(vext * 1.57542e9 / 299792458.0 * (double)nfft / fs) + 3));
    }
    if (g_ctx) { dpe_ctx_destroy(g_ctx); g_ctx = NULL; }
    BR_CALL(dpe_ctx_create(&g_ctx, &cfg));
    BR_CALL(dpe_grid_set(g_ctx, grid.data(), G, stream));
    if (nv > 0) BR_CALL(dpe_vel_grid_set(g_ctx, vgrid.data(), cfg.Gv, stream));
    g_have_vel = nv > 0;
    BR_CALL(dpe_stream_sync(stream));
    // "TimeGrid" is a CUDA_DEVICE port in the reference (cuChanMgr's CHM_GridPrep reads it on the device)
    if (cudaMalloc((void**)&g_time_grid_d, sizeof(double) * n) != cudaSuccess) return -1;
    cudaMemcpy(g_time_grid_d, tgrid.data(), sizeof(double) * n, cudaMemcpyHostToDevice);
    outputs[0].Data = const_cast<void*>(dpe_dev_ptr(g_ctx, DPE_PTR_ZVAL));
    outputs[0].VectorLength = 8;
    outputs[1].Data = const_cast<void*>(dpe_dev_ptr(g_ctx, DPE_PTR_RVAL));
    outputs[1].VectorLength = 64;
    outputs[2].Data = g_time_grid_d;
    outputs[2].VectorLength = (unsigned short)n;
    outputs[3].Data = const_cast<void*>(dpe_dev_ptr(g_ctx, DPE_PTR_POS_SCORES));
    outputs[3].VectorLength = (unsigned short)G;
    // the identity the reference leaves in RVal (rows are rewritten every epoch by the estimate kernels)
    Started = 1;
    return 0;
}

int dsp::BatchCorrManifold::Stop(void) {
    if (g_ctx) { dpe_ctx_destroy(g_ctx); g_ctx = NULL; }
    if (g_time_grid_d) { cudaFree(g_time_grid_d); g_time_grid_d = NULL; }
    Started = 0;
    return 0;
}

int dsp::BatchCorrManifold::Update(void* cuFlowStream) {
    if (!Started || !g_ctx) return -1;
    void* stream = (void*)*(cudaStream_t*)cuFlowStream;
    dpe_epoch_dev ep;
    memset(&ep, 0, sizeof(ep));
    ep.C = inputs[8]->VectorLength;
    ep.fc = (const double*)inputs[8]->Data;
    ep.rc_end = (const double*)inputs[14]->Data;
    ep.cp_ref_tow = (const int32_t*)inputs[16]->Data;
    ep.cp_end = (const int32_t*)inputs[17]->Data;
    ep.cp_ref = (const int32_t*)inputs[18]->Data;
    ep.center = (const double*)inputs[2]->Data;
    ep.enu2ecef = (const double*)inputs[12]->Data;
    ep.sat_states = (const double*)inputs[4]->Data;
    ep.rx_time = *(const double*)inputs[5]->Data;           // HOST port (cuchanmgr.cu:973)
    BR_CALL(dpe_epoch_set_device(g_ctx, &ep, DPE_PART_GEOMETRY, stream));
    BR_CALL(dpe_score_pos(g_ctx, g_score_mode, DPE_SAT_MIDDLE, stream));
    BR_CALL(dpe_estimate(g_ctx, DPE_EST_ARGMAX, NULL, 1, stream));
    if (g_have_vel) BR_CALL(dpe_score_vel(g_ctx, stream));
    // the reference ends its Update with cudaStreamSynchronize on its streams (batchcorrmanifold.cu:2620-2626)
    BR_CALL(dpe_stream_sync(stream));
    return 0;
}
