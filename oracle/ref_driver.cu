// oracle/ref_driver.cu -- TEST INFRASTRUCTURE (never linked into the product).
//
// Harness around the UNMODIFIED reference (CUDARecv, compiled from
// /root/reference/cudarecv by oracle/Makefile): it derives from dsp::DPEFlow,
// lets the reference wire its own seven modules (dpeflow.cpp:26-222), overrides
// the file / grid parameters through the reference's public Flow::SetModParam
// (flow.h:42-47), then steps the modules itself -- the same
// `for m in Mods: m->Update(&cuStream)` loop as Flow::FlowThread (flow.cu:122-137),
// without the SCHED_RR thread -- and dumps, for every epoch, the reference's own
// device buffers through Flow::GetOutput:
//   before the epoch : channel parameters + satellite states (cuChanMgr outputs),
//                      grid centre xCurrkk1 (cuEKF)
//   after BCM        : CodeScores window (BatchCorrScores), PosScores, zVal
//   after the epoch  : xCurrk1k1
// Files: <out>/e<epoch>_<name>.bin (raw little-endian), <out>/meta.txt.
// Also prints the mean wall time per epoch (the reference kernels on this GPU).
//
// usage: ref_dpe <samples.dat> <handoff.csv> <rinex.n> <grid.csv> <pos_dim> <vel_dim>
//                <epochs> <out_dir> [lag_halfwidth=32] [fs=2.5e6]
//
// Built a second time with -DREF_WEIGHTED (oracle/_ref/ref_dpe_weighted): this translation unit then
// takes the place of batchcorrmanifold.o by #including the reference's batchcorrmanifold.cu where it
// lies, so that its DORMANT score-weighted estimator -- BCM_PosMeasReduction (:816-1056) +
// BCM_ReduceAndPosMeas (:1365-1510), launches commented out at :2547-2567 -- can be launched on the
// module's own device buffers after every BatchCorrManifold::Update, with exactly the arguments of the
// commented launch.  Dumps e<epoch>_zval_weighted.bin (8 doubles: position-clock, velocity-drift) and
// e<epoch>_weighted_parts.bin / e<epoch>_weighted_vel_parts.bin (8 x weightState_t {a,b,c,d,score} each); the velocity
// twins BCM_VelMeasReduction (:1090-1347) + BCM_ReduceAndVelMeas (:1525-) are launched the same way.
#ifdef REF_WEIGHTED
#include "batchcorrmanifold.cu"
#endif
#include <cuda_runtime.h>
#include <cufft.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "dsp.h"
#include "flow.h"
#include "dpeflow.h"

static void dump_host(const std::string& dir, int epoch, const char* name, const void* p, size_t bytes) {
    char path[512];
    snprintf(path, sizeof(path), "%s/e%03d_%s.bin", dir.c_str(), epoch, name);
    FILE* f = fopen(path, "wb");
    if (!f) { perror(path); exit(3); }
    fwrite(p, 1, bytes, f);
    fclose(f);
}

#ifdef REF_WEIGHTED
// adds no data: only a door to the protected members of the module the reference flow created
class BcmProbe : public dsp::BatchCorrManifold {
  public:
    int RunWeighted(double z[8], double parts[40], double vparts[40]) {
        double *z_d = NULL, *R_d = NULL;
        if (cudaMalloc((void**)&z_d, 16 * sizeof(double)) != cudaSuccess) return -1;
        if (cudaMalloc((void**)&R_d, 64 * sizeof(double)) != cudaSuccess) return -1;
        cudaMemset(z_d, 0, 16 * sizeof(double));
        // batchcorrmanifold.cu:2548-2553 and :2562-2563, verbatim arguments
        BCM_PosMeasReduction<<<threadsPerMeasRedBlock, threadsPerPosManiBlock,
                               sizeof(dsp::utils::weightState_t<double>) * threadsPerPosManiBlock, posStream>>>(
            satStates_d, codeScores_d, xCurr_d, gridPosLocs_d, enu2ecefMat_d, codeFrequency_d, txTimePtr_d, *rxTimePtr,
            numChan, currSamplingFreq, numSamps, weightedPosStates_d);
        BCM_ReduceAndPosMeas<<<1, threadsPerMeasRedBlock,
                               sizeof(dsp::utils::weightState_t<double>) * threadsPerMeasRedBlock, posStream>>>(
            weightedPosStates_d, threadsPerMeasRedBlock, z_d, R_d);
        if (cudaStreamSynchronize(posStream) != cudaSuccess) return -2;
        // batchcorrmanifold.cu:2555-2560 and :2565-2566, verbatim arguments
        BCM_VelMeasReduction<<<threadsPerMeasRedBlock, threadsPerVelManiBlock,
                               sizeof(dsp::utils::weightState_t<double>) * threadsPerVelManiBlock, velStream>>>(
            satStates_d, carrScores_d, xCurr_d, gridVelLocs_d, enu2ecefMat_d, carrFrequency_d, txTimePtr_d, *rxTimePtr,
            numChan, currSamplingFreq, *numfftPointsPtr, dopplerSign_d, weightedVelStates_d);
        BCM_ReduceAndVelMeas<<<1, threadsPerMeasRedBlock,
                               sizeof(dsp::utils::weightState_t<double>) * threadsPerMeasRedBlock, velStream>>>(
            weightedVelStates_d, threadsPerMeasRedBlock, z_d, R_d);
        if (cudaStreamSynchronize(velStream) != cudaSuccess) return -4;
        cudaMemcpy(z, z_d, 8 * sizeof(double), cudaMemcpyDeviceToHost);
        cudaMemcpy(parts, weightedPosStates_d, 40 * sizeof(double), cudaMemcpyDeviceToHost);
        cudaMemcpy(vparts, weightedVelStates_d, 40 * sizeof(double), cudaMemcpyDeviceToHost);
        cudaFree(z_d);
        cudaFree(R_d);
        return cudaGetLastError() == cudaSuccess ? 0 : -3;
    }
};
#endif

class RefHarness : public dsp::DPEFlow {
  public:
    int Run(int epochs, const std::string& out, int W, int dump) {   // dump: 0 none, 1 everything, 2 light (long runs)
        if (cudaStreamCreate(&cuStream) != cudaSuccess) return -1;
        for (size_t i = 0; i < Mods.size(); ++i)
            if (Mods[i]->Start((void*)&cuStream)) {
                fprintf(stderr, "module %s failed to start\n", Mods[i]->GetModuleName().c_str());
                return -1;
            }
        cudaDeviceSynchronize();
        double total_us = 0;
        int done = 0;
        for (int e = 0; e < epochs; ++e) {
            if (dump) DumpInputs(out, e, dump == 2);
            struct timeval t0, t1;
            int rc = 0;
            for (size_t i = 0; i < Mods.size() && !rc; ++i) {
                rc = Mods[i]->Update((void*)&cuStream);
                if (i == 0) gettimeofday(&t0, NULL);           // as flow.cu:132-135
                if (dump && !rc && Mods[i]->GetModuleName() == "BatchCorrManifold") {
                    cudaStreamSynchronize(cuStream);
                    DumpOutputs(out, e, W, dump == 2);
#ifdef REF_WEIGHTED
                    double zw[8], parts[40], vparts[40];
                    int wr = static_cast<BcmProbe*>(static_cast<dsp::BatchCorrManifold*>(Mods[i]))->RunWeighted(zw, parts, vparts);
                    if (wr) { fprintf(stderr, "weighted kernels failed (%d)\n", wr); exit(6); }
                    dump_host(out, e, "zval_weighted", zw, 4 * sizeof(double));
                    dump_host(out, e, "zval_weighted_vel", zw + 4, 4 * sizeof(double));
                    dump_host(out, e, "weighted_parts", parts, sizeof(parts));
                    dump_host(out, e, "weighted_vel_parts", vparts, sizeof(vparts));
#endif
                }
            }
            cudaStreamSynchronize(cuStream);
            gettimeofday(&t1, NULL);
            if (rc) { fprintf(stderr, "epoch %d: Update failed, stopping\n", e); break; }
            total_us += (t1.tv_sec - t0.tv_sec) * 1e6 + (t1.tv_usec - t0.tv_usec);
            ++done;
            if (dump) DumpDev(out, e, "x_k1k1", "cuEKF", "xCurrk1k1", 8 * sizeof(double));
        }
        for (size_t i = 0; i < Mods.size(); ++i) Mods[i]->Stop();
        printf("REF_EPOCHS %d\nREF_MEAN_EPOCH_US %.1f\n", done, done ? total_us / done : 0.0);
        return done;
    }

  private:
    dsp::Port* P(const char* mod, const char* port) {
        dsp::Port* p = NULL;
        if (GetOutput(mod, port, &p) || !p) { fprintf(stderr, "no port %s.%s\n", mod, port); exit(4); }
        return p;
    }
    void DumpDev(const std::string& out, int e, const char* name, const char* mod, const char* port,
                 size_t bytes, size_t offset = 0) {
        dsp::Port* p = P(mod, port);
        std::vector<char> h(bytes);
        if (p->MemLoc == dsp::CUDA_DEVICE) {
            if (cudaMemcpy(h.data(), (const char*)p->Data + offset, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) {
                fprintf(stderr, "D2H of %s.%s failed\n", mod, port);
                exit(5);
            }
        } else {
            memcpy(h.data(), (const char*)p->Data + offset, bytes);
        }
        dump_host(out, e, name, h.data(), bytes);
    }
    void DumpInputs(const std::string& out, int e, bool light) {
        cudaDeviceSynchronize();
        const int C = P("cuChanMgr", "CodeFrequency")->VectorLength;
        const int CT = P("cuChanMgr", "SatStates")->VectorLength;
        const char* m = "cuChanMgr";
        DumpDev(out, e, "rx_time", m, "rxTime", sizeof(double));
        DumpDev(out, e, "tx_time", m, "txTime", C * sizeof(double));
        DumpDev(out, e, "rc_start", m, "CodePhaseStart", C * sizeof(double));
        DumpDev(out, e, "ri_start", m, "CarrierPhaseStart", C * sizeof(double));
        DumpDev(out, e, "rc_end", m, "CodePhaseEnd", C * sizeof(double));
        DumpDev(out, e, "ri_end", m, "CarrierPhaseEnd", C * sizeof(double));
        DumpDev(out, e, "fc", m, "CodeFrequency", C * sizeof(double));
        DumpDev(out, e, "fi", m, "CarrierFrequency", C * sizeof(double));
        if (!light) DumpDev(out, e, "sat_states", m, "SatStates", (size_t)CT * 8 * sizeof(double));
        DumpDev(out, e, "sat_raw", m, "SatStatesOld", (size_t)C * 8 * sizeof(double));
        DumpDev(out, e, "prn", m, "ValidPRNs", C);
        DumpDev(out, e, "cp_ref", m, "cpReference", C * sizeof(int));
        DumpDev(out, e, "cp_start", m, "cpElapsedStart", C * sizeof(int));
        DumpDev(out, e, "cp_end", m, "cpElapsedEnd", C * sizeof(int));
        DumpDev(out, e, "cp_ref_tow", m, "cpRefTOW", C * sizeof(int));
        DumpDev(out, e, "enu2ecef", m, "ENU2ECEFMat", 9 * sizeof(double));
        DumpDev(out, e, "x_kk1", "cuEKF", "xCurrkk1", 8 * sizeof(double));
        if (e == 0) {
            FILE* f = fopen((out + "/meta.txt").c_str(), "w");
            fprintf(f, "C %d\nCT %d\n", C, CT);
            fclose(f);
        }
    }
    void DumpOutputs(const std::string& out, int e, int W, bool light) {
        const int C = P("cuChanMgr", "CodeFrequency")->VectorLength;
        dsp::Port* smp = P("SampleBlock", "Samples");
        const size_t S = smp->VectorLength;
        // the 20 ms block the reference just processed (int16 I,Q)
        if (!light) DumpDev(out, e, "iq", "SampleBlock", "Samples", S * 2 * sizeof(short));
#ifdef REF_BRIDGE
        // bridge build (oracle/ref_bridge.cu): "CodeScores" is this repository's lag window, not [C][S] rows
        light = true;
#else
        // CodeScores window: fft-shifted bins S/2-W .. S/2+W+1 of every channel
        dsp::Port* cs = P("BatchCorrScores", "CodeScores");
        const int NL = 2 * W + 2;
        std::vector<double> win((size_t)C * NL * 2);
        for (int c = 0; c < C; ++c)
            cudaMemcpy(&win[(size_t)c * NL * 2],
                       (const char*)cs->Data + ((size_t)c * S + S / 2 - W) * 2 * sizeof(double),
                       (size_t)NL * 2 * sizeof(double), cudaMemcpyDeviceToHost);
        dump_host(out, e, "code_scores_win", win.data(), win.size() * sizeof(double));
#endif
        DumpDev(out, e, "zval", "BatchCorrManifold", "zVal", 8 * sizeof(double));
#ifdef REF_BRIDGE
        {
            int Gb = 0;
            GetModParam("BatchCorrManifold", "PosGridDimSize", &Gb);
            DumpDev(out, e, "pos_scores", "BatchCorrManifold", "PosScores", (size_t)Gb * Gb * Gb * Gb * sizeof(double));
        }
#endif
        if (light) return;
        // CarrScores window: fft-shifted bins N_c/2-Wd .. N_c/2+Wd+1 of the zero-padded carrier spectrum
        dsp::Port* cr = P("BatchCorrScores", "CarrScores");
        const int Nc = *(int*)P("BatchCorrScores", "NumFFTPoints")->Data;
        const int Wd = 64, NBd = 2 * Wd + 2;
        std::vector<double> cw((size_t)C * NBd * 2);
        for (int c = 0; c < C; ++c)
            cudaMemcpy(&cw[(size_t)c * NBd * 2],
                       (const char*)cr->Data + ((size_t)c * Nc + Nc / 2 - Wd) * 2 * sizeof(double),
                       (size_t)NBd * 2 * sizeof(double), cudaMemcpyDeviceToHost);
        dump_host(out, e, "carr_scores_win", cw.data(), cw.size() * sizeof(double));
        dump_host(out, e, "n_fft", &Nc, sizeof(int));
        int G = 0;
        GetModParam("BatchCorrManifold", "PosGridDimSize", &G);
        const size_t Gtot = (size_t)G * G * G * G;
        DumpDev(out, e, "pos_scores", "BatchCorrManifold", "PosScores", Gtot * sizeof(double));
        DumpDev(out, e, "time_grid", "BatchCorrManifold", "TimeGrid", (size_t)G * sizeof(double));
    }
};

int main(int argc, char** argv) {
    if (argc < 9) {
        fprintf(stderr, "usage: %s samples.dat handoff.csv rinex grid.csv|none pos_dim vel_dim epochs out_dir [W] [fs] [dump] [ekf] "
                        "[grid_type] [grid_spacing] [lpower]\n",
                argv[0]);
        return 2;
    }
    const int pos_dim = atoi(argv[5]), vel_dim = atoi(argv[6]), epochs = atoi(argv[7]);
    const std::string out = argv[8];
    const int W = argc > 9 ? atoi(argv[9]) : 32;
    const double fs = argc > 10 ? atof(argv[10]) : 2.5e6;
    const int dump = argc > 11 ? atoi(argv[11]) : 1;              // 0 none, 1 everything, 2 light
    const bool ekf = argc > 12 ? atoi(argv[12]) != 0 : false;      // run the reference with its 8-state KF enabled
    // grid.csv == "none": let the reference generate its grid (BCM_InitPosGrid, batchcorrmanifold.cu:148-255)
    const int grid_type = argc > 13 ? atoi(argv[13]) : 0;           // ManifoldGridTypes: 0 Uniform, 2 ArthurBasis
    const float grid_spacing = argc > 14 ? (float)atof(argv[14]) : 1.0f;
    const int lpower = argc > 15 ? atoi(argv[15]) : 1;              // BatchCorrManifold LPower (score = sum |v|^L)
    mkdir(out.c_str(), 0755);
    if (!getenv("HOME")) setenv("HOME", "/tmp", 1);

    // Load cuFFT's kernels before the flow starts: the reference creates its plans inside the first
    // Update (batchcorrscores.cu:1007-1035) and its reader thread gives up after 1.5 s
    // (sampleblock.cu:432); a cold library load on a fresh box can take longer than that.
    {
        cufftHandle warm;
        if (cufftPlan1d(&warm, 50000, CUFFT_Z2Z, 8) == CUFFT_SUCCESS) cufftDestroy(warm);
        if (cufftPlan1d(&warm, 524288, CUFFT_Z2Z, 8) == CUFFT_SUCCESS) cufftDestroy(warm);
        cudaDeviceSynchronize();
    }
    RefHarness flow;
    if (flow.LoadFlow(NULL)) { fprintf(stderr, "LoadFlow failed\n"); return 1; }
    int rc = 0;
    rc |= flow.SetModParam("SampleBlock", "Filename", argv[1]);
    rc |= flow.SetModParam("SampleBlock", "SamplingFrequency", fs);
    rc |= flow.SetModParam("DPInit", "HandoffFilename", argv[2]);
    rc |= flow.SetModParam("DPInit", "RINEXFilename", argv[3]);
    rc |= flow.SetModParam("BatchCorrManifold", "LoadPosGridFilename", argv[4]);
    rc |= flow.SetModParam("BatchCorrManifold", "LoadPosGrid", strcmp(argv[4], "none") != 0);
    rc |= flow.SetModParam("BatchCorrManifold", "LPower", lpower);
    rc |= flow.SetModParam("BatchCorrManifold", "GridType", grid_type);
    rc |= flow.SetModParam("BatchCorrManifold", "GridDimSpacing", grid_spacing);
    rc |= flow.SetModParam("BatchCorrManifold", "PosGridDimSize", pos_dim);
    rc |= flow.SetModParam("BatchCorrManifold", "VelGridDimSize", vel_dim);
    rc |= flow.SetModParam("BatchCorrManifold", "GridLogFileName", (out + "/grid_log.csv").c_str());
    rc |= flow.SetModParam("XECEFLogger", "Filename", (out + "/XFile.csv").c_str());
    if (ekf) rc |= flow.SetModParam("cuEKF", "EnableEKF", true);
    if (rc) { fprintf(stderr, "SetModParam failed\n"); return 1; }
    const int done = flow.Run(epochs, out, W, dump);
    fflush(stdout);
    // the reference's module destructors free buffers they do not own; leave
    // without running them once the results are on disk
    _exit(done == epochs ? 0 : 1);
}
