// modules.h -- the seven modules of the DPE flow, same names / ports / params as the reference
// (cudarecv/dsp/src/dpeflow.cpp:56-62, port tables in each module's constructor).  The three
// hot-path modules (SampleBlock, BatchCorrScores, BatchCorrManifold) do their device work
// exclusively through the C ABI of libdpe_b200.so; the others are host C++ (north star).
#ifndef DPE_HOST_MODULES_H_
#define DPE_HOST_MODULES_H_

#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>
#include "../../include/dpe_b200.h"
#include "gnss.h"
#include "module.h"

namespace dsp {

/** One dpe_ctx per flow, shared by the modules of that flow (they all receive the same
 *  `void* cuFlowStream`, which keys the registry).  BatchCorrManifold::Start creates it. */
struct RankPool;
struct SharedCtx {
    dpe_ctx* ctx = nullptr;       // rank 0 (the only one when NumGPUs = 1)
    dpe_epoch ep;                 // filled in two halves by BatchCorrScores / BatchCorrManifold
    int score_mode = DPE_SCORE_LOOKUP;
    int est_mode = DPE_EST_ARGMAX;
    // NumGPUs > 1: the grid is sharded over one context per GPU, each driven by its own host thread; the
    // whole epoch is then ONE dpe_epoch_run_dist per rank, issued from BatchCorrManifold::Update
    // (BatchCorrScores::Update only records the block and the channel half of `ep`)
    RankPool* pool = nullptr;
    const int16_t* iq = nullptr;  // the block SampleBlock handed over this epoch
    int64_t iq_len = 0;
};
SharedCtx* SharedFor(void* cuFlowStream);
void SharedRelease(void* cuFlowStream);
inline void* StreamOf(void* cuFlowStream) { return *static_cast<void**>(cuFlowStream); }

/** Parses the RINEX navigation file and the handoff CSV (dpinit.cpp:118-201); ends the flow
 *  after MaxEpochs iterations (3000 in the reference, dpinit.cpp:231). */
class DPInit : public Module {
  public:
    DPInit();
    int Start(void*) override;
    int Update(void*) override;
    int Stop() override { Started = false; return 0; }

  private:
    static const unsigned int kNameCap = 1024;
    char HandoffFilename[kNameCap] = {0}, RINEXFilename[kNameCap] = {0};
    float initDeltaX = 0, initDeltaY = 0, initDeltaZ = 0, initDeltaT = 0;
    int MaxEpochs = 3000;
    bool Started = false;
    long loopCounter = 0;
    long long initByte = 0;
    char initPRN[gnss::kPrnMax] = {0};
    double initRC[gnss::kPrnMax], initRI[gnss::kPrnMax], initFC[gnss::kPrnMax], initFI[gnss::kPrnMax];
    int initCP[gnss::kPrnMax], initCPTimestamp[gnss::kPrnMax], initCPRefTOW[gnss::kPrnMax];
    double initX[8], initP[64], initRxTime = 0;
    int initK = 0;
    std::vector<gnss::EphSet> initEph;
};

/** Reader thread (capture file or TCP stream) -> ring of page-locked 20 ms blocks
 *  (sampleblock.cu:102-156,312-515). */
class SampleBlock : public Module {
  public:
    SampleBlock();
    ~SampleBlock() override;
    int Start(void*) override;
    int Update(void*) override;
    int Stop() override;

  private:
    static const unsigned int kNameCap = 1024;
    static const int kNumBlocks = 32;              // NumBlocksDefault of the reference
    char Filename[kNameCap] = {0}, Hostname[kNameCap] = {0};
    int PortNo = 0;
    double SamplingFrequency = 2.5e6, SampleLength = 0.02;
    bool RunLive = false;
    char InputSourceType = 0;
    int64_t BlockLength = 0;                       // samples per block (64-bit: 10 MHz works)
    std::vector<int16_t*> Blocks;                  // pinned host ring
    std::vector<int16_t*> DevBlocks;               // device ring: the reader thread uploads on its own stream (sampleblock.cu:221,403)
    void* readerStream = nullptr;
    int Device = 0;
    int fd = -1;                                   // capture file or connected TCP socket
    int OpenSource();
    std::thread reader;
    std::mutex mu;
    std::condition_variable cv;
    int filled = 0, freeSlots = 0, loadIdx = 0, procIdx = -1;
    bool eof = false, firstUpdate = true, Started = false;
    std::atomic<bool> KeepRunning{false};
    void ReaderThread();
};

/** int16 unpack, wipe-off, replica, windowed code correlogram: dpe_block_stage +
 *  dpe_epoch_set_part(CHANNELS) + dpe_replica_prepare + dpe_correlogram. */
class BatchCorrScores : public Module {
  public:
    BatchCorrScores();
    int Start(void*) override;
    int Update(void*) override;
    int Stop() override { Started = false; return 0; }

  private:
    bool Started = false;
    int numFFTPoints = 0;
};

/** Position-clock manifold: owns the grid (generated or CSV), dpe_epoch_set_part(GEOMETRY) +
 *  dpe_score_pos + dpe_estimate + dpe_result_fetch -> zVal / RVal. */
class BatchCorrManifold : public Module {
  public:
    BatchCorrManifold();
    int Start(void*) override;
    int Update(void*) override;
    int Stop() override;

  private:
    static const unsigned int kNameCap = 1024;
    int posGridDimSize = 25, velGridDimSize = 25, gridType = 0, LPower = 1;
    float gridDimSpacing = 1.0f;
    bool loadPosGrid = false;
    char Filename[kNameCap] = "", loadPosGridFilename[kNameCap] = "";
    // extensions (not in the reference): scoring path, estimator, lag window
    bool bruteForce = false, weightedMean = false;
    int lagHalfwidth = 0, doppHalfwidth = 0;       // 0 = sized from the extent of the grids
    int numGPUs = 1, device = 0;                   // NumGPUs GPUs starting at ordinal Device
    bool Started = false, haveVel = false;
    void* flowStream = nullptr;
    std::vector<double> grid, timeGrid;
    double zVal[16], RVal[64];
    dpe_result last;
};

/** 8-state filter between the manifold and the channel manager.  The DPE flow runs it disabled
 *  (EnableEKF=false, dpeflow.cpp:90): the measurement is copied into the state (EKF_PassMeas,
 *  cuekf.cu:147-159,577-592).  Enabled: linear KF with H = I, random-walk F, speed-adaptive Q
 *  (cuekf.cu:42-81,625-742), on the host -- 8x8 FP64 matrices. */
class cuEKF : public Module {
  public:
    cuEKF();
    int Start(void*) override;
    int Update(void*) override;
    int Stop() override { Started = false; return 0; }

  private:
    double SampleLength = 0.02;
    bool EnableEKF = false, Started = false;
    double xkk1[8], xk1k1[8], Pkk1[64], Pk1k1[64], F[64], Q[64];
    double lpfVals[20], lpfAvg = 0;
    int lpfIdx = 0;
    long measIdx = 0, prevMeasIdx = 0;
    int StepUpdate(const double* z, const double* R);
    void StepPredict();
};

/** Channel manager + satellite states on the host (cuchanmgr.cu:240-923 runs them on the GPU). */
class cuChanMgr : public Module {
  public:
    cuChanMgr();
    int Start(void*) override;
    int Update(void*) override;
    int Stop() override { Started = false; return 0; }

  private:
    int dopplerSign = 1;
    bool Started = false;
    int numChan = 0, timeDim = 1;
    double T = 0.02, rxTime = 0;
    uint8_t PRNs[gnss::kPrnMax];
    double rcStart[gnss::kPrnMax], rcEnd[gnss::kPrnMax], riStart[gnss::kPrnMax], riEnd[gnss::kPrnMax];
    double fc[gnss::kPrnMax], fi[gnss::kPrnMax], txTime[gnss::kPrnMax];
    int cpStart[gnss::kPrnMax], cpEnd[gnss::kPrnMax], cpRef[gnss::kPrnMax], cpRefTOW[gnss::kPrnMax];
    int dopplerSignArr[gnss::kPrnMax];
    gnss::SatState sat[gnss::kPrnMax];
    std::vector<gnss::SatState> batchSat;
    double enu2ecef[9];
    const std::vector<gnss::EphSet>* nav = nullptr;
    int TimeUpdate(int i, const double* x, double rxT);
    void GridPrep(const double* xkk1, const double* timeGrid);
};

/** Port tap -> CSV ("%f, " per value, one row per Update) or binary (datalogger.cu:141-213). */
class DataLogger : public Module {
  public:
    explicit DataLogger(const char* name = "DataLogger");
    int Start(void*) override;
    int Update(void*) override;
    int Stop() override;

  private:
    static const unsigned int kNameCap = 1024;
    char Filename[kNameCap] = {0};
    bool csv = true, Started = false;
    FILE* fp = nullptr;
    struct Row { int dtype = 0; int64_t n = 0; std::vector<unsigned char> bytes; };
    std::deque<Row> rows;                          // filled by the flow thread, drained by the writer thread
    std::thread writer;
    std::mutex mu;
    std::condition_variable cv;
    bool KeepRunning = false, writeFailed = false;
    void WriterThread();
};

}  // namespace dsp
#endif
