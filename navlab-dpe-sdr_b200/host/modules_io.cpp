// modules_io.cpp -- DPInit, SampleBlock, cuEKF (pass-through), DataLogger: the host modules
// either side of the hot path.
#include <fcntl.h>
#include <netdb.h>
#include <sys/socket.h>
#include <sys/types.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstring>
#include <iostream>
#include "modules.h"

namespace dsp {

// ------------------------------------------------------------------------------------------
// DPInit (cudarecv/modules/src/dpinit.cpp:66-238)
// ------------------------------------------------------------------------------------------
DPInit::DPInit() {
    ModuleName = "DPInit";
    AllocateInputs(0);
    AllocateOutputs(14);
    ConfigOutput(0, "StartByte", INT_t, VALUE, HOST, 1, &initByte, 0);
    ConfigOutput(1, "InitPRN", CHAR_t, VALUE, HOST, VECTORLENGTH_ANY, initPRN, 0);
    ConfigOutput(2, "InitCodePhase", DOUBLE_t, VALUE, HOST, VECTORLENGTH_ANY, initRC, 0);
    ConfigOutput(3, "InitCarrierPhase", DOUBLE_t, VALUE, HOST, VECTORLENGTH_ANY, initRI, 0);
    ConfigOutput(4, "InitCodeFrequency", DOUBLE_t, FREQUENCY_HZ, HOST, VECTORLENGTH_ANY, initFC, 0);
    ConfigOutput(5, "InitCarrierFrequency", DOUBLE_t, FREQUENCY_HZ, HOST, VECTORLENGTH_ANY, initFI, 0);
    ConfigOutput(6, "InitElapsedCodePeriods", INT_t, VALUE, HOST, VECTORLENGTH_ANY, initCP, 0);
    ConfigOutput(7, "InitReferenceCodePeriods", INT_t, VALUE, HOST, VECTORLENGTH_ANY, initCPTimestamp, 0);
    ConfigOutput(8, "InitCPRefTOW", INT_t, VALUE, HOST, VECTORLENGTH_ANY, initCPRefTOW, 0);
    ConfigOutput(9, "InitX", DOUBLE_t, STATE, HOST, VECTORLENGTH_ANY, initX, 0);
    ConfigOutput(10, "InitP", DOUBLE_t, COVARIANCE, HOST, VECTORLENGTH_ANY, initP, 0);
    ConfigOutput(11, "InitK", INT_t, VALUE, HOST, 1, &initK, 0);
    ConfigOutput(12, "InitRXTime", DOUBLE_t, VALUE, HOST, 1, &initRxTime, 0);
    ConfigOutput(13, "InitEph", UNDEFINED_t, EPHEMS, HOST, 1, &initEph, 0);
    InsertParam("HandoffFilename", HandoffFilename, CHAR_t, kNameCap, 0);
    InsertParam("RINEXFilename", RINEXFilename, CHAR_t, kNameCap, 0);
    InsertParam("InitDeltaX", &initDeltaX, FLOAT_t, sizeof(float), sizeof(float));
    InsertParam("InitDeltaY", &initDeltaY, FLOAT_t, sizeof(float), sizeof(float));
    InsertParam("InitDeltaZ", &initDeltaZ, FLOAT_t, sizeof(float), sizeof(float));
    InsertParam("InitDeltaT", &initDeltaT, FLOAT_t, sizeof(float), sizeof(float));
    InsertParam("MaxEpochs", &MaxEpochs, INT_t, sizeof(int), sizeof(int));      // 3000 in the reference
    std::memset(initX, 0, sizeof(initX));
}

int DPInit::Start(void*) {
    if (Started) return 0;
    if (gnss::ReadRinexNav(RINEXFilename, &initEph)) {
        std::clog << "[" << ModuleName << "] Open RINEXParamsFile failed: " << RINEXFilename << std::endl;
        return -1;
    }
    gnss::Handoff h;
    if (gnss::ReadHandoff(HandoffFilename, &h)) {
        std::clog << "[" << ModuleName << "] Open handoffParamsFile failed: " << HandoffFilename << std::endl;
        return -1;
    }
    const int n = (int)h.prn.size();
    if (n > gnss::kPrnMax) return -1;
    for (int i = 0; i < n; ++i) {
        initPRN[i] = (char)h.prn[i];
        initRC[i] = h.rc[i]; initRI[i] = h.ri[i]; initFC[i] = h.fc[i]; initFI[i] = h.fi[i];
        initCP[i] = h.cp[i]; initCPTimestamp[i] = h.cp_timestamp[i]; initCPRefTOW[i] = h.TOW[i];
    }
    for (int i = 0; i < 8; ++i) initX[i] = h.X_ECEF[i];
    // InitDelta{X,Y,Z,T}: ECEF / clock offset of the first grid centre (PerturbInitialization, dpinit.cpp:55-61)
    initX[0] += initDeltaX; initX[1] += initDeltaY; initX[2] += initDeltaZ; initX[3] += initDeltaT;
    for (int i = 0; i < 64; ++i) initP[i] = (i % 9 == 0) ? 1.0 : 0.0;
    initK = 0;
    initRxTime = h.rxTime;
    initByte = h.bytes_read;
    loopCounter = 0;
    for (int id = 1; id <= 8; ++id) UpdateOutput((unsigned char)id, n, outputs[id].Data, 0);
    UpdateOutput(9, 8, initX, 0);
    UpdateOutput(10, 64, initP, 0);
    UpdateOutput(13, (int64_t)initEph.size(), &initEph, 0);
    std::clog << "[" << ModuleName << "] Updated outputs" << std::endl;
    Started = true;
    return 0;
}

int DPInit::Update(void*) {
    if (loopCounter % 500 == 0) std::clog << "[" << ModuleName << "] Started iteration " << loopCounter << std::endl;
    loopCounter++;
    return (loopCounter >= MaxEpochs) ? -1 : 0;       // "bootleg way to end the pipeline", dpinit.cpp:229-235
}

// ------------------------------------------------------------------------------------------
// SampleBlock (cudarecv/modules/src/sampleblock.cu)
// ------------------------------------------------------------------------------------------
SampleBlock::SampleBlock() {
    ModuleName = "SampleBlock";
    AllocateInputs(1);
    ConfigExpectedInput(0, "StartByte", INT_t, VALUE, 1);
    AllocateOutputs(3);
    // the reference hands over a device pointer filled by its reader thread; here the block stays in
    // page-locked host memory and BatchCorrScores stages it with dpe_block_stage on the flow stream
    ConfigOutput(0, "Samples", UNDEFINED_t, VALUE_CMPX, CUDA_DEVICE, 2, nullptr, 0);      // a device block, like the reference's
    ConfigOutput(1, "SamplingFrequency", DOUBLE_t, FREQUENCY_HZ, HOST, 1, &SamplingFrequency, 0);
    ConfigOutput(2, "SampleLength", DOUBLE_t, VALUE, HOST, 1, &SampleLength, 0);
    InsertParam("Filename", Filename, CHAR_t, kNameCap, 0);
    InsertParam("Hostname", Hostname, CHAR_t, kNameCap, 0);
    InsertParam("PortNo", &PortNo, INT_t, sizeof(int), sizeof(int));
    InsertParam("SamplingFrequency", &SamplingFrequency, DOUBLE_t, sizeof(double), sizeof(double));
    InsertParam("SampleLength", &SampleLength, DOUBLE_t, sizeof(double), sizeof(double));
    InsertParam("RunLive", &RunLive, BOOL_t, sizeof(bool), sizeof(bool));
    InsertParam("InputSourceType", &InputSourceType, CHAR_t, sizeof(char), sizeof(char));
    InsertParam("Device", &Device, INT_t, sizeof(int), sizeof(int));           // CUDA ordinal of the device ring
}

SampleBlock::~SampleBlock() { Stop(); }

// SampleBlock::Open (sampleblock.cu:102-156): a capture file positioned at StartByte, or a TCP
// client socket to Hostname:PortNo streaming the same interleaved int16 I/Q.
int SampleBlock::OpenSource() {
    if (InputSourceType == 0) {                                    // SAMPLE_INPUT_SOURCE_FILE
        fd = ::open(Filename, O_RDONLY);
        if (fd < 0) { std::cerr << "[SampleBlock] Unable to open file: " << Filename << std::endl; return -1; }
        const long long start = *In<long long>(0);
        // the reference rejects StartByte 0 (lseek()==0 is read as a failure, sampleblock.cu:123-128); accepted here
        if (start < 0 || lseek(fd, (off_t)start, SEEK_SET) != (off_t)start) {
            std::cerr << "[SampleBlock] Failed to skip ahead in file: " << Filename << std::endl;
            ::close(fd); fd = -1;
            return -1;
        }
        std::clog << "[" << ModuleName << "] Starting reading at byte " << start << " in file " << Filename << std::endl;
        return 0;
    }
    if (InputSourceType == 1) {                                    // SAMPLE_INPUT_SOURCE_SOCKET
        struct addrinfo hints, *res = nullptr;
        std::memset(&hints, 0, sizeof(hints));
        hints.ai_family = AF_INET;
        hints.ai_socktype = SOCK_STREAM;
        char port[16];
        std::snprintf(port, sizeof(port), "%d", PortNo);
        if (getaddrinfo(Hostname, port, &hints, &res) || !res) {
            std::cerr << "[" << ModuleName << "] Open: Invalid hostname." << std::endl;
            return -1;
        }
        fd = ::socket(res->ai_family, res->ai_socktype, res->ai_protocol);
        if (fd < 0) {
            std::cerr << "[" << ModuleName << "] Open: Unable to open socket." << std::endl;
            freeaddrinfo(res);
            return -1;
        }
        if (::connect(fd, res->ai_addr, res->ai_addrlen) < 0) {
            std::cerr << "[" << ModuleName << "] Open: Failed to connect." << std::endl;
            freeaddrinfo(res);
            ::close(fd); fd = -1;
            return -1;
        }
        freeaddrinfo(res);
        return 0;
    }
    std::cerr << "[" << ModuleName << "] Open: Invalid input source type." << std::endl;
    return -1;
}

int SampleBlock::Start(void*) {
    if (Started) return 0;
    if (!InputsConnected()) return -1;
    if (OpenSource()) return -1;
    BlockLength = (int64_t)(SamplingFrequency * SampleLength + 0.5);
    Blocks.assign(kNumBlocks, nullptr);
    DevBlocks.assign(kNumBlocks, nullptr);
    for (int i = 0; i < kNumBlocks; ++i)
        if (dpe_host_alloc((void**)&Blocks[i], sizeof(int16_t) * 2 * (size_t)BlockLength) ||
            dpe_device_alloc((void**)&DevBlocks[i], sizeof(int16_t) * 2 * (size_t)BlockLength, Device)) {
            std::cerr << "[SampleBlock] Unable to allocate sample buffers: " << dpe_last_error() << std::endl;
            return -1;
        }
    if (dpe_stream_create_on(&readerStream, Device)) {
        std::cerr << "[SampleBlock] Unable to create the upload stream: " << dpe_last_error() << std::endl;
        return -1;
    }
    UpdateOutput(0, BlockLength, nullptr, 0);
    filled = 0; freeSlots = kNumBlocks; loadIdx = 0; procIdx = -1; eof = false; firstUpdate = true;
    KeepRunning = true;
    reader = std::thread(&SampleBlock::ReaderThread, this);
    Started = true;
    return 0;
}

void SampleBlock::ReaderThread() {
    const size_t bytes = sizeof(int16_t) * 2 * (size_t)BlockLength;
    long blockCnt = 0;
    while (KeepRunning) {
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return freeSlots > 0 || !KeepRunning; });
            if (!KeepRunning) break;
        }
        size_t got = 0;                           // read() until the whole block is in (sampleblock.cu:353-373)
        bool failed = false;
        while (got < bytes) {
            const ssize_t n = ::read(fd, reinterpret_cast<char*>(Blocks[loadIdx]) + got, bytes - got);
            if (n < 0 && errno == EINTR) continue;
            if (n < 0) { failed = true; std::perror("[SampleBlock] Read Error"); break; }
            if (n == 0) break;
            got += (size_t)n;
        }
        if (failed || got != bytes) {             // partial block at EOF is dropped like the reference does
            std::lock_guard<std::mutex> lk(mu);
            eof = true;
            cv.notify_all();
            std::clog << "[SampleBlock] Reached EOF." << std::endl << "[SampleBlock] blockCnt = " << blockCnt << std::endl;
            break;
        }
        // H2D ahead of time on the reader's own stream (sampleblock.cu:403): the flow thread gets a resident block
        if (dpe_copy_h2d(DevBlocks[loadIdx], Blocks[loadIdx], bytes, readerStream, Device) || dpe_stream_sync(readerStream)) {
            std::cerr << "[SampleBlock] upload failed: " << dpe_last_error() << std::endl;
            std::lock_guard<std::mutex> lk(mu);
            eof = true;
            cv.notify_all();
            break;
        }
        ++blockCnt;
        loadIdx = (loadIdx + 1) % kNumBlocks;
        std::lock_guard<std::mutex> lk(mu);
        --freeSlots; ++filled;
        // live capture with every buffer full: the next samples will be lost (sampleblock.cu:419-421)
        if (RunLive && freeSlots == 0) std::clog << "[SampleBlock] Fail real-time." << std::endl;
        cv.notify_all();
    }
}

int SampleBlock::Update(void*) {
    if (!Started) { std::cerr << "[SampleBlock] Thread not running." << std::endl; return -1; }
    std::unique_lock<std::mutex> lk(mu);
    if (firstUpdate) firstUpdate = false;
    else { ++freeSlots; cv.notify_all(); }       // the block of the previous epoch may be refilled now
    // 1.5 s like the reference (sampleblock.cu:484); buffers are only released in Stop(), after the flow ended
    if (!cv.wait_for(lk, std::chrono::milliseconds(1500), [&] { return filled > 0 || eof; })) {
        std::cerr << "[SampleBlock] Error: sem_timewait timeout: samplesAvailSem" << std::endl;
        return -1;
    }
    if (filled == 0) return -1;                    // EOF: ends the flow
    --filled;
    procIdx = (procIdx + 1) % kNumBlocks;
    outputs[0].Data = DevBlocks[procIdx];
    return 0;
}

int SampleBlock::Stop() {
    // (also after a Start() that failed half way: the source, the buffers and the stream it got that far are released here)
    if (!Started && fd < 0 && Blocks.empty() && !readerStream) return 0;
    KeepRunning = false;
    { std::lock_guard<std::mutex> lk(mu); cv.notify_all(); }
    // a reader blocked in read() on an idle TCP sender must come back: shut the socket down first
    // (ENOTSOCK on a capture file, harmless)
    if (fd >= 0) ::shutdown(fd, SHUT_RDWR);
    if (reader.joinable()) reader.join();
    for (size_t i = 0; i < Blocks.size(); ++i) dpe_host_free(Blocks[i]);
    Blocks.clear();
    for (size_t i = 0; i < DevBlocks.size(); ++i) dpe_device_free(DevBlocks[i], Device);
    DevBlocks.clear();
    if (readerStream) { dpe_stream_destroy(readerStream); readerStream = nullptr; }
    if (fd >= 0) { ::close(fd); fd = -1; }
    Started = false;
    return 0;
}

// ------------------------------------------------------------------------------------------
// cuEKF: the DPE flow runs it with EnableEKF=false (dpeflow.cpp:90) => EKF_PassMeas
// ------------------------------------------------------------------------------------------
cuEKF::cuEKF() {
    ModuleName = "cuEKF";
    AllocateInputs(5);
    ConfigExpectedInput(0, "InitX", DOUBLE_t, STATE, VECTORLENGTH_ANY);
    ConfigExpectedInput(1, "InitP", DOUBLE_t, COVARIANCE, VECTORLENGTH_ANY);
    ConfigExpectedInput(2, "InitK", INT_t, VALUE, 1);
    ConfigExpectedInput(3, "zVal", DOUBLE_t, STATE, VECTORLENGTH_ANY);
    ConfigExpectedInput(4, "RVal", DOUBLE_t, COVARIANCE, VECTORLENGTH_ANY);
    AllocateOutputs(3);
    ConfigOutput(0, "xCurrkk1", DOUBLE_t, STATE, HOST, 8, xkk1, 0);
    ConfigOutput(1, "PCurrkk1", DOUBLE_t, COVARIANCE, HOST, 64, Pkk1, 0);
    ConfigOutput(2, "xCurrk1k1", DOUBLE_t, STATE, HOST, 8, xk1k1, 0);
    InsertParam("SampleLength", &SampleLength, DOUBLE_t, sizeof(double), sizeof(double));
    InsertParam("EnableEKF", &EnableEKF, BOOL_t, sizeof(bool), sizeof(bool));
}

// 8x8 row-major helpers
static void mat_mul(const double* A, const double* B, double* C, bool transB) {
    for (int r = 0; r < 8; ++r)
        for (int c = 0; c < 8; ++c) {
            double s = 0;
            for (int k = 0; k < 8; ++k) s += A[r * 8 + k] * (transB ? B[c * 8 + k] : B[k * 8 + c]);
            C[r * 8 + c] = s;
        }
}

static bool mat_inv(const double* A, double* inv) {            // Gauss-Jordan, partial pivoting
    double a[8][16];
    for (int r = 0; r < 8; ++r)
        for (int c = 0; c < 8; ++c) { a[r][c] = A[r * 8 + c]; a[r][8 + c] = (r == c) ? 1.0 : 0.0; }
    for (int p = 0; p < 8; ++p) {
        int best = p;
        for (int r = p + 1; r < 8; ++r)
            if (std::fabs(a[r][p]) > std::fabs(a[best][p])) best = r;
        if (a[best][p] == 0.0) return false;
        if (best != p)
            for (int c = 0; c < 16; ++c) std::swap(a[p][c], a[best][c]);
        const double d = a[p][p];
        for (int c = 0; c < 16; ++c) a[p][c] /= d;
        for (int r = 0; r < 8; ++r)
            if (r != p && a[r][p] != 0.0) {
                const double f = a[r][p];
                for (int c = 0; c < 16; ++c) a[r][c] -= f * a[p][c];
            }
    }
    for (int r = 0; r < 8; ++r)
        for (int c = 0; c < 8; ++c) inv[r * 8 + c] = a[r][8 + c];
    return true;
}

int cuEKF::Start(void*) {
    if (Started) return 0;
    if (!InputsConnected()) return -1;
    for (int i = 0; i < 8; ++i) xkk1[i] = xk1k1[i] = In<double>(0)[i];      // cuekf.cu:338-344
    for (int i = 0; i < 64; ++i) {
        Pk1k1[i] = In<double>(1)[i];                                         // InitP
        Pkk1[i] = Q[i] = F[i] = (i % 9 == 0) ? 1.0 : 0.0;                    // EKF_MakeIMatrix (:392-400)
    }
    for (int i = 0; i < 4; ++i) F[i * 8 + 4 + i] = SampleLength;            // EKF_MakeDPERandomWalkFMatrix (:107-135)
    for (int i = 0; i < 20; ++i) lpfVals[i] = 0;
    lpfAvg = 0; lpfIdx = 0; measIdx = 0; prevMeasIdx = -1;      // cuekf.cu:446-447: one predict per update
    if (EnableEKF) StepPredict();   // "the initial value is a k-1|k-1 measurement; predict the next state" (:487-491)
    Started = true;
    return 0;
}

// y = z - H x_kk1, S = H P H' + R, K = P H' S^-1, x = x_kk1 + K y, P = (I - K H) P_kk1 with H = I
// (cuEKF::StepUpdate, cuekf.cu:660-715)
int cuEKF::StepUpdate(const double* z, const double* R) {
    double S[64], Sinv[64], K[64], IKH[64], y[8];
    for (int i = 0; i < 8; ++i) y[i] = z[i] - xkk1[i];
    for (int i = 0; i < 64; ++i) S[i] = Pkk1[i] + R[i];
    if (!mat_inv(S, Sinv)) { std::cerr << "[" << ModuleName << "] Error: StepUpdate() S inversion failed" << std::endl; return -1; }
    mat_mul(Pkk1, Sinv, K, false);
    for (int r = 0; r < 8; ++r) {
        double s = xkk1[r];
        for (int c = 0; c < 8; ++c) s += K[r * 8 + c] * y[c];
        xk1k1[r] = s;
    }
    for (int i = 0; i < 64; ++i) IKH[i] = ((i % 9 == 0) ? 1.0 : 0.0) - K[i];
    mat_mul(IKH, Pkk1, Pk1k1, false);
    return 0;
}

// Q from the 20-tap mean speed (EKF_Update_Q, cuekf.cu:42-81; Q = F Qd F', GetQVal :733-742), then
// x_kk1 = F x_k1k1, P_kk1 = F P_k1k1 F' + Q (StepPredict :636-655)
void cuEKF::StepPredict() {
    const double v = std::sqrt(xk1k1[4] * xk1k1[4] + xk1k1[5] * xk1k1[5] + xk1k1[6] * xk1k1[6]);
    lpfAvg = lpfAvg - lpfVals[lpfIdx] + (v / 20.0);
    lpfVals[lpfIdx] = v / 20.0;
    if (++lpfIdx >= 20) lpfIdx = 0;
    const double vval = 1.0 + 250.0 / std::fmin(std::fmax(lpfAvg * lpfAvg, 50.0), 125.0);
    double Qd[64], T1[64];
    for (int i = 0; i < 64; ++i) Qd[i] = 0;
    Qd[4 * 8 + 4] = Qd[5 * 8 + 5] = Qd[6 * 8 + 6] = vval;
    Qd[63] = (2.5e-10) * (2.5e-10) * 299792458.0 * 299792458.0;            // Q_CLOCK_DRIFT, cuekf.h:28
    mat_mul(F, Qd, T1, false);
    mat_mul(T1, F, Q, true);
    for (int r = 0; r < 8; ++r) {
        double s = 0;
        for (int c = 0; c < 8; ++c) s += F[r * 8 + c] * xk1k1[c];
        xkk1[r] = s;
    }
    mat_mul(F, Pk1k1, T1, false);
    mat_mul(T1, F, Pkk1, true);
    for (int i = 0; i < 64; ++i) Pkk1[i] += Q[i];
}

int cuEKF::Update(void*) {
    if (!Started) return -1;
    if (EnableEKF) {                                   // cuEKF::Update, cuekf.cu:577-589
        if (StepUpdate(In<double>(3), In<double>(4))) return -1;
        while (prevMeasIdx < measIdx) { StepPredict(); ++prevMeasIdx; }
        ++measIdx;
        return 0;
    }
    // exactly 8 values (the reference copies 16, reading past zVal)
    for (int i = 0; i < 8; ++i) xk1k1[i] = xkk1[i] = In<double>(3)[i];
    return 0;
}

// ------------------------------------------------------------------------------------------
// DataLogger (cudarecv/modules/src/datalogger.cu): "%f" values, ", " separated, one row per epoch
// ------------------------------------------------------------------------------------------
DataLogger::DataLogger(const char* name) {
    ModuleName = name;
    AllocateInputs(1);
    ConfigExpectedInput(0, "Data", DATATYPE_ANY, VALUETYPE_ANY, VECTORLENGTH_ANY);
    AllocateOutputs(0);
    InsertParam("Filename", Filename, CHAR_t, kNameCap, 0);
    InsertParam("CSV", &csv, BOOL_t, sizeof(bool), sizeof(bool));
}

int DataLogger::Start(void*) {
    if (Started) return 0;
    if (!InputsConnected()) return -1;
    if (inputs[0]->MemLoc != HOST) { std::cerr << "[" << ModuleName << "] only HOST ports can be logged" << std::endl; return -1; }
    fp = std::fopen(Filename, csv ? "w" : "wb");
    if (!fp) { std::cerr << "[" << ModuleName << "] cannot open " << Filename << std::endl; return -1; }
    KeepRunning = true;
    writeFailed = false;
    writer = std::thread(&DataLogger::WriterThread, this);
    Started = true;
    return 0;
}

// The flow thread only copies the port into a queued row; formatting and file I/O happen on the writer
// thread (the reference: double-buffered D2H + writer pthread, datalogger.cu:113-213,215-278).
int DataLogger::Update(void*) {
    const Port* p = inputs[0];
    const int64_t n = p->Length;
    const size_t sz = (p->Datatype == DOUBLE_t) ? 8 : (p->Datatype == INT_t || p->Datatype == FLOAT_t) ? 4 : 1;
    Row r;
    r.dtype = p->Datatype;
    r.n = n;
    r.bytes.assign(static_cast<const unsigned char*>(p->Data), static_cast<const unsigned char*>(p->Data) + sz * (size_t)n);
    {
        std::lock_guard<std::mutex> lk(mu);
        if (writeFailed) return -1;
        rows.push_back(std::move(r));
    }
    cv.notify_one();
    return 0;
}

void DataLogger::WriterThread() {
    for (;;) {
        Row r;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return !rows.empty() || !KeepRunning; });
            if (rows.empty()) break;                  // stop requested and everything written
            r = std::move(rows.front());
            rows.pop_front();
        }
        bool ok = true;
        if (!csv) {
            ok = std::fwrite(r.bytes.data(), 1, r.bytes.size(), fp) == r.bytes.size();
        } else {
            for (int64_t i = 0; i < r.n && ok; ++i) {
                double v = 0;
                switch (r.dtype) {
                    case DOUBLE_t: v = reinterpret_cast<const double*>(r.bytes.data())[i]; break;
                    case FLOAT_t: v = reinterpret_cast<const float*>(r.bytes.data())[i]; break;
                    case INT_t: v = reinterpret_cast<const int*>(r.bytes.data())[i]; break;
                    default: v = reinterpret_cast<const char*>(r.bytes.data())[i]; break;
                }
                ok = std::fprintf(fp, (i + 1 < r.n) ? "%f, " : "%f\n", v) > 0;
            }
        }
        if (!ok) { std::lock_guard<std::mutex> lk(mu); writeFailed = true; }
    }
}

int DataLogger::Stop() {
    if (writer.joinable()) {
        { std::lock_guard<std::mutex> lk(mu); KeepRunning = false; }
        cv.notify_all();
        writer.join();                                // drains the queue first
    }
    if (fp) { std::fclose(fp); fp = nullptr; }
    Started = false;
    return 0;
}

}  // namespace dsp
