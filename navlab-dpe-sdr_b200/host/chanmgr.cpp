// chanmgr.cpp -- cuChanMgr on the host: channel parameters referenced to
// the start and the end of each block, satellite states from broadcast ephemeris, earth-rotation
// corrected per time-grid point, ENU->ECEF matrix.  Equations: cudarecv/modules/src/cuchanmgr.cu
// CHM_ComputeSatStates :240-306, CHM_PropagateChannels :338-608, CHM_TimeUpdateChannels :641-829,
// CHM_GridPrep :853-923 (the reference runs them as <<<1,64>>> kernels; <= 37 channels of scalar
// FP64 work belong on the host, as the north star says).
#include <cmath>
#include <cstring>
#include <iostream>
#include "modules.h"

namespace dsp {

using namespace gnss;

cuChanMgr::cuChanMgr() {
    ModuleName = "cuChanMgr";
    AllocateInputs(14);
    ConfigExpectedInput(0, "InitEph", UNDEFINED_t, EPHEMS, VECTORLENGTH_ANY);
    ConfigExpectedInput(1, "InitPRN", CHAR_t, VALUE, VECTORLENGTH_ANY);
    ConfigExpectedInput(2, "InitCodePhase", DOUBLE_t, VALUE, VECTORLENGTH_ANY);
    ConfigExpectedInput(3, "InitCarrierPhase", DOUBLE_t, VALUE, VECTORLENGTH_ANY);
    ConfigExpectedInput(4, "InitCodeFrequency", DOUBLE_t, FREQUENCY_HZ, VECTORLENGTH_ANY);
    ConfigExpectedInput(5, "InitCarrierFrequency", DOUBLE_t, FREQUENCY_HZ, VECTORLENGTH_ANY);
    ConfigExpectedInput(6, "InitElapsedCodePeriods", INT_t, VALUE, VECTORLENGTH_ANY);
    ConfigExpectedInput(7, "InitReferenceCodePeriods", INT_t, VALUE, VECTORLENGTH_ANY);
    ConfigExpectedInput(8, "InitCPRefTOW", INT_t, VALUE, VECTORLENGTH_ANY);
    ConfigExpectedInput(9, "InitRXTime", DOUBLE_t, VALUE, 1);
    ConfigExpectedInput(10, "SampleLength", DOUBLE_t, VALUE, 1);
    ConfigExpectedInput(11, "TimeGrid", DOUBLE_t, VALUE, VECTORLENGTH_ANY);
    ConfigExpectedInput(12, "xCurrk1k1", DOUBLE_t, STATE, VECTORLENGTH_ANY);
    ConfigExpectedInput(13, "xCurrkk1", DOUBLE_t, STATE, VECTORLENGTH_ANY);
    InsertParam("DopplerSign", &dopplerSign, INT_t, sizeof(int), sizeof(int));
    AllocateOutputs(18);
    ConfigOutput(0, "rxTime", DOUBLE_t, VALUE, HOST, 1, &rxTime, 0);
    ConfigOutput(1, "txTime", DOUBLE_t, VALUE, HOST, VECTORLENGTH_ANY, txTime, 0);
    ConfigOutput(2, "CodePhaseStart", DOUBLE_t, VALUE, HOST, VECTORLENGTH_ANY, rcStart, 0);
    ConfigOutput(3, "CarrierPhaseStart", DOUBLE_t, VALUE, HOST, VECTORLENGTH_ANY, riStart, 0);
    ConfigOutput(4, "CodePhaseEnd", DOUBLE_t, VALUE, HOST, VECTORLENGTH_ANY, rcEnd, 0);
    ConfigOutput(5, "CarrierPhaseEnd", DOUBLE_t, VALUE, HOST, VECTORLENGTH_ANY, riEnd, 0);
    ConfigOutput(6, "CodeFrequency", DOUBLE_t, FREQUENCY_HZ, HOST, VECTORLENGTH_ANY, fc, 0);
    ConfigOutput(7, "CarrierFrequency", DOUBLE_t, FREQUENCY_HZ, HOST, VECTORLENGTH_ANY, fi, 0);
    ConfigOutput(8, "SatStates", DOUBLE_t, STATE, HOST, VECTORLENGTH_ANY, nullptr, 0);
    ConfigOutput(9, "DopplerSign", INT_t, VALUE, HOST, VECTORLENGTH_ANY, dopplerSignArr, 0);
    ConfigOutput(10, "ValidPRNs", CHAR_t, VALUE, HOST, VECTORLENGTH_ANY, PRNs, 0);
    ConfigOutput(11, "cpReference", INT_t, VALUE, HOST, VECTORLENGTH_ANY, cpRef, 0);
    ConfigOutput(12, "cpElapsedStart", INT_t, VALUE, HOST, VECTORLENGTH_ANY, cpStart, 0);
    ConfigOutput(13, "cpElapsedEnd", INT_t, VALUE, HOST, VECTORLENGTH_ANY, cpEnd, 0);
    ConfigOutput(14, "ENU2ECEFMat", DOUBLE_t, VALUE, HOST, 9, enu2ecef, 0);
    ConfigOutput(15, "SatStatesOld", DOUBLE_t, STATE, HOST, VECTORLENGTH_ANY, sat, 0);
    ConfigOutput(16, "cpRef", INT_t, VALUE, HOST, VECTORLENGTH_ANY, cpRef, 0);
    ConfigOutput(17, "cpRefTOW", INT_t, VALUE, HOST, VECTORLENGTH_ANY, cpRefTOW, 0);
}

static double posfmod(double a, double m) { double r = std::fmod(a, m); return r < 0 ? r + m : r; }

// Enhanced time update of channel i (shared tail of CHM_PropagateChannels :451-602 and
// CHM_TimeUpdateChannels :675-823): predict the code phase T ahead, then re-derive it from the
// geometry at the state x, roll start <- end, recompute txTime and the satellite state.
int cuChanMgr::TimeUpdate(int i, const double* x, double rxT) {
    const double adv = fc[i] * T + rcEnd[i];
    const int cpPred = cpEnd[i] + (int)std::floor(adv / kLCA);
    const double rcPred = posfmod(adv, kLCA);
    const double txPred = cpRefTOW[i] + ((cpPred - cpRef[i]) * kTCA) + (rcPred / kFCA);
    const Eph* eph = SelectEph(*nav, PRNs[i], txPred);
    SatState sp;
    if (!eph || !SatPosition(*eph, txPred, &sp)) {
        std::cerr << "[" << ModuleName << "] no usable ephemeris for PRN " << (int)PRNs[i] << std::endl;
        return -1;
    }
    const double tau = rxT + T - (txPred + (x[3] / kC)) + sp.clkb;
    const SatState r = RotateSat(sp, tau);
    const double lx = r.x - x[0], ly = r.y - x[1], lz = r.z - x[2];
    const double range = std::sqrt(lx * lx + ly * ly + lz * lz);
    const double pr = range - kC * r.clkb + x[3];
    const double bcTx = rxT + T - pr / kC;
    const double frac = bcTx - cpRefTOW[i] - ((cpEnd[i] - cpRef[i]) * kTCA);
    const double bcRc = frac * kFCA;
    cpStart[i] = cpEnd[i];
    rcStart[i] = rcEnd[i];
    cpEnd[i] += (int)std::floor(bcRc / kLCA);
    rcEnd[i] = posfmod(bcRc, kLCA);
    riStart[i] = riEnd[i];
    riEnd[i] = posfmod(fi[i] * T + riEnd[i], 1.0);
    txTime[i] = TxTime(cpRefTOW[i], cpEnd[i], cpRef[i], rcEnd[i]);
    return SatPosition(*eph, txTime[i], &sat[i]) ? 0 : -1;
}

void cuChanMgr::GridPrep(const double* xkk1, const double* timeGrid) {
    batchSat.resize((size_t)numChan * timeDim);
    for (int c = 0; c < numChan; ++c)
        for (int it = 0; it < timeDim; ++it) {
            const double tau = rxTime - (txTime[c] + ((timeGrid[it] + xkk1[3]) / kC)) + sat[c].clkb;
            batchSat[(size_t)c * timeDim + it] = RotateSat(sat[c], tau);
        }
    double lat, lon;
    EcefToLatLon(xkk1, &lat, &lon);
    EnuToEcefMatrix(lat, lon, enu2ecef);
    UpdateOutput(8, (int64_t)numChan * timeDim, batchSat.data(), 0);
}

int cuChanMgr::Start(void*) {
    if (Started) return 0;
    if (!InputsConnected()) return -1;
    nav = static_cast<const std::vector<EphSet>*>(inputs[0]->Data);
    numChan = (int)InLen(1);
    if (numChan < 1 || numChan > kPrnMax) return -1;
    T = std::floor(*In<double>(10) * 1.0e6 + 0.5) / 1.0e6;        // cuchanmgr.cu:1037
    rxTime = *In<double>(9);
    for (int i = 0; i < numChan; ++i) {
        PRNs[i] = (uint8_t)In<char>(1)[i];
        rcStart[i] = 0; riStart[i] = 0; cpStart[i] = 0;
        rcEnd[i] = In<double>(2)[i];                               // handoff lands in the "end" slots (:1053-1069)
        riEnd[i] = In<double>(3)[i];
        fc[i] = In<double>(4)[i];
        fi[i] = In<double>(5)[i];
        cpEnd[i] = In<int>(6)[i];
        cpRef[i] = In<int>(7)[i];
        cpRefTOW[i] = In<int>(8)[i];
        dopplerSignArr[i] = dopplerSign;
        txTime[i] = TxTime(cpRefTOW[i], cpEnd[i], cpRef[i], rcEnd[i]);
        const Eph* eph = SelectEph(*nav, PRNs[i], txTime[i]);
        if (!eph || !SatPosition(*eph, txTime[i], &sat[i])) {
            std::cerr << "[" << ModuleName << "] no usable ephemeris for PRN " << (int)PRNs[i] << std::endl;
            return -1;
        }
    }
    const double* x = In<double>(12);                              // cuEKF state (InitX at this point)
    for (int i = 0; i < numChan; ++i)
        if (TimeUpdate(i, x, rxTime)) return -1;
    rxTime += T;                                                   // :1121
    timeDim = (int)InLen(11);
    for (int id : {1, 2, 3, 4, 5, 6, 7, 9, 10, 11, 12, 13, 15, 16, 17}) UpdateOutput((unsigned char)id, numChan, outputs[id].Data, 0);
    GridPrep(In<double>(13), In<double>(11));
    Started = true;
    return 0;
}

// CHM_PropagateChannels: measurement update of fi / fc from the new fix x = x_{k|k}, then the
// enhanced time update; rxTime += T; GridPrep for the next epoch (cuchanmgr.cu:1224-1268).
int cuChanMgr::Update(void*) {
    if (!Started) return -1;
    const double* x = In<double>(12);
    for (int i = 0; i < numChan; ++i) {
        const double tau = rxTime - (txTime[i] + (x[3] / kC)) + sat[i].clkb;
        const SatState r = RotateSat(sat[i], tau);
        const double ve[4] = {x[4] - kOmegaE * x[1], x[5] + kOmegaE * x[0], x[6], x[7]};
        const double lx = r.x - x[0], ly = r.y - x[1], lz = r.z - x[2];
        const double range = std::sqrt(lx * lx + ly * ly + lz * lz);
        const double rate = ((lx / range) * (ve[0] - r.vx)) + ((ly / range) * (ve[1] - r.vy)) + ((lz / range) * (ve[2] - r.vz));
        const double bcFi = kFL1 * ((rate - ve[3]) / kC + r.clkd) / dopplerSign;
        const double pr = range - kC * r.clkb + x[3];
        const double bcTx = rxTime - pr / kC;
        const double frac = bcTx - cpRefTOW[i] - ((cpEnd[i] - cpRef[i]) * kTCA);
        const double bcRc = frac * kFCA;
        const double bcFc = kFCA + (dopplerSign * kFCA / kFL1) * bcFi + (bcRc - rcEnd[i]) / T;
        fi[i] = bcFi;
        fc[i] = bcFc;
        if (TimeUpdate(i, x, rxTime)) return -1;
    }
    rxTime += T;
    GridPrep(In<double>(13), In<double>(11));
    return 0;
}

}  // namespace dsp
