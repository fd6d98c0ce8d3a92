// dpeflow.cpp -- the DPE flow graph as data: seven modules, their default parameters and the
// port-to-port wiring of the reference's flow (cudarecv/dsp/src/dpeflow.cpp:56-62 modules,
// :67-133 parameters, :140-213 connections).  Module, port and parameter names are the interface
// and therefore identical; everything else is table-driven here.
#include "dpeflow.h"
#include <cstdlib>
#include <ctime>
#include <iostream>
#include <string>
#include "modules.h"

namespace dsp {

namespace {

struct Wire { const char* srcMod; const char* srcPort; const char* dstMod; const char* dstPort; };

// source module -> destination module, grouped by source
const Wire kWires[] = {
    // DPInit: start byte, initial state / covariance, handoff channel parameters, ephemerides
    {"DPInit", "StartByte", "SampleBlock", "StartByte"},
    {"DPInit", "InitX", "cuEKF", "InitX"},
    {"DPInit", "InitP", "cuEKF", "InitP"},
    {"DPInit", "InitK", "cuEKF", "InitK"},
    {"DPInit", "InitEph", "cuChanMgr", "InitEph"},
    {"DPInit", "InitPRN", "cuChanMgr", "InitPRN"},
    {"DPInit", "InitCodePhase", "cuChanMgr", "InitCodePhase"},
    {"DPInit", "InitCarrierPhase", "cuChanMgr", "InitCarrierPhase"},
    {"DPInit", "InitCodeFrequency", "cuChanMgr", "InitCodeFrequency"},
    {"DPInit", "InitCarrierFrequency", "cuChanMgr", "InitCarrierFrequency"},
    {"DPInit", "InitElapsedCodePeriods", "cuChanMgr", "InitElapsedCodePeriods"},
    {"DPInit", "InitReferenceCodePeriods", "cuChanMgr", "InitReferenceCodePeriods"},
    {"DPInit", "InitCPRefTOW", "cuChanMgr", "InitCPRefTOW"},
    {"DPInit", "InitRXTime", "cuChanMgr", "InitRXTime"},
    // SampleBlock: the 20 ms block and its rate / length
    {"SampleBlock", "Samples", "BatchCorrScores", "Samples"},
    {"SampleBlock", "SamplingFrequency", "BatchCorrScores", "SamplingFrequency"},
    {"SampleBlock", "SampleLength", "BatchCorrScores", "SampleLength"},
    {"SampleBlock", "SamplingFrequency", "BatchCorrManifold", "SamplingFrequency"},
    {"SampleBlock", "SampleLength", "BatchCorrManifold", "SampleLength"},
    {"SampleBlock", "SampleLength", "cuChanMgr", "SampleLength"},
    // BatchCorrScores: code correlogram and carrier spectrum
    {"BatchCorrScores", "CodeScores", "BatchCorrManifold", "CodeScores"},
    {"BatchCorrScores", "CarrScores", "BatchCorrManifold", "CarrScores"},
    {"BatchCorrScores", "NumFFTPoints", "BatchCorrManifold", "NumFFTPoints"},
    // cuChanMgr -> BatchCorrScores: parameters referenced to the START of the block
    {"cuChanMgr", "CodePhaseStart", "BatchCorrScores", "CodePhaseStart"},
    {"cuChanMgr", "CodeFrequency", "BatchCorrScores", "CodeFrequency"},
    {"cuChanMgr", "CarrierPhaseStart", "BatchCorrScores", "CarrierPhaseStart"},
    {"cuChanMgr", "CarrierFrequency", "BatchCorrScores", "CarrierFrequency"},
    {"cuChanMgr", "cpReference", "BatchCorrScores", "cpReference"},
    {"cuChanMgr", "cpElapsedStart", "BatchCorrScores", "cpElapsedStart"},
    {"cuChanMgr", "DopplerSign", "BatchCorrScores", "DopplerSign"},
    {"cuChanMgr", "ValidPRNs", "BatchCorrScores", "ValidPRNs"},
    // cuChanMgr -> BatchCorrManifold: parameters referenced to the END of the block (note the renames)
    {"cuChanMgr", "CodeFrequency", "BatchCorrManifold", "CodeFrequency"},
    {"cuChanMgr", "CarrierFrequency", "BatchCorrManifold", "CarrierFrequency"},
    {"cuChanMgr", "rxTime", "BatchCorrManifold", "rxTime"},
    {"cuChanMgr", "txTime", "BatchCorrManifold", "txTime"},
    {"cuChanMgr", "DopplerSign", "BatchCorrManifold", "DopplerSign"},
    {"cuChanMgr", "SatStates", "BatchCorrManifold", "SatStates"},
    {"cuChanMgr", "ENU2ECEFMat", "BatchCorrManifold", "ENU2ECEFMat"},
    {"cuChanMgr", "SatStatesOld", "BatchCorrManifold", "SatStatesOld"},
    {"cuChanMgr", "CodePhaseEnd", "BatchCorrManifold", "CodePhase"},
    {"cuChanMgr", "CarrierPhaseEnd", "BatchCorrManifold", "CarrierPhase"},
    {"cuChanMgr", "cpRefTOW", "BatchCorrManifold", "cpRefTOW"},
    {"cuChanMgr", "cpRef", "BatchCorrManifold", "cpRef"},
    {"cuChanMgr", "cpElapsedEnd", "BatchCorrManifold", "cpElapsedEnd"},
    // BatchCorrManifold: measurement to the filter, time grid to the channel manager
    {"BatchCorrManifold", "zVal", "cuEKF", "zVal"},
    {"BatchCorrManifold", "RVal", "cuEKF", "RVal"},
    {"BatchCorrManifold", "TimeGrid", "cuChanMgr", "TimeGrid"},
    // cuEKF: the estimate (grid centre for the next epoch, channel update, log)
    {"cuEKF", "xCurrk1k1", "cuChanMgr", "xCurrk1k1"},
    {"cuEKF", "xCurrkk1", "cuChanMgr", "xCurrkk1"},
    {"cuEKF", "xCurrkk1", "BatchCorrManifold", "xCurrkk1"},
    {"cuEKF", "xCurrk1k1", "XECEFLogger", "Data"},
};

}  // namespace

int DPEFlow::LoadFlow(const char*) {
    if (!Mods.empty()) { std::cerr << "[DPEFlow] already loaded" << std::endl; return -1; }
    Mods = {new DPInit, new SampleBlock, new BatchCorrScores, new BatchCorrManifold, new cuEKF, new cuChanMgr,
            new DataLogger("XECEFLogger")};

    // default parameters (the reference's values); paths are expected to be overridden with setparam
    const char* home = std::getenv("HOME");
    const std::string demo = std::string(home ? home : ".") + "/Desktop/demofiles/";
    char stamp[80];
    std::time_t now = std::time(nullptr);
    std::strftime(stamp, sizeof(stamp), "%d-%m-%Y_%H-%M-%S", std::localtime(&now));
    const std::string out = "/home/ubuntu/output/inves/" + std::string(stamp);
    const double T = 0.02;       // the reference: "THINGS WILL BREAK IF T IS CHANGED" -- here S is 64-bit, T is free

    int bad = 0;
    bad |= SetModParam("SampleBlock", "SamplingFrequency", 2.5e6);
    bad |= SetModParam("SampleBlock", "SampleLength", T);
    bad |= SetModParam("SampleBlock", "RunLive", false);
    bad |= SetModParam("SampleBlock", "Filename", (demo + "static_opensky_20180705_190000_usrp6_2500kHz.dat").c_str());
    bad |= SetModParam("DPInit", "HandoffFilename", (demo + "handoff_params_usrp6.csv").c_str());
    bad |= SetModParam("DPInit", "RINEXFilename", (demo + "nist1860.18n").c_str());
    for (const char* k : {"InitDeltaX", "InitDeltaY", "InitDeltaZ", "InitDeltaT"}) bad |= SetModParam("DPInit", k, 0.0f);
    bad |= SetModParam("cuEKF", "SampleLength", T);
    bad |= SetModParam("cuEKF", "EnableEKF", false);
    bad |= SetModParam("BatchCorrManifold", "PosGridDimSize", 25);
    bad |= SetModParam("BatchCorrManifold", "VelGridDimSize", 25);
    bad |= SetModParam("BatchCorrManifold", "GridDimSpacing", 1.0f);
    bad |= SetModParam("BatchCorrManifold", "GridType", 0);            // ManifoldGridTypes::Uniform
    bad |= SetModParam("BatchCorrManifold", "LPower", 1);
    bad |= SetModParam("BatchCorrManifold", "GridLogFileName", (out + "-Grid.csv").c_str());
    bad |= SetModParam("BatchCorrManifold", "LoadPosGridFilename", (demo + "rngrid3.csv").c_str());
    bad |= SetModParam("cuChanMgr", "DopplerSign", 1);
    bad |= SetModParam("XECEFLogger", "Filename", (out + "-XFile.csv").c_str());
    bad |= SetModParam("XECEFLogger", "CSV", true);
    if (bad) { std::cerr << "[DPEFlow] default parameters rejected" << std::endl; return -1; }

    for (const Wire& w : kWires)
        if (ConnectPort(w.srcMod, w.srcPort, w.dstMod, w.dstPort)) {
            std::cerr << "[DPEFlow] cannot connect " << w.srcMod << "." << w.srcPort << " -> " << w.dstMod << "."
                      << w.dstPort << std::endl;
            return -1;
        }
    std::clog << "[DPEFlow] Completed LoadFlow." << std::endl;
    return 0;
}

}  // namespace dsp
