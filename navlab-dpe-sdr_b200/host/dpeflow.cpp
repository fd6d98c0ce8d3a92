#include "dpeflow.h"
#include <cstdlib>
#include <ctime>
#include <iostream>
#include <string>
#include "modules.h"

#define errCheck(stmt) do { if ((stmt) != 0) { std::cerr << "[DPEFlow] " #stmt " failed" << std::endl; return -1; } } while (0)

namespace dsp {

int DPEFlow::LoadFlow(const char*) {
    if (!Mods.empty()) { std::cerr << "[DPEFlow] already loaded" << std::endl; return -1; }
    const char* home = std::getenv("HOME");
    const std::string demo = std::string(home ? home : ".") + "/Desktop/demofiles/";

    Mods.reserve(7);
    Mods.push_back(new DPInit);
    Mods.push_back(new SampleBlock);
    Mods.push_back(new BatchCorrScores);
    Mods.push_back(new BatchCorrManifold);
    Mods.push_back(new cuEKF);
    Mods.push_back(new cuChanMgr);
    Mods.push_back(new DataLogger("XECEFLogger"));

    // parameters: the reference's values (dpeflow.cpp:67-133)
    errCheck(SetModParam("SampleBlock", "SamplingFrequency", 2.5e6));
    errCheck(SetModParam("SampleBlock", "RunLive", false));
    errCheck(SetModParam("SampleBlock", "Filename", (demo + "static_opensky_20180705_190000_usrp6_2500kHz.dat").c_str()));
    errCheck(SetModParam("DPInit", "HandoffFilename", (demo + "handoff_params_usrp6.csv").c_str()));
    errCheck(SetModParam("DPInit", "RINEXFilename", (demo + "nist1860.18n").c_str()));
    errCheck(SetModParam("DPInit", "InitDeltaX", 0.0f));
    errCheck(SetModParam("DPInit", "InitDeltaY", 0.0f));
    errCheck(SetModParam("DPInit", "InitDeltaZ", 0.0f));
    errCheck(SetModParam("DPInit", "InitDeltaT", 0.0f));
    const double T = 0.02;
    errCheck(SetModParam("SampleBlock", "SampleLength", T));
    errCheck(SetModParam("cuEKF", "SampleLength", T));
    errCheck(SetModParam("BatchCorrManifold", "PosGridDimSize", 25));
    errCheck(SetModParam("BatchCorrManifold", "VelGridDimSize", 25));
    errCheck(SetModParam("BatchCorrManifold", "GridDimSpacing", 1.0f));
    errCheck(SetModParam("BatchCorrManifold", "GridType", 0));          // ManifoldGridTypes::Uniform
    errCheck(SetModParam("BatchCorrManifold", "LPower", 1));
    errCheck(SetModParam("cuChanMgr", "DopplerSign", 1));
    errCheck(SetModParam("cuEKF", "EnableEKF", false));

    char stamp[80];
    std::time_t now = std::time(nullptr);
    std::strftime(stamp, sizeof(stamp), "%d-%m-%Y_%H-%M-%S", std::localtime(&now));
    const std::string prefix = "/home/ubuntu/output/inves/" + std::string(stamp);
    errCheck(SetModParam("XECEFLogger", "Filename", (prefix + "-XFile.csv").c_str()));
    errCheck(SetModParam("XECEFLogger", "CSV", true));
    errCheck(SetModParam("BatchCorrManifold", "GridLogFileName", (prefix + "-Grid.csv").c_str()));
    errCheck(SetModParam("BatchCorrManifold", "LoadPosGridFilename", (demo + "rngrid3.csv").c_str()));

    // port connections, grouped by source module (dpeflow.cpp:140-213)
    errCheck(ConnectPort("DPInit", "StartByte", "SampleBlock", "StartByte"));
    errCheck(ConnectPort("DPInit", "InitX", "cuEKF", "InitX"));
    errCheck(ConnectPort("DPInit", "InitP", "cuEKF", "InitP"));
    errCheck(ConnectPort("DPInit", "InitK", "cuEKF", "InitK"));
    errCheck(ConnectPort("DPInit", "InitEph", "cuChanMgr", "InitEph"));
    errCheck(ConnectPort("DPInit", "InitPRN", "cuChanMgr", "InitPRN"));
    errCheck(ConnectPort("DPInit", "InitCodePhase", "cuChanMgr", "InitCodePhase"));
    errCheck(ConnectPort("DPInit", "InitCarrierPhase", "cuChanMgr", "InitCarrierPhase"));
    errCheck(ConnectPort("DPInit", "InitCodeFrequency", "cuChanMgr", "InitCodeFrequency"));
    errCheck(ConnectPort("DPInit", "InitCarrierFrequency", "cuChanMgr", "InitCarrierFrequency"));
    errCheck(ConnectPort("DPInit", "InitElapsedCodePeriods", "cuChanMgr", "InitElapsedCodePeriods"));
    errCheck(ConnectPort("DPInit", "InitReferenceCodePeriods", "cuChanMgr", "InitReferenceCodePeriods"));
    errCheck(ConnectPort("DPInit", "InitCPRefTOW", "cuChanMgr", "InitCPRefTOW"));
    errCheck(ConnectPort("DPInit", "InitRXTime", "cuChanMgr", "InitRXTime"));

    errCheck(ConnectPort("SampleBlock", "Samples", "BatchCorrScores", "Samples"));
    errCheck(ConnectPort("SampleBlock", "SamplingFrequency", "BatchCorrScores", "SamplingFrequency"));
    errCheck(ConnectPort("SampleBlock", "SampleLength", "BatchCorrScores", "SampleLength"));
    errCheck(ConnectPort("SampleBlock", "SamplingFrequency", "BatchCorrManifold", "SamplingFrequency"));
    errCheck(ConnectPort("SampleBlock", "SampleLength", "BatchCorrManifold", "SampleLength"));
    errCheck(ConnectPort("SampleBlock", "SampleLength", "cuChanMgr", "SampleLength"));

    errCheck(ConnectPort("BatchCorrScores", "CodeScores", "BatchCorrManifold", "CodeScores"));
    errCheck(ConnectPort("BatchCorrScores", "CarrScores", "BatchCorrManifold", "CarrScores"));
    errCheck(ConnectPort("BatchCorrScores", "NumFFTPoints", "BatchCorrManifold", "NumFFTPoints"));

    errCheck(ConnectPort("cuChanMgr", "CodePhaseStart", "BatchCorrScores", "CodePhaseStart"));
    errCheck(ConnectPort("cuChanMgr", "CodeFrequency", "BatchCorrScores", "CodeFrequency"));
    errCheck(ConnectPort("cuChanMgr", "CarrierPhaseStart", "BatchCorrScores", "CarrierPhaseStart"));
    errCheck(ConnectPort("cuChanMgr", "CarrierFrequency", "BatchCorrScores", "CarrierFrequency"));
    errCheck(ConnectPort("cuChanMgr", "cpReference", "BatchCorrScores", "cpReference"));
    errCheck(ConnectPort("cuChanMgr", "cpElapsedStart", "BatchCorrScores", "cpElapsedStart"));
    errCheck(ConnectPort("cuChanMgr", "DopplerSign", "BatchCorrScores", "DopplerSign"));
    errCheck(ConnectPort("cuChanMgr", "ValidPRNs", "BatchCorrScores", "ValidPRNs"));
    errCheck(ConnectPort("cuChanMgr", "CodeFrequency", "BatchCorrManifold", "CodeFrequency"));
    errCheck(ConnectPort("cuChanMgr", "CarrierFrequency", "BatchCorrManifold", "CarrierFrequency"));
    errCheck(ConnectPort("cuChanMgr", "rxTime", "BatchCorrManifold", "rxTime"));
    errCheck(ConnectPort("cuChanMgr", "txTime", "BatchCorrManifold", "txTime"));
    errCheck(ConnectPort("cuChanMgr", "DopplerSign", "BatchCorrManifold", "DopplerSign"));
    errCheck(ConnectPort("cuChanMgr", "SatStates", "BatchCorrManifold", "SatStates"));
    errCheck(ConnectPort("cuChanMgr", "ENU2ECEFMat", "BatchCorrManifold", "ENU2ECEFMat"));
    errCheck(ConnectPort("cuChanMgr", "SatStatesOld", "BatchCorrManifold", "SatStatesOld"));
    errCheck(ConnectPort("cuChanMgr", "CodePhaseEnd", "BatchCorrManifold", "CodePhase"));
    errCheck(ConnectPort("cuChanMgr", "CarrierPhaseEnd", "BatchCorrManifold", "CarrierPhase"));
    errCheck(ConnectPort("cuChanMgr", "cpRefTOW", "BatchCorrManifold", "cpRefTOW"));
    errCheck(ConnectPort("cuChanMgr", "cpRef", "BatchCorrManifold", "cpRef"));
    errCheck(ConnectPort("cuChanMgr", "cpElapsedEnd", "BatchCorrManifold", "cpElapsedEnd"));

    errCheck(ConnectPort("BatchCorrManifold", "zVal", "cuEKF", "zVal"));
    errCheck(ConnectPort("BatchCorrManifold", "RVal", "cuEKF", "RVal"));
    errCheck(ConnectPort("BatchCorrManifold", "TimeGrid", "cuChanMgr", "TimeGrid"));

    errCheck(ConnectPort("cuEKF", "xCurrk1k1", "cuChanMgr", "xCurrk1k1"));
    errCheck(ConnectPort("cuEKF", "xCurrkk1", "cuChanMgr", "xCurrkk1"));
    errCheck(ConnectPort("cuEKF", "xCurrkk1", "BatchCorrManifold", "xCurrkk1"));
    errCheck(ConnectPort("cuEKF", "xCurrk1k1", "XECEFLogger", "Data"));

    std::clog << "[DPEFlow] Completed LoadFlow." << std::endl;
    return 0;
}

}  // namespace dsp
