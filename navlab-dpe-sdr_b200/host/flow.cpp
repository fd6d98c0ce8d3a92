#include "flow.h"
#include <chrono>
#include <vector>
#include <cstdlib>
#include <iostream>
#include "../../include/dpe_b200.h"

namespace dsp {

Flow::~Flow() {
    if (started) Stop();
    for (size_t i = 0; i < Mods.size(); ++i) delete Mods[i];
    if (cuStream) dpe_stream_destroy(cuStream);
}

int Flow::LoadFlow(const char*) { return -1; }

int Flow::GetModID(const std::string& name) const {
    for (size_t i = 0; i < Mods.size(); ++i)
        if (Mods[i]->GetModuleName() == name) return (int)i;
    std::cerr << "[Flow] no module named " << name << std::endl;
    return -1;
}

Module* Flow::GetModule(const std::string& name) const {
    const int id = GetModID(name);
    return id < 0 ? nullptr : Mods[id];
}

int Flow::ConnectPort(const std::string& srcMod, const char* srcPort, const std::string& dstMod, const char* dstPort) {
    const int s = GetModID(srcMod), d = GetModID(dstMod);
    if (s < 0 || d < 0) return -1;
    const int sp = Mods[s]->GetOutputID(srcPort), dp = Mods[d]->GetInputID(dstPort);
    if (sp < 0 || dp < 0) return -1;
    Port* p = nullptr;
    if (Mods[s]->GetOutput((unsigned char)sp, &p)) return -1;
    return Mods[d]->SetInput((unsigned char)dp, p);
}

int Flow::GetOutput(const std::string& modName, const std::string& portName, Port** out) const {
    const int m = GetModID(modName);
    if (m < 0) return -1;
    const int p = Mods[m]->GetOutputID(portName.c_str());
    if (p < 0) return -1;
    return Mods[m]->GetOutput((unsigned char)p, out);
}

#define DPE_SETMODPARAM(T)                                                                     \
    int Flow::SetModParam(const std::string& mod, const std::string& key, const T val) {        \
        const int m = GetModID(mod);                                                           \
        return m < 0 ? -1 : Mods[m]->SetParam(key, val);                                       \
    }
DPE_SETMODPARAM(int)
DPE_SETMODPARAM(char)
DPE_SETMODPARAM(float)
DPE_SETMODPARAM(double)
DPE_SETMODPARAM(bool)
int Flow::SetModParam(const std::string& mod, const std::string& key, const char* str) {
    const int m = GetModID(mod);
    return m < 0 ? -1 : Mods[m]->SetParam(key, str);
}

int Flow::StartModules() {
    if (Mods.empty()) { std::cerr << "[Flow] Flow not loaded." << std::endl; return -1; }
    if (!cuStream && dpe_stream_create(&cuStream)) {
        std::cerr << "[Flow] cannot create a CUDA stream: " << dpe_last_error() << std::endl;
        return -1;
    }
    for (size_t i = 0; i < Mods.size(); ++i)
        if (Mods[i]->Start((void*)&cuStream)) {
            std::cerr << "[Flow] Unable to start module " << Mods[i]->GetModuleName() << std::endl;
            for (size_t j = 0; j <= i; ++j) Mods[j]->Stop();     // the one that failed too: it may hold what it got half way
            return -1;
        }
    started = true;
    return 0;
}

int Flow::Start() {
    std::clog << "[FLOW] Starting." << std::endl;
    if (StartModules()) return -1;
    FlowDone = false;
    KeepRunning = true;
    thread = std::thread(&Flow::FlowThread, this, -1L);
    std::clog << "[FLOW] Started." << std::endl;
    return 0;
}

int Flow::RunBlocking(long maxEpochs) {
    if (StartModules()) return -1;
    FlowDone = false;
    KeepRunning = true;
    FlowThread(maxEpochs);
    started = false;
    return 0;
}

int Flow::Stop() {
    if (!started) { std::clog << "[Flow] Stop: Flow wasn't running." << std::endl; return 0; }
    std::clog << "[Flow] Stopping Flow." << std::endl;
    KeepRunning = false;
    if (thread.joinable()) thread.join();
    started = false;
    return 0;
}

void Flow::FlowThread(long maxEpochs) {
    using clk = std::chrono::steady_clock;
    const clk::time_point t_begin = clk::now();
    clk::time_point t_iter = t_begin;
    double total_us = 0;
    stats = FlowStats();
    stats.min_us = 1e300;
    // DPE_FLOW_PROFILE=1: host time spent inside each module's Update (an asynchronous module only enqueues; the module
    // that fetches a result pays for the GPU work it waits on)
    const char* pf = getenv("DPE_FLOW_PROFILE");
    const bool profile = pf && pf[0] == '1';
    std::vector<double> mod_us(Mods.size() + 1, 0.0);
    while (KeepRunning && (maxEpochs < 0 || (long)stats.runCount < maxEpochs)) {
        bool failed = false;
        for (size_t i = 0; i < Mods.size(); ++i) {
            const clk::time_point t_mod = profile ? clk::now() : clk::time_point();
            const int rc_mod = Mods[i]->Update((void*)&cuStream);
            if (profile) mod_us[i] += std::chrono::duration<double, std::micro>(clk::now() - t_mod).count();
            if (rc_mod) {
                std::cerr << "[Flow] " << Mods[i]->GetModuleName() << "->Update() Failed. \nStopping Flow." << std::endl;
                failed = true;
                break;
            }
            if (i == 0) t_iter = clk::now();       // the reference starts its timer after module 0 (flow.cu:132-135)
        }
        if (failed) break;
        if (dpe_stream_sync(cuStream)) { std::cerr << "[Flow] stream sync: " << dpe_last_error() << std::endl; break; }
        const double us = std::chrono::duration<double, std::micro>(clk::now() - t_iter).count();
        stats.runCount++;
        total_us += us;
        if (us > stats.max_us) { stats.max_us = us; stats.maxCount = stats.runCount; }
        if (us < stats.min_us) { stats.min_us = us; stats.minCount = stats.runCount; }
    }
    KeepRunning = false;
    std::clog << "[Flow] Flow Stopped." << std::endl << "[Flow] Stopping Modules." << std::endl;
    for (size_t i = 0; i < Mods.size(); ++i) Mods[i]->Stop();
    stats.avg_us = stats.runCount ? total_us / stats.runCount : 0;
    stats.total_s = std::chrono::duration<double>(clk::now() - t_begin).count();
    if (!stats.runCount) stats.min_us = 0;
    std::clog << "[Flow] runCount = " << stats.runCount << std::endl
              << "[Flow] Average 20ms block duration = " << stats.avg_us << " us" << std::endl
              << "[Flow] Max block duration = " << stats.max_us << " us, run count = " << stats.maxCount << std::endl
              << "[Flow] Min block duration = " << stats.min_us << " us, run count = " << stats.minCount << std::endl
              << "[Flow] Total time = " << stats.total_s << " seconds." << std::endl;
    if (profile && stats.runCount)
        for (size_t i = 0; i < Mods.size(); ++i)
            std::clog << "[Flow] profile: " << Mods[i]->GetModuleName() << " " << mod_us[i] / stats.runCount << " us per epoch" << std::endl;
    FlowDone = true;
}

}  // namespace dsp
