// flow.h -- a flow = an ordered list of modules stepped by one thread.
// Interface of cudarecv/dsp/inc/flow.h:14-102 (LoadFlow / Start / Stop / CheckFlowState /
// SetModParam / GetOutput / ConnectPort); the thread body is Flow::FlowThread (flow.cu:105-197):
// `for m in Mods: m->Update(&cuStream)` until a module returns non-zero, with the same
// per-epoch timing statistics printed at the end.
#ifndef DPE_HOST_FLOW_H_
#define DPE_HOST_FLOW_H_

#include <atomic>
#include <string>
#include <thread>
#include <vector>
#include "module.h"

namespace dsp {

struct FlowStats {
    unsigned long runCount = 0;
    double avg_us = 0, min_us = 0, max_us = 0, total_s = 0;
    unsigned long minCount = 0, maxCount = 0;
};

class Flow {
    friend class FlowMgr;

  public:
    virtual ~Flow();
    virtual int LoadFlow(const char* filename);
    virtual int Start();
    virtual int Stop();
    virtual bool CheckFlowState() { return FlowDone; }

    int GetOutput(const std::string& modName, const std::string& portName, Port** out) const;
    int SetModParam(const std::string& mod, const std::string& key, const int val);
    int SetModParam(const std::string& mod, const std::string& key, const char val);
    int SetModParam(const std::string& mod, const std::string& key, const float val);
    int SetModParam(const std::string& mod, const std::string& key, const double val);
    int SetModParam(const std::string& mod, const std::string& key, const bool val);
    int SetModParam(const std::string& mod, const std::string& key, const char* str);

    /** Run synchronously on the caller's thread for at most maxEpochs epochs (tests, batch runs). */
    int RunBlocking(long maxEpochs);
    const FlowStats& Stats() const { return stats; }
    Module* GetModule(const std::string& name) const;

  protected:
    std::vector<Module*> Mods;
    std::thread thread;
    std::atomic<bool> KeepRunning{false};
    std::atomic<bool> FlowDone{false};
    bool started = false;
    void* cuStream = nullptr;          // cudaStream_t created through the C ABI; modules get &cuStream
    FlowStats stats;

    void FlowThread(long maxEpochs);
    int StartModules();
    int GetModID(const std::string& name) const;
    int ConnectPort(const std::string& srcMod, const char* srcPort, const std::string& dstMod, const char* dstPort);
};

}  // namespace dsp
#endif
