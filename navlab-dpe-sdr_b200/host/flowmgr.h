// flowmgr.h -- registry of flow types and live flows behind the console commands
// (cudarecv/dsp/inc/flowmgr.h:41-93, dsp/src/flowmgr.cpp:42-330).
#ifndef DPE_HOST_FLOWMGR_H_
#define DPE_HOST_FLOWMGR_H_

#include <map>
#include <string>
#include <vector>
#include "flow.h"

namespace dsp {

class FlowMgr {
  public:
    static const size_t NPOS = (size_t)-1;
    FlowMgr();
    ~FlowMgr();
    /** `type` is matched case-insensitively on its mandatory prefix ("DPE", 3 letters). */
    int createFlow(const std::string& type, const std::string& alias = "");
    int loadFlow(const std::string& key, const char* filename = nullptr) const;
    int startFlow(const std::string& key) const;
    int stopFlow(const std::string& key) const;
    int destroyFlow(const std::string& key);
    /** value literal typing as in flowmgr.cpp:215-261: \xHH char, "..." string, true/false bool,
     *  contains . e or f -> float, else int; plus (extension) a trailing `d` or `lf` for a double,
     *  which the reference's console cannot express. */
    int setParam(const std::string& key, const std::string& mod, const std::string& param, const std::string& value) const;
    int listOutput(const std::string& key, const std::string& mod, const std::string& port) const;
    void addAlias(const std::string& alias, const size_t& idx);
    void listAlias() const;
    void listFlow() const;
    void flowType() const;
    size_t getFlowIdx(const std::string& key) const;
    Flow* getFlowPtr(const std::string& key) const;
    size_t EmergencyStop() const;

  private:
    struct Registrar { std::string name, desc; size_t mandatory; Flow* (*create)(); };
    std::vector<Registrar> _regis;
    std::vector<std::pair<Flow*, std::string>> _flowlist;
    std::map<std::string, size_t> _alias;
};

}  // namespace dsp
#endif
