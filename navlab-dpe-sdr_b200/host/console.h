// console.h -- the command shell in front of FlowMgr.  Same command set and the same
// "upper-case letters are mandatory" prefix matching as the reference's console
// (cudarecv/cudarecv/src/cmdFlow.cpp:21-31, cmdCommon.cpp:23-26, console/src/cmdParser.cpp:28-113):
//   NEWFlow <type> [alias]      LOADFlow <flow> [file]    STARTFlow <flow>     STOPFlow <flow>
//   DELFlow <flow>              SETParam <flow> <module> <param> <value>       PRINTport <flow> <module> <port>
//   ADDAlias <alias> <flow>     ACTAlias                  ACTFlow              LSFlow
//   WAITFlow <flow> [seconds]   (extension: block until the flow has ended)
//   Quit [-f]    HIStory    HELp    DOfile <file>
// Line based (stdin or dofile); the raw-tty editor of the reference is UI, not part of the path.
#ifndef DPE_HOST_CONSOLE_H_
#define DPE_HOST_CONSOLE_H_

#include <iosfwd>
#include <string>
#include <vector>
#include "flowmgr.h"

namespace console {

class Shell {
  public:
    explicit Shell(dsp::FlowMgr* mgr) : mgr_(mgr) {}
    /** Executes one command line.  Returns 0 ok, -1 error, 1 quit requested. */
    int execOneCmd(const std::string& line);
    int run(std::istream& in, bool prompt);
    const std::vector<std::string>& history() const { return history_; }

  private:
    dsp::FlowMgr* mgr_;
    std::vector<std::string> history_;
    int depth_ = 0;
};

}  // namespace console
#endif
