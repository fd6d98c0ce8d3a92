#include "console.h"
#include <chrono>
#include <cctype>
#include <fstream>
#include <iostream>
#include <sstream>
#include <thread>

namespace console {

struct Cmd { const char* name; const char* help; };
static const Cmd kCmds[] = {
    {"NEWFlow", "NEWFlow <type> [alias]: create a flow (types: ACTFlow)"},
    {"LOADFlow", "LOADFlow <flow> [file]: build the module graph"},
    {"STARTFlow", "STARTFlow <flow>: start the flow thread"},
    {"STOPFlow", "STOPFlow <flow>"},
    {"DELFlow", "DELFlow <flow>"},
    {"SETParam", "SETParam <flow> <module> <param> <value>: only between LOADFlow and STARTFlow"},
    {"PRINTport", "PRINTport <flow> <module> <port>"},
    {"ADDAlias", "ADDAlias <alias> <flow index>"},
    {"ACTAlias", "ACTAlias: list aliases"},
    {"ACTFlow", "ACTFlow: list flow types"},
    {"LSFlow", "LSFlow: list flows"},
    {"WAITFlow", "WAITFlow <flow> [seconds]: wait until the flow has ended"},
    {"Quit", "Quit [-f]"},
    {"HIStory", "HIStory"},
    {"HELp", "HELp"},
    {"DOfile", "DOfile <file>: run commands from a file"},
};

// upper-case letters of `name` are the mandatory prefix; the rest may be abbreviated
static bool matches(const char* name, const std::string& word) {
    size_t mand = 0;
    while (name[mand] && std::isupper((unsigned char)name[mand])) ++mand;
    const size_t n = std::char_traits<char>::length(name);
    if (word.size() < mand || word.size() > n) return false;
    for (size_t i = 0; i < word.size(); ++i)
        if (std::toupper((unsigned char)word[i]) != std::toupper((unsigned char)name[i])) return false;
    return true;
}

static std::vector<std::string> tokens(const std::string& line) {
    std::vector<std::string> t;
    std::string cur;
    bool quoted = false;
    for (size_t i = 0; i < line.size(); ++i) {
        const char c = line[i];
        if (c == '"') { quoted = !quoted; cur += c; continue; }
        if (!quoted && (c == '#')) break;
        if (!quoted && std::isspace((unsigned char)c)) { if (!cur.empty()) { t.push_back(cur); cur.clear(); } continue; }
        cur += c;
    }
    if (!cur.empty()) t.push_back(cur);
    return t;
}

int Shell::execOneCmd(const std::string& line) {
    const std::vector<std::string> a = tokens(line);
    if (a.empty()) return 0;
    history_.push_back(line);
    const Cmd* cmd = nullptr;
    for (const Cmd& c : kCmds)
        if (matches(c.name, a[0])) { cmd = &c; break; }
    if (!cmd) { std::cerr << "Illegal command!! (" << a[0] << ")" << std::endl; return -1; }
    const std::string n = cmd->name;
    auto need = [&](size_t k) { if (a.size() < k + 1) { std::cerr << "Error: missing option!! usage: " << cmd->help << std::endl; return false; } return true; };
    if (n == "NEWFlow") return need(1) ? mgr_->createFlow(a[1], a.size() > 2 ? a[2] : "") : -1;
    if (n == "LOADFlow") return need(1) ? mgr_->loadFlow(a[1], a.size() > 2 ? a[2].c_str() : nullptr) : -1;
    if (n == "STARTFlow") return need(1) ? mgr_->startFlow(a[1]) : -1;
    if (n == "STOPFlow") return need(1) ? mgr_->stopFlow(a[1]) : -1;
    if (n == "DELFlow") return need(1) ? mgr_->destroyFlow(a[1]) : -1;
    if (n == "SETParam") return need(4) ? mgr_->setParam(a[1], a[2], a[3], a[4]) : -1;
    if (n == "PRINTport") return need(3) ? mgr_->listOutput(a[1], a[2], a[3]) : -1;
    if (n == "ADDAlias") {
        if (!need(2)) return -1;
        const size_t idx = mgr_->getFlowIdx(a[2]);
        if (idx == dsp::FlowMgr::NPOS) return -1;
        mgr_->addAlias(a[1], idx);
        return 0;
    }
    if (n == "ACTAlias") { mgr_->listAlias(); return 0; }
    if (n == "ACTFlow") { mgr_->flowType(); return 0; }
    if (n == "LSFlow") { mgr_->listFlow(); return 0; }
    if (n == "WAITFlow") {
        if (!need(1)) return -1;
        dsp::Flow* f = mgr_->getFlowPtr(a[1]);
        if (!f) return -1;
        const double limit = a.size() > 2 ? std::atof(a[2].c_str()) : 3600.0;
        const auto t0 = std::chrono::steady_clock::now();
        while (!f->CheckFlowState() && std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < limit)
            std::this_thread::sleep_for(std::chrono::milliseconds(5));
        return f->CheckFlowState() ? 0 : -1;
    }
    if (n == "Quit") { mgr_->EmergencyStop(); return 1; }
    if (n == "HIStory") { for (size_t i = 0; i < history_.size(); ++i) std::cout << i << ": " << history_[i] << std::endl; return 0; }
    if (n == "HELp") { for (const Cmd& c : kCmds) std::cout << c.help << std::endl; return 0; }
    if (n == "DOfile") {
        if (!need(1)) return -1;
        if (depth_ > 8) { std::cerr << "dofile nesting too deep" << std::endl; return -1; }
        std::ifstream f(a[1].c_str());
        if (!f) { std::cerr << "Error: cannot open file \"" << a[1] << "\"!!" << std::endl; return -1; }
        ++depth_;
        const int rc = run(f, false);
        --depth_;
        return rc == 1 ? 1 : rc;
    }
    return -1;
}

int Shell::run(std::istream& in, bool prompt) {
    std::string line;
    int last = 0;
    while (true) {
        if (prompt) std::cout << "cudarecv> " << std::flush;
        if (!std::getline(in, line)) break;
        last = execOneCmd(line);
        if (last == 1) return 1;
    }
    return last;
}

}  // namespace console
