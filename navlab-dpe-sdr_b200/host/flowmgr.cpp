#include "flowmgr.h"
#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <iostream>
#include "dpeflow.h"

namespace dsp {

static std::string upper(std::string s) {
    for (size_t i = 0; i < s.size(); ++i) s[i] = (char)std::toupper((unsigned char)s[i]);
    return s;
}

/** Command-style prefix match: `cmd` must start with the first `mandatory` letters of `name`
 *  and be a prefix of it (case-insensitive), e.g. "dpe", "DPE". */
static bool prefixMatch(const std::string& name, const std::string& cmd, size_t mandatory) {
    const std::string n = upper(name), c = upper(cmd);
    return c.size() >= mandatory && c.size() <= n.size() && n.compare(0, c.size(), c) == 0;
}

static Flow* makeDPE() { return new DPEFlow; }

FlowMgr::FlowMgr() { _regis.push_back(Registrar{"DPE", "Direct Position Estimation Flow", 3, &makeDPE}); }

FlowMgr::~FlowMgr() {
    for (size_t i = 0; i < _flowlist.size(); ++i) delete _flowlist[i].first;
}

int FlowMgr::createFlow(const std::string& type, const std::string& alias) {
    for (size_t i = 0; i < _regis.size(); ++i)
        if (prefixMatch(_regis[i].name, type, _regis[i].mandatory)) {
            _flowlist.push_back(std::make_pair(_regis[i].create(), _regis[i].name));
            if (!alias.empty()) addAlias(alias, _flowlist.size() - 1);
            std::cout << "Flow " << _flowlist.size() - 1 << " (" << _regis[i].name << ") created." << std::endl;
            return 0;
        }
    std::cerr << "[FlowMgr] unknown flow type: " << type << std::endl;
    return -1;
}

size_t FlowMgr::getFlowIdx(const std::string& key) const {
    std::map<std::string, size_t>::const_iterator it = _alias.find(key);
    if (it != _alias.end()) return it->second < _flowlist.size() ? it->second : NPOS;
    if (key.empty() || !std::all_of(key.begin(), key.end(), [](char c) { return std::isdigit((unsigned char)c); })) return NPOS;
    const size_t idx = (size_t)std::strtoul(key.c_str(), nullptr, 10);
    return (idx < _flowlist.size() && _flowlist[idx].first) ? idx : NPOS;
}

Flow* FlowMgr::getFlowPtr(const std::string& key) const {
    const size_t i = getFlowIdx(key);
    if (i == NPOS) { std::cerr << "[FlowMgr] no such flow: " << key << std::endl; return nullptr; }
    return _flowlist[i].first;
}

int FlowMgr::loadFlow(const std::string& key, const char* filename) const {
    Flow* f = getFlowPtr(key);
    return f ? f->LoadFlow(filename) : -1;
}
int FlowMgr::startFlow(const std::string& key) const { Flow* f = getFlowPtr(key); return f ? f->Start() : -1; }
int FlowMgr::stopFlow(const std::string& key) const { Flow* f = getFlowPtr(key); return f ? f->Stop() : -1; }

int FlowMgr::destroyFlow(const std::string& key) {
    const size_t i = getFlowIdx(key);
    if (i == NPOS) return -1;
    delete _flowlist[i].first;
    _flowlist[i].first = nullptr;
    for (std::map<std::string, size_t>::iterator it = _alias.begin(); it != _alias.end();)
        if (it->second == i) it = _alias.erase(it); else ++it;
    return 0;
}

int FlowMgr::setParam(const std::string& key, const std::string& mod, const std::string& param,
                      const std::string& value) const {
    Flow* f = getFlowPtr(key);
    if (!f || value.empty()) return -1;
    if (value.size() >= 3 && value[0] == '\\' && value[1] == 'x')
        return f->SetModParam(mod, param, (char)std::strtol(value.c_str() + 2, nullptr, 16));
    if (value.size() >= 2 && value[0] == '"' && value[value.size() - 1] == '"')
        return f->SetModParam(mod, param, value.substr(1, value.size() - 2).c_str());
    if (value == "true" || value == "false") return f->SetModParam(mod, param, value == "true");
    const bool dbl = (value.size() > 1 && value[value.size() - 1] == 'd') ||
                     (value.size() > 2 && value.compare(value.size() - 2, 2, "lf") == 0);
    if (dbl) return f->SetModParam(mod, param, std::strtod(value.c_str(), nullptr));
    if (value.find_first_of(".ef") != std::string::npos)
        return f->SetModParam(mod, param, (float)std::strtod(value.c_str(), nullptr));
    return f->SetModParam(mod, param, (int)std::strtol(value.c_str(), nullptr, 10));
}

int FlowMgr::listOutput(const std::string& key, const std::string& mod, const std::string& port) const {
    Flow* f = getFlowPtr(key);
    Port* p = nullptr;
    if (!f || f->GetOutput(mod, port, &p) || !p) return -1;
    std::cout << mod << "." << port << " [" << p->Length << "]";
    if (p->MemLoc != HOST || !p->Data) { std::cout << " (not a readable HOST port)" << std::endl; return 0; }
    std::cout << ":";
    for (int64_t i = 0; i < p->Length && i < 64; ++i) {
        switch (p->Datatype) {
            case DOUBLE_t: std::cout << " " << static_cast<double*>(p->Data)[i]; break;
            case FLOAT_t: std::cout << " " << static_cast<float*>(p->Data)[i]; break;
            case INT_t: std::cout << " " << static_cast<int*>(p->Data)[i]; break;
            case BOOL_t: std::cout << " " << static_cast<bool*>(p->Data)[i]; break;
            case CHAR_t: std::cout << " " << (int)static_cast<char*>(p->Data)[i]; break;
            default: std::cout << " ?"; break;
        }
    }
    std::cout << std::endl;
    return 0;
}

void FlowMgr::addAlias(const std::string& alias, const size_t& idx) { _alias[alias] = idx; }

void FlowMgr::listAlias() const {
    for (std::map<std::string, size_t>::const_iterator it = _alias.begin(); it != _alias.end(); ++it)
        std::cout << it->first << " -> " << it->second << std::endl;
}

void FlowMgr::listFlow() const {
    for (size_t i = 0; i < _flowlist.size(); ++i)
        if (_flowlist[i].first)
            std::cout << i << ": " << _flowlist[i].second << (_flowlist[i].first->CheckFlowState() ? " (done)" : "") << std::endl;
}

void FlowMgr::flowType() const {
    for (size_t i = 0; i < _regis.size(); ++i) std::cout << _regis[i].name << " : " << _regis[i].desc << std::endl;
}

size_t FlowMgr::EmergencyStop() const {
    size_t n = 0;
    for (size_t i = 0; i < _flowlist.size(); ++i)
        if (_flowlist[i].first) { _flowlist[i].first->Stop(); ++n; }
    return n;
}

}  // namespace dsp
