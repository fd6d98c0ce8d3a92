// flow_capi.cpp -- extern "C" handle on the flow mirror (libdpe_flow.so) so the parity tests can
// drive `newflow dpe / setparam / loadflow / startflow` exactly like a console user and read
// HOST ports back.  Declared in include/dpe_flow.h.
#include <cstring>
#include <sstream>
#include "../../include/dpe_flow.h"
#include "console.h"
#include "gnss.h"

struct dpe_shell {
    dsp::FlowMgr mgr;
    console::Shell sh;
    dpe_shell() : sh(&mgr) {}
};

extern "C" {

dpe_shell* dpe_shell_create(void) { return new dpe_shell(); }
void dpe_shell_destroy(dpe_shell* s) { if (s) { s->mgr.EmergencyStop(); delete s; } }
int dpe_shell_exec(dpe_shell* s, const char* line) { return (s && line) ? s->sh.execOneCmd(line) : -1; }

int dpe_shell_run_blocking(dpe_shell* s, const char* flow, long max_epochs) {
    dsp::Flow* f = s ? s->mgr.getFlowPtr(flow) : nullptr;
    return f ? f->RunBlocking(max_epochs) : -1;
}

int dpe_shell_flow_stats(dpe_shell* s, const char* flow, double* out5) {
    dsp::Flow* f = s ? s->mgr.getFlowPtr(flow) : nullptr;
    if (!f || !out5) return -1;
    const dsp::FlowStats& st = f->Stats();
    out5[0] = (double)st.runCount; out5[1] = st.avg_us; out5[2] = st.min_us; out5[3] = st.max_us; out5[4] = st.total_s;
    return 0;
}

long dpe_shell_read_port(dpe_shell* s, const char* flow, const char* mod, const char* port, double* out, long cap) {
    dsp::Flow* f = s ? s->mgr.getFlowPtr(flow) : nullptr;
    dsp::Port* p = nullptr;
    if (!f || f->GetOutput(mod, port, &p) || !p || p->MemLoc != dsp::HOST || !p->Data) return -1;
    long n = (long)p->Length;
    if (p->ValueType == dsp::STATE && p->Datatype == dsp::DOUBLE_t && std::strncmp(p->Name, "SatStates", 9) == 0) n *= 8;
    if (n > cap) n = cap;
    for (long i = 0; i < n; ++i) {
        switch (p->Datatype) {
            case dsp::DOUBLE_t: out[i] = static_cast<double*>(p->Data)[i]; break;
            case dsp::FLOAT_t: out[i] = static_cast<float*>(p->Data)[i]; break;
            case dsp::INT_t: out[i] = static_cast<int*>(p->Data)[i]; break;
            case dsp::CHAR_t: out[i] = static_cast<unsigned char*>(p->Data)[i]; break;
            case dsp::BOOL_t: out[i] = static_cast<bool*>(p->Data)[i]; break;
            default: return -1;
        }
    }
    return n;
}

// host GPS helpers exposed for CPU parity tests against the oracle
int dpe_host_sat_position(const char* rinex, int prn, double tx_time, double* state8) {
    std::vector<gnss::EphSet> nav;
    if (gnss::ReadRinexNav(rinex, &nav)) return -1;
    const gnss::Eph* e = gnss::SelectEph(nav, prn, tx_time);
    gnss::SatState s;
    if (!e || !gnss::SatPosition(*e, tx_time, &s)) return -1;
    std::memcpy(state8, &s, sizeof(s));
    return 0;
}

int dpe_host_make_grid(const int* dims4, const double* spacing4, int grid_type, double* out, long cap) {
    std::vector<double> g, t;
    gnss::MakeGrid(dims4, spacing4, grid_type, &g, &t);
    if ((long)g.size() > cap) return -1;
    std::memcpy(out, g.data(), g.size() * sizeof(double));
    return (int)(g.size() / 4);
}

// handoff CSV (dpinit.cpp:247-400) flattened as doubles:
//   [0] n channels  [1] rxTime  [2] bytes_read  [3] t_oe  [4..11] X_ECEF(8)
//   then 8 rows of n: prn, rc, ri, fc, fi, cp, cp_timestamp, TOW
long dpe_host_read_handoff(const char* path, double* out, long cap) {
    gnss::Handoff h;
    if (gnss::ReadHandoff(path, &h)) return -1;
    const long n = (long)h.prn.size(), need = 12 + 8 * n;
    if (need > cap) return -1;
    out[0] = (double)n; out[1] = h.rxTime; out[2] = (double)h.bytes_read; out[3] = (double)h.t_oe;
    for (int i = 0; i < 8; ++i) out[4 + i] = h.X_ECEF[i];
    double* q = out + 12;
    for (long i = 0; i < n; ++i) {
        q[0 * n + i] = h.prn[i]; q[1 * n + i] = h.rc[i]; q[2 * n + i] = h.ri[i]; q[3 * n + i] = h.fc[i];
        q[4 * n + i] = h.fi[i]; q[5 * n + i] = h.cp[i]; q[6 * n + i] = h.cp_timestamp[i]; q[7 * n + i] = h.TOW[i];
    }
    return need;
}

// grid CSV (`x,y,z,delta_t` per line, batchcorrmanifold.cu:2433-2444): returns the number of candidates
long dpe_host_read_grid(const char* path, double* out, long cap) {
    std::vector<double> g;
    if (gnss::ReadGridCsv(path, &g)) return -1;
    if ((long)g.size() > cap) return -1;
    std::memcpy(out, g.data(), g.size() * sizeof(double));
    return (long)(g.size() / 4);
}

}  // extern "C"
