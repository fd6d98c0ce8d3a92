#include "gnss.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace gnss {

// ---- time helpers (GPS epoch 1980-01-06) -------------------------------------------------
static long long days_from_civil(int y, int m, int d) {     // days since 1970-01-01
    y -= m <= 2;
    const long long era = (y >= 0 ? y : y - 399) / 400;
    const unsigned yoe = (unsigned)(y - era * 400);
    const unsigned doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
    const unsigned doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + (long long)doe - 719468;
}

static void gps_week_tow(int y, int mo, int d, int h, int mi, double s, int* week, double* tow) {
    const long long days = days_from_civil(y, mo, d) - days_from_civil(1980, 1, 6);
    const double secs = (double)days * 86400.0 + h * 3600.0 + mi * 60.0 + s;
    *week = (int)std::floor(secs / 604800.0);
    *tow = secs - (double)*week * 604800.0;
}

static double field(const std::string& line, size_t pos, size_t n) {   // RINEX D-exponent number
    if (line.size() <= pos) return 0.0;
    std::string s = line.substr(pos, n);
    for (size_t i = 0; i < s.size(); ++i)
        if (s[i] == 'D' || s[i] == 'd') s[i] = 'E';
    return std::strtod(s.c_str(), nullptr);
}

// RINEX 2.x GPS navigation message: header up to END OF HEADER, then 8 lines per record
// (PRN / epoch / clock, then 7 broadcast-orbit lines of four 19-character fields).
int ReadRinexNav(const std::string& path, std::vector<EphSet>* out) {
    std::ifstream f(path.c_str());
    if (!f) return -1;
    out->clear();
    std::string line;
    bool header_done = false;
    while (std::getline(f, line))
        if (line.find("END OF HEADER") != std::string::npos) { header_done = true; break; }
    if (!header_done) return -1;
    std::vector<double> d;
    int prn = 0, week = 0;
    double tocs = 0;
    while (std::getline(f, line)) {
        if (line.size() < 22) continue;
        if (d.empty()) {
            prn = (int)field(line, 0, 2);
            int y, mo, dd, h, mi;
            double s;
            if (std::sscanf(line.substr(3, 19).c_str(), "%d %d %d %d %d %lf", &y, &mo, &dd, &h, &mi, &s) != 6) continue;
            y += (y < 80) ? 2000 : (y < 100 ? 1900 : 0);
            gps_week_tow(y, mo, dd, h, mi, s, &week, &tocs);
            for (int j = 0; j < 3; ++j) d.push_back(field(line, 22 + 19 * j, 19));
        } else {
            for (int j = 0; j < 4; ++j) d.push_back(field(line, 3 + 19 * j, 19));
            if (d.size() >= 31) {
                if (prn >= 1 && prn <= 32) {
                    Eph e;
                    e.sat = prn; e.tocs = tocs;
                    e.f0 = d[0]; e.f1 = d[1]; e.f2 = d[2];
                    e.crs = d[4]; e.deln = d[5]; e.M0 = d[6]; e.cuc = d[7]; e.e = d[8]; e.cus = d[9];
                    e.sqrtA = d[10]; e.toes = d[11]; e.cic = d[12]; e.OMG0 = d[13]; e.cis = d[14]; e.i0 = d[15];
                    e.crc = d[16]; e.omg = d[17]; e.OMGd = d[18]; e.idot = d[19]; e.week = (int)d[21]; e.tgd = d[25];
                    e.A = e.sqrtA * e.sqrtA;
                    EphSet* set = nullptr;
                    for (size_t i = 0; i < out->size(); ++i)
                        if ((*out)[i].toes == (double)(long long)e.toes) { set = &(*out)[i]; break; }
                    if (!set) { out->push_back(EphSet()); set = &out->back(); set->toes = e.toes; }
                    set->eph[prn] = e;
                    set->valid[prn] = true;
                }
                d.clear();
            }
        }
    }
    return out->empty() ? -1 : 0;
}

static std::vector<std::string> split_csv(const std::string& line) {
    std::vector<std::string> v;
    std::stringstream ss(line);
    std::string tok;
    while (std::getline(ss, tok, ',')) {
        while (!tok.empty() && (tok.back() == '\r' || tok.back() == '\n' || tok.back() == ' ')) tok.pop_back();
        v.push_back(tok);
    }
    return v;
}

int ReadHandoff(const std::string& path, Handoff* h) {
    std::ifstream f(path.c_str());
    if (!f) return -1;
    *h = Handoff();
    std::string line;
    while (std::getline(f, line)) {
        std::vector<std::string> t = split_csv(line);
        if (t.size() < 2) continue;
        const std::string& key = t[0];
        std::vector<double> v;
        for (size_t i = 1; i < t.size(); ++i)
            if (!t[i].empty()) v.push_back(std::atof(t[i].c_str()));
        if (v.empty()) continue;
        if (key == "rxTime") h->rxTime = v[0];
        else if (key == "X_ECEF") h->X_ECEF = v;
        else if (key == "bytes_read") h->bytes_read = std::atoll(t[1].c_str());
        else if (key == "prn_list") h->prn.assign(v.begin(), v.end());
        else if (key == "rc") h->rc = v;
        else if (key == "ri") h->ri = v;
        else if (key == "fc") h->fc = v;
        else if (key == "fi") h->fi = v;
        else if (key == "cp") h->cp.assign(v.begin(), v.end());
        else if (key == "cp_timestamp") h->cp_timestamp.assign(v.begin(), v.end());
        else if (key == "TOW") h->TOW.assign(v.begin(), v.end());
        else if (key == "t_oe") h->t_oe = (int)v[0];
        // every other key (rxTime_a, ephemeris rows, ...) is ignored like the reference does
    }
    const size_t n = h->prn.size();
    if (!n || h->rc.size() < n || h->ri.size() < n || h->fc.size() < n || h->fi.size() < n || h->cp.size() < n ||
        h->cp_timestamp.size() < n || h->TOW.size() < n || h->X_ECEF.size() < 4)
        return -1;
    h->X_ECEF.resize(8, 0.0);
    return 0;
}

int ReadGridCsv(const std::string& path, std::vector<double>* g) {
    std::ifstream f(path.c_str());
    if (!f) return -1;
    g->clear();
    std::string line;
    while (std::getline(f, line)) {
        std::vector<std::string> t = split_csv(line);
        if (t.size() < 4) continue;
        for (int i = 0; i < 4; ++i) g->push_back(std::atof(t[i].c_str()));
    }
    return g->empty() ? -1 : 0;
}

const Eph* SelectEph(const std::vector<EphSet>& nav, int prn, double t) {
    const Eph* best = nullptr;
    double bestd = 0;
    if (prn < 1 || prn > kPrnMax) return nullptr;
    for (size_t i = 0; i < nav.size(); ++i) {
        if (!nav[i].valid[prn]) continue;
        const double d = std::fabs(nav[i].toes - t);
        if (!best || d < bestd) { best = &nav[i].eph[prn]; bestd = d; }
    }
    return best;
}

static double wk(double t) { return t > 302400.0 ? t - 604800.0 : (t < -302400.0 ? t + 604800.0 : t); }

static bool kepler(double M, double ecc, double* Eout) {     // Newton, <= 10 iterations, 1e-12 (cuchanmgr.h:13-14)
    double E = M, dE = 1.0;
    int it = 0;
    while (it < 10 && std::fabs(dE) > 1e-12) {
        dE = -(M - E + ecc * std::sin(E)) / (-1.0 + ecc * std::cos(E));
        E = std::fmod(E + dE, kTwoPi);
        ++it;
    }
    *Eout = E;
    return std::fabs(dE) <= 1e-12;
}

bool SatPosition(const Eph& q, double tx, SatState* out) {
    const double n = std::sqrt(kMu / (q.A * q.A * q.A)) + q.deln;
    double tc = wk(tx - q.tocs);
    double clkb = q.f2 * tc * tc + q.f1 * tc + q.f0 - q.tgd;
    double tk = wk(tx - clkb - q.toes);
    double E;
    if (!kepler(std::fmod(q.M0 + n * tk, kTwoPi), q.e, &E)) return false;
    const double dtr = kFRel * q.e * q.sqrtA * std::sin(E);
    tc = tx - (clkb + dtr) - q.tocs;
    clkb = q.f2 * tc * tc + q.f1 * tc + q.f0 + dtr - q.tgd;
    const double clkd = q.f1 + 2.0 * q.f2 * tc;
    tk = wk(tx - clkb - q.toes);
    if (!kepler(std::fmod(q.M0 + n * tk, kTwoPi), q.e, &E)) return false;
    const double sE = std::sin(E), cE = std::cos(E), esq = q.e * q.e;
    const double v = std::atan2(std::sqrt(1.0 - esq) * sE / (1.0 - q.e * cE), (cE - q.e) / (1.0 - q.e * cE));
    double u = std::fmod(v + q.omg, kTwoPi);
    double c2 = std::cos(2.0 * u), s2 = std::sin(2.0 * u);
    u += q.cuc * c2 + q.cus * s2;
    const double r = q.A * (1.0 - q.e * cE) + q.crc * c2 + q.crs * s2;
    const double inc = q.i0 + q.idot * tk + q.cic * c2 + q.cis * s2;
    const double om = std::fmod(q.OMG0 + (q.OMGd - kOmegaE) * tk - kOmegaE * q.toes, kTwoPi);
    const double xo = r * std::cos(u), yo = r * std::sin(u);
    const double co = std::cos(om), so = std::sin(om), ci = std::cos(inc), si = std::sin(inc);
    out->x = xo * co - yo * so * ci;
    out->y = xo * so + yo * co * ci;
    out->z = yo * si;
    out->clkb = clkb;
    // velocity (Remondi form, cuchanmgr.cu:177-204); harmonics re-evaluated at the corrected u
    c2 = std::cos(2.0 * u); s2 = std::sin(2.0 * u);
    const double edot = n / (1.0 - q.e * cE);
    const double vdot = sE * edot * (1.0 + q.e * std::cos(v)) / (std::sin(v) * (1.0 - q.e * cE));
    const double udot = vdot + 2.0 * (q.cus * c2 - q.cuc * s2) * vdot;
    const double rdot = q.A * q.e * sE * edot + 2.0 * (q.crs * c2 - q.crc * s2) * vdot;
    const double idd = q.idot + (q.cis * c2 - q.cic * s2) * 2 * vdot;
    const double vxo = rdot * std::cos(u) - yo * udot, vyo = rdot * std::sin(u) + xo * udot;
    const double od = q.OMGd - kOmegaE;
    const double ta = vxo - yo * ci * od, tb = xo * od + vyo * ci - yo * si * idd;
    out->vx = ta * co - tb * so;
    out->vy = ta * so + tb * co;
    out->vz = vyo * si + yo * ci * idd;
    out->clkd = clkd;
    return true;
}

SatState RotateSat(const SatState& s, double tau) {
    const double c = std::cos(-kOmegaE * tau), sn = std::sin(-kOmegaE * tau);
    SatState o = s;
    o.x = c * s.x - sn * s.y;
    o.y = sn * s.x + c * s.y;
    o.vx = c * s.vx - sn * s.vy - kOmegaE * sn * s.x - kOmegaE * c * s.y;
    o.vy = sn * s.vx + c * s.vy + kOmegaE * c * s.x - kOmegaE * sn * s.y;
    return o;
}

void EcefToLatLon(const double* p, double* lat, double* lon) {
    const double pn = std::sqrt(p[0] * p[0] + p[1] * p[1]);
    const double th = std::atan2(p[2] * kWgsA, pn * kWgsB);
    const double st = std::sin(th), ct = std::cos(th);
    *lat = std::atan2(p[2] + kWgsEp * kWgsEp * kWgsB * st * st * st, pn - kWgsE * kWgsE * kWgsA * ct * ct * ct);
    *lon = std::atan2(p[1], p[0]);
}

void EnuToEcefMatrix(double lat, double lon, double* R) {
    const double sl = std::sin(lat), cl = std::cos(lat), so = std::sin(lon), co = std::cos(lon);
    R[0] = -so; R[1] = -sl * co; R[2] = cl * co;
    R[3] = co;  R[4] = -sl * so; R[5] = cl * so;
    R[6] = 0.0; R[7] = cl;       R[8] = sl;
}

double TxTime(int cpRefTow, int cpElapsed, int cpRef, double codePhase) {
    return cpRefTow + ((cpElapsed - cpRef) * kTCA) + (codePhase / kFCA);
}

void MakeGrid(const int dims[4], const double spacing[4], int gridType, std::vector<double>* g,
              std::vector<double>* timeGrid) {
    std::vector<double> ax[4];
    for (int k = 0; k < 4; ++k) {
        const int n = dims[k], h = (n - 1) / 2;
        ax[k].resize(n);
        for (int i = 0; i < n; ++i) {
            double v = spacing[k] * (i - h);
            if (gridType == 2) {                      // ArthurBasis: outer quarters stretched x3
                const bool outer = (i < h / 2) || ((n - i) < h / 2);
                if (outer) v = 3 * spacing[k] * (i - h) + (i < h ? 1.0 : -1.0) * spacing[k] * ((h / 2) + 1) * 2;
            }
            ax[k][i] = v;
        }
    }
    const size_t G = (size_t)dims[0] * dims[1] * dims[2] * dims[3];
    g->resize(G * 4);
    size_t i = 0;
    for (int a = 0; a < dims[0]; ++a)
        for (int b = 0; b < dims[1]; ++b)
            for (int c = 0; c < dims[2]; ++c)
                for (int d = 0; d < dims[3]; ++d, ++i) {
                    (*g)[4 * i] = ax[0][a]; (*g)[4 * i + 1] = ax[1][b]; (*g)[4 * i + 2] = ax[2][c]; (*g)[4 * i + 3] = ax[3][d];
                }
    if (timeGrid) *timeGrid = ax[3];
}

}  // namespace gnss
