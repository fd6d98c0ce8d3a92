// gnss.h -- host-side GPS support for the flow: broadcast-ephemeris satellite state, RINEX 2.x
// navigation reader, handoff / grid CSV readers, geodetic helpers.  The north star keeps
// "satellite states from RINEX ephemeris, the ENU-clock grid generation and rngrid csv read-in
// on the host in C++" -- this is that code (reference: utils/src/rinexparse.cpp,
// modules/src/dpinit.cpp:247-400, modules/src/cuchanmgr.cu:26-210 which runs it on the GPU).
#ifndef DPE_HOST_GNSS_H_
#define DPE_HOST_GNSS_H_

#include <cstdint>
#include <string>
#include <vector>

namespace gnss {

// utils/inc/consthelper.h:5-27
constexpr double kC = 299792458.0;
constexpr double kTwoPi = 6.2831853071796;
constexpr double kFL1 = 1.57542e9;
constexpr double kFCA = 1.023e6;
constexpr double kLCA = 1023.0;
constexpr double kTCA = 0.001;
constexpr double kMu = 3.986005e14;
constexpr double kFRel = -4.442807633e-10;
constexpr double kOmegaE = 7.2921151467e-5;
constexpr double kWgsA = 6378137.0, kWgsB = 6356752.314245;
constexpr double kWgsE = 0.08181919084262149, kWgsEp = 0.08209443794969568;
constexpr int kPrnMax = 37;

struct Eph {                       // fields of eph_t the path uses (utils/inc/ephhelper.h:98-124)
    int sat = 0, week = 0;
    double A = 0, sqrtA = 0, e = 0, i0 = 0, OMG0 = 0, omg = 0, M0 = 0, deln = 0, OMGd = 0, idot = 0;
    double crc = 0, crs = 0, cuc = 0, cus = 0, cic = 0, cis = 0, toes = 0, tocs = 0, f0 = 0, f1 = 0, f2 = 0, tgd = 0;
};

struct EphSet {                    // ephemerides sharing one TOE (ephhelper.h:148-169)
    double toes = -1;
    bool valid[kPrnMax + 1] = {false};
    Eph eph[kPrnMax + 1];
};

struct SatState { double x, y, z, clkb, vx, vy, vz, clkd; };   // state_t<double>, statehelper.h:11-21

struct Handoff {                   // demofiles/handoff_params_usrp6.csv grammar (dpinit.cpp:247-400)
    double rxTime = 0;
    std::vector<double> X_ECEF;
    long long bytes_read = 0;
    std::vector<int> prn, cp, cp_timestamp, TOW;
    std::vector<double> rc, ri, fc, fi;
    int t_oe = 0;
};

int ReadRinexNav(const std::string& path, std::vector<EphSet>* out);       // 0 = ok
int ReadHandoff(const std::string& path, Handoff* out);
int ReadGridCsv(const std::string& path, std::vector<double>* enu_dt);     // rows of x,y,z,delta_t

const Eph* SelectEph(const std::vector<EphSet>& nav, int prn, double t);   // nearest TOE (cuchanmgr.cu:276-292)
bool SatPosition(const Eph& eph, double txTime, SatState* out);            // CHM_Get_Sat_Pos, cuchanmgr.cu:85-210
SatState RotateSat(const SatState& s, double tau);                         // z-rotation by -OmegaE*tau (+ w x r)
void EcefToLatLon(const double* p, double* lat, double* lon);              // cuchanmgr.cu:37-50
void EnuToEcefMatrix(double lat, double lon, double* R9);                  // cuchanmgr.cu:54-73, row-major
double TxTime(int cpRefTow, int cpElapsed, int cpRef, double codePhase);   // cuchanmgr.cu:258-260

/** Uniform / ArthurBasis 4-D grid (BCM_InitPosGrid, batchcorrmanifold.cu:148-255): t fastest. */
void MakeGrid(const int dims[4], const double spacing[4], int gridType, std::vector<double>* enu_dt,
              std::vector<double>* timeGrid);

}  // namespace gnss
#endif
