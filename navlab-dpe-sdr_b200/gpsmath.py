"""Scenario-side GPS math for the synthetic L1 C/A generator (host, NumPy).

This is the *simulator's* physics (where the satellites are, what the antenna
would receive), not receiver code: the receiver-side channel manager lives in
``host/chanmgr.cpp`` behind the C ABI.  Standard IS-GPS-200 broadcast-orbit
evaluation; constants follow the reference's ``utils/inc/consthelper.h`` so a
scenario is self-consistent with the receiver equations
(``modules/src/cuchanmgr.cu:85-210``).
"""
from __future__ import annotations

import math

import numpy as np

C = 299792458.0
F_L1 = 1.57542e9
F_CA = 1.023e6
L_CA = 1023
T_CA = 1.0e-3
MU = 3.986005e14
OMEGA_E = 7.2921151467e-5
F_REL = -4.442807633e-10
TWO_PI = 6.2831853071796
WGS84_A = 6378137.0
WGS84_B = 6356752.314245
WGS84_E = 0.08181919084262149
WGS84_EP = 0.08209443794969568

_FIELDS = ("af0", "af1", "af2", "iode", "crs", "deln", "M0", "cuc", "e", "cus", "sqrtA",
           "toe", "cic", "OMG0", "cis", "i0", "crc", "omg", "OMGd", "idot", "codes", "week",
           "l2p", "sva", "svh", "tgd", "iodc", "ttr", "fit")


def _num(s):
    s = s.replace("D", "E").replace("d", "E").strip()
    return float(s) if s else 0.0


def _gps_tow(y, mo, d, h, mi, s):
    """Calendar -> GPS seconds of week (valid 1980-2099)."""
    doy = [1, 32, 60, 91, 121, 152, 182, 213, 244, 274, 305, 335]
    days = (y - 1970) * 365 + (y - 1969) // 4 + doy[mo - 1] + d - 2 + (1 if (y % 4 == 0 and mo >= 3) else 0)
    t = days * 86400 + h * 3600 + mi * 60 + s
    t0 = 3657 * 86400                   # 1980-01-06
    return (t - t0) % (86400 * 7)


def read_rinex_nav(path):
    """RINEX 2.x GPS nav file -> list of dicts (one per record, file order)."""
    with open(path, "r") as f:
        lines = f.read().splitlines()
    k = next(i for i, l in enumerate(lines) if "END OF HEADER" in l) + 1
    recs = []
    body = lines[k:]
    for i in range(0, len(body) - 7, 8):
        r = [l.ljust(80) for l in body[i:i + 8]]
        prn = int(r[0][0:2])
        ep = r[0][3:22].split()
        y = int(ep[0]); y += 2000 if y < 80 else 1900
        toc = _gps_tow(y, int(ep[1]), int(ep[2]), int(ep[3]), int(ep[4]), float(ep[5]))
        vals = [_num(r[0][22 + 19 * j:41 + 19 * j]) for j in range(3)]
        for l in r[1:]:
            vals += [_num(l[3 + 19 * j:22 + 19 * j]) for j in range(4)]
        e = dict(zip(_FIELDS, vals))
        e["prn"] = prn
        e["toc"] = float(toc)
        recs.append(e)
    return recs


def pick_eph(recs, prn, t):
    best = None
    for e in recs:
        if e["prn"] == prn and (best is None or abs(e["toe"] - t) < abs(best["toe"] - t)):
            best = e
    if best is None:
        raise KeyError("no ephemeris for PRN %d" % prn)
    return best


def _wk(t):
    return t - 604800.0 if t > 302400.0 else t + 604800.0 if t < -302400.0 else t


def sat_state(e, t_sv):
    """Satellite ECEF position/velocity and clock at SV-clock time ``t_sv``.

    Returns (pos[3], vel[3], clk_bias_s, clk_drift).  ``clk_bias`` includes the
    relativistic term and -TGD, the convention the receiver back-calculation
    expects (cuchanmgr.cu:110-113,173).
    """
    A = e["sqrtA"] ** 2
    n = math.sqrt(MU / A ** 3) + e["deln"]

    def kepler(tk):
        M = math.fmod(e["M0"] + n * tk, TWO_PI)
        E = M
        for _ in range(15):
            dE = (M - E + e["e"] * math.sin(E)) / (1.0 - e["e"] * math.cos(E))
            E += dE
            if abs(dE) < 1e-14:
                break
        return E

    tc = _wk(t_sv - e["toc"])
    clkb = e["af2"] * tc * tc + e["af1"] * tc + e["af0"] - e["tgd"]
    E = kepler(_wk(t_sv - clkb - e["toe"]))
    dtr = F_REL * e["e"] * e["sqrtA"] * math.sin(E)
    tc = t_sv - (clkb + dtr) - e["toc"]
    clkb = e["af2"] * tc * tc + e["af1"] * tc + e["af0"] + dtr - e["tgd"]
    clkd = e["af1"] + 2.0 * e["af2"] * tc
    tk = _wk(t_sv - clkb - e["toe"])
    E = kepler(tk)
    sinE, cosE = math.sin(E), math.cos(E)
    v = math.atan2(math.sqrt(1.0 - e["e"] ** 2) * sinE, cosE - e["e"])
    u0 = v + e["omg"]
    c2, s2 = math.cos(2 * u0), math.sin(2 * u0)
    u = u0 + e["cuc"] * c2 + e["cus"] * s2
    r = A * (1.0 - e["e"] * cosE) + e["crc"] * c2 + e["crs"] * s2
    inc = e["i0"] + e["idot"] * tk + e["cic"] * c2 + e["cis"] * s2
    om = e["OMG0"] + (e["OMGd"] - OMEGA_E) * tk - OMEGA_E * e["toe"]
    xo, yo = r * math.cos(u), r * math.sin(u)
    co, so, ci, si = math.cos(om), math.sin(om), math.cos(inc), math.sin(inc)
    pos = np.array([xo * co - yo * so * ci, xo * so + yo * co * ci, yo * si])
    # velocity (Remondi form)
    c2, s2 = math.cos(2 * u), math.sin(2 * u)
    edot = n / (1.0 - e["e"] * cosE)
    vdot = sinE * edot * (1.0 + e["e"] * math.cos(v)) / (math.sin(v) * (1.0 - e["e"] * cosE))
    udot = vdot + 2.0 * (e["cus"] * c2 - e["cuc"] * s2) * vdot
    rdot = A * e["e"] * sinE * edot + 2.0 * (e["crs"] * c2 - e["crc"] * s2) * vdot
    idd = e["idot"] + (e["cis"] * c2 - e["cic"] * s2) * 2 * vdot
    vxo = rdot * math.cos(u) - yo * udot
    vyo = rdot * math.sin(u) + xo * udot
    od = e["OMGd"] - OMEGA_E
    ta = vxo - yo * ci * od
    tb = xo * od + vyo * ci - yo * si * idd
    vel = np.array([ta * co - tb * so, ta * so + tb * co, vyo * si + yo * ci * idd])
    return pos, vel, clkb, clkd


def rotate_z(pos, vel, tau):
    """Earth-rotation (Sagnac) correction over flight time ``tau``: rotate the
    satellite about z by -OMEGA_E*tau (cuchanmgr.cu:383-404)."""
    c, s = math.cos(-OMEGA_E * tau), math.sin(-OMEGA_E * tau)
    p = np.array([c * pos[0] - s * pos[1], s * pos[0] + c * pos[1], pos[2]])
    v = np.array([c * vel[0] - s * vel[1] - OMEGA_E * s * pos[0] - OMEGA_E * c * pos[1],
                  s * vel[0] + c * vel[1] + OMEGA_E * c * pos[0] - OMEGA_E * s * pos[1],
                  vel[2]])
    return p, v


def ecef_to_latlon(p):
    pn = math.hypot(p[0], p[1])
    th = math.atan2(p[2] * WGS84_A, pn * WGS84_B)
    lat = math.atan2(p[2] + WGS84_EP ** 2 * WGS84_B * math.sin(th) ** 3,
                     pn - WGS84_E ** 2 * WGS84_A * math.cos(th) ** 3)
    return lat, math.atan2(p[1], p[0])


def enu_to_ecef_matrix(lat, lon):
    """Row-major 3x3, columns = East, North, Up unit vectors in ECEF."""
    sl, so, cl, co = math.sin(lat), math.sin(lon), math.cos(lat), math.cos(lon)
    return np.array([-so, -sl * co, cl * co, co, -sl * so, cl * so, 0.0, cl, sl])


def ca_code(prn):
    """C/A Gold code (+1/-1, chip +1 <-> binary 1) from the IS-GPS-200 G2 delay
    table; independent construction from the receiver's LFSR-tap kernel."""
    delays = [5, 6, 7, 8, 17, 18, 139, 140, 141, 251, 252, 254, 255, 256, 257, 258, 469, 470,
              471, 472, 473, 474, 509, 512, 513, 514, 515, 516, 859, 860, 861, 862, 863, 950,
              947, 948, 950]
    g1 = np.ones(10, dtype=np.int8)
    g2 = np.ones(10, dtype=np.int8)
    o1 = np.empty(1023, dtype=np.int8)
    o2 = np.empty(1023, dtype=np.int8)
    for i in range(1023):
        o1[i] = g1[9]
        o2[i] = g2[9]
        f1 = g1[2] ^ g1[9]
        f2 = g2[1] ^ g2[2] ^ g2[5] ^ g2[7] ^ g2[8] ^ g2[9]
        g1 = np.concatenate(([f1], g1[:9]))
        g2 = np.concatenate(([f2], g2[:9]))
    o2 = np.roll(o2, delays[prn - 1])
    return np.where((o1 ^ o2) == 1, 1, -1).astype(np.int8)
