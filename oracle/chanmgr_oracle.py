"""NumPy/float64 restatement of the reference's channel manager, broadcast
ephemeris satellite state, RINEX 2.x nav parser and handoff CSV parser
(TEST INFRASTRUCTURE; see oracle/__init__.py).

Follows CUDARecv ``modules/src/cuchanmgr.cu``, ``utils/src/rinexparse.cpp``,
``utils/src/converters.cpp`` and ``modules/src/dpinit.cpp`` (paths relative to
/root/reference/cudarecv).  Scalar Python on purpose: <= 12 channels.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from .dpe_oracle import (CONST_C, CONST_2PI, CONST_F_CA, CONST_F_L1, CONST_L_CA,
                         CONST_T_CA, CONST_OEDot, CONST_F, CONST_WGS84_A,
                         CONST_WGS84_B, CONST_WGS84_E, CONST_WGS84_EP)

MU_GPS = 3.9860050e14            # utils/inc/ephhelper.h:52
CHM_RTOL_KEPLER = 1e-12          # modules/inc/cuchanmgr.h:13-14
CHM_MAX_ITER_KEPLER = 10


# --------------------------------------------------------------------------
# RINEX 2.x navigation parser (utils/src/rinexparse.cpp)
# --------------------------------------------------------------------------

def _str2num(s: str, i: int, n: int) -> float:
    """auxil/src/auxil.cpp:64-72."""
    if i < 0 or len(s) < i:
        return 0.0
    sub = s[i:i + n].replace("d", "E").replace("D", "E").strip()
    try:
        # sscanf("%lf") parses the longest valid prefix
        tok = sub.split()[0] if sub else ""
        return float(tok)
    except (ValueError, IndexError):
        return 0.0


def _epoch2time(ep):
    """utils/src/converters.cpp epoch2time (rtklib): returns (time_t, frac)."""
    doy = [1, 32, 60, 91, 121, 152, 182, 213, 244, 274, 305, 335]
    year, mon, day = int(ep[0]), int(ep[1]), int(ep[2])
    if year < 1970 or 2099 < year or mon < 1 or 12 < mon:
        return 0, 0.0
    days = (year - 1970) * 365 + (year - 1969) // 4 + doy[mon - 1] + day - 2 + \
        (1 if (year % 4 == 0 and mon >= 3) else 0)
    sec = int(math.floor(ep[5]))
    return days * 86400 + int(ep[3]) * 3600 + int(ep[4]) * 60 + sec, ep[5] - sec


_GPST0 = _epoch2time([1980, 1, 6, 0, 0, 0])


def _time2gpst(t):
    """converters.cpp time2gpst: (week, tow)."""
    sec = t[0] - _GPST0[0]
    w = int(sec // (86400 * 7))
    return w, float(sec - w * 86400 * 7) + t[1]


@dataclass
class Eph:
    """utils/inc/ephhelper.h:98-124 (eph_t), fields the path uses."""
    sat: int = 0
    iode: int = 0
    iodc: int = 0
    week: int = 0
    A: float = 0.0
    sqrt_A: float = 0.0
    e: float = 0.0
    e_sqr: float = 0.0
    i0: float = 0.0
    OMG0: float = 0.0
    omg: float = 0.0
    M0: float = 0.0
    deln: float = 0.0
    OMGd: float = 0.0
    idot: float = 0.0
    crc: float = 0.0
    crs: float = 0.0
    cuc: float = 0.0
    cus: float = 0.0
    cic: float = 0.0
    cis: float = 0.0
    toes: float = 0.0
    tocs: float = 0.0
    f0: float = 0.0
    f1: float = 0.0
    f2: float = 0.0
    tgd: float = 0.0


@dataclass
class EphSet:
    """utils/inc/ephhelper.h:148-169 (ephSet_t): ephemerides sharing one TOE."""
    toes: float = -1.0
    eph: dict = field(default_factory=dict)    # prn -> Eph (ephValid == key present)


def read_rinex_nav(path: str) -> list:
    """utils/src/rinexparse.cpp:19-57,185-366.  Returns [EphSet] in file order
    of first appearance of each TOE (std::find_if + push_back, :200-216)."""
    with open(path, "r") as f:
        lines = f.read().splitlines()
    # header (:64-129): skip until END OF HEADER
    k = 0
    while k < len(lines):
        ln = lines[k]
        k += 1
        if len(ln) > 60 and "END OF HEADER" in ln[60:]:
            break
    nav: list = []
    data: list = []
    prn = 0
    toc = None
    for ln in lines[k:]:
        ln = ln.ljust(80)
        if not data:                                            # :243-275
            prn = int(_str2num(ln, 0, 2))
            ep = ln[3:22].split()
            if len(ep) < 6:
                continue
            ep = [float(v) for v in ep]
            ep[0] += 2000.0 if ep[0] < 80.0 else 1900.0 if ep[0] < 100.0 else 0.0
            toc = _epoch2time(ep)
            data = [_str2num(ln, 22 + 19 * j, 19) for j in range(3)]
        else:                                                   # :276-300
            data += [_str2num(ln, 3 + 19 * j, 19) for j in range(4)]
            if len(data) >= 31:
                e = _decode_eph(prn, toc, data)
                data = []
                if e is None:
                    continue
                tgt = None
                for s in nav:                                   # :203-216
                    if s.toes == int(e.toes):
                        tgt = s
                        break
                if tgt is None:
                    tgt = EphSet(toes=e.toes)
                    nav.append(tgt)
                tgt.eph[e.sat] = e                              # :219-220
    return nav


def _decode_eph(prn, toc, d):
    """rinexparse.cpp:308-366 (DecodeRinexBody)."""
    if not (1 <= prn <= 32):
        return None
    e = Eph()
    e.sat = prn
    e.week, e.tocs = _time2gpst(toc)
    e.f0, e.f1, e.f2 = d[0], d[1], d[2]
    e.sqrt_A = d[10]; e.e = d[8]; e.i0 = d[15]; e.OMG0 = d[13]
    e.omg = d[17]; e.M0 = d[6]; e.deln = d[5]; e.OMGd = d[18]
    e.idot = d[19]; e.crc = d[16]; e.crs = d[4]; e.cuc = d[7]
    e.cus = d[9]; e.cic = d[12]; e.cis = d[14]
    e.A = e.sqrt_A * e.sqrt_A
    e.e_sqr = e.e * e.e
    e.iode = int(d[3]); e.iodc = int(d[26])
    e.toes = d[11]
    e.week = int(d[21])
    e.tgd = d[25]
    return e


# --------------------------------------------------------------------------
# Handoff CSV (modules/src/dpinit.cpp:247-400)
# --------------------------------------------------------------------------

def read_handoff(path: str) -> dict:
    out = {}
    with open(path, "r") as f:
        for line in f:
            p = line.strip().split(",")
            key, vals = p[0], [v for v in p[1:] if v != ""]
            if key in ("rxTime", "rxTime_a"):
                out[key] = float(vals[0])
            elif key == "X_ECEF":
                out[key] = np.array([float(v) for v in vals])
            elif key == "bytes_read":
                out[key] = int(vals[0])
            elif key == "prn_list":
                out[key] = np.array([int(v) for v in vals], dtype=np.uint8)
            elif key in ("rc", "ri", "fc", "fi"):
                out[key] = np.array([float(v) for v in vals])
            elif key in ("cp", "cp_timestamp", "TOW"):
                out[key] = np.array([int(float(v)) for v in vals], dtype=np.int32)
            elif key == "t_oe":
                out[key] = int(float(vals[0]))
    return out


# --------------------------------------------------------------------------
# cuChanMgr (modules/src/cuchanmgr.cu)
# --------------------------------------------------------------------------

def _week_crossover(t):
    """cuchanmgr.cu:26-31."""
    if t > 302400.0:
        return t - 604800.0
    if t < -302400.0:
        return t + 604800.0
    return t


def _kepler(M, e):
    """cuchanmgr.cu:97-107: Newton, <=10 iterations, tol 1e-12, fmod each step."""
    E = M
    dE = 1.0
    it = 0
    while it < CHM_MAX_ITER_KEPLER and abs(dE) > CHM_RTOL_KEPLER:
        f = M - E + e * math.sin(E)
        dfdE = -1.0 + e * math.cos(E)
        dE = -f / dfdE
        E = math.fmod(E + dE, CONST_2PI)
        it += 1
    return E, abs(dE) <= CHM_RTOL_KEPLER


def get_sat_pos(eph: Eph, tx_time: float):
    """cuchanmgr.cu:85-210 (CHM_Get_Sat_Pos).  Returns state[8] =
    (x,y,z,clkb, vx,vy,vz,clkd) in ECEF at tx_time, clock terms in seconds."""
    n = math.sqrt(MU_GPS / (eph.A * eph.A * eph.A)) + eph.deln
    tc = _week_crossover(tx_time - eph.tocs)
    clkb = eph.f2 * tc * tc + eph.f1 * tc + eph.f0 - eph.tgd
    tk = _week_crossover(tx_time - clkb - eph.toes)
    M = math.fmod(eph.M0 + n * tk, CONST_2PI)
    E, ok = _kepler(M, eph.e)
    if not ok:
        return None
    dtr = CONST_F * eph.e * eph.sqrt_A * math.sin(E)
    tc = tx_time - (clkb + dtr) - eph.tocs
    clkb = eph.f2 * tc * tc + eph.f1 * tc + eph.f0 + dtr - eph.tgd
    clkd = eph.f1 + 2.0 * eph.f2 * tc
    tk = _week_crossover(tx_time - clkb - eph.toes)
    M = math.fmod(eph.M0 + n * tk, CONST_2PI)
    E, ok = _kepler(M, eph.e)
    if not ok:
        return None
    sinE, cosE = math.sin(E), math.cos(E)
    v = math.atan2(math.sqrt(1.0 - eph.e_sqr) * sinE / (1.0 - eph.e * cosE),
                   (cosE - eph.e) / (1.0 - eph.e * cosE))
    u = math.fmod(v + eph.omg, CONST_2PI)
    cos2u, sin2u = math.cos(2.0 * u), math.sin(2.0 * u)
    u += eph.cuc * cos2u + eph.cus * sin2u
    r = eph.A * (1.0 - eph.e * cosE) + eph.crc * cos2u + eph.crs * sin2u
    i = eph.i0 + eph.idot * tk + eph.cic * cos2u + eph.cis * sin2u
    omegak = math.fmod(eph.OMG0 + (eph.OMGd - CONST_OEDot) * tk - CONST_OEDot * eph.toes, CONST_2PI)
    x_op, y_op = r * math.cos(u), r * math.sin(u)
    cos_o, sin_o = math.cos(omegak), math.sin(omegak)
    cosi, sini = math.cos(i), math.sin(i)
    sx = x_op * cos_o - y_op * sin_o * cosi
    sy = x_op * sin_o + y_op * cos_o * cosi
    sz = y_op * sini
    # velocity (:177-204) -- note cos2u/sin2u recomputed from the CORRECTED u
    cos2u, sin2u = math.cos(2.0 * u), math.sin(2.0 * u)
    edot = n / (1.0 - eph.e * cosE)
    vdot = sinE * edot * (1.0 + eph.e * math.cos(v)) / (math.sin(v) * (1.0 - eph.e * cosE))
    udot = vdot + 2.0 * (eph.cus * cos2u - eph.cuc * sin2u) * vdot
    rdot = eph.A * eph.e * sinE * edot + 2.0 * (eph.crs * cos2u - eph.crc * sin2u) * vdot
    idotdot = eph.idot + (eph.cis * cos2u - eph.cic * sin2u) * 2 * vdot
    vx_op = rdot * math.cos(u) - y_op * udot
    vy_op = rdot * math.sin(u) + x_op * udot
    omegadot = eph.OMGd - CONST_OEDot
    tmpa = vx_op - y_op * cosi * omegadot
    tmpb = x_op * omegadot + vy_op * cosi - y_op * sini * idotdot
    vx = tmpa * cos_o - tmpb * sin_o
    vy = tmpa * sin_o + tmpb * cos_o
    vz = vy_op * sini + y_op * cosi * idotdot
    return np.array([sx, sy, sz, clkb, vx, vy, vz, clkd])


def select_eph(nav: list, prn: int, t: float):
    """cuchanmgr.cu:276-292: valid set with minimal |toe - t| (first wins ties)."""
    best = None
    for s in nav:
        if prn in s.eph:
            if best is None or abs(s.toes - t) < abs(best.toes - t):
                best = s
    return None if best is None else best.eph[prn]


def tx_time_of(cp_ref_tow, cp_elapsed, cp_ref, code_phase):
    """cuchanmgr.cu:258-260."""
    return cp_ref_tow + ((cp_elapsed - cp_ref) * CONST_T_CA) + (code_phase / CONST_F_CA)


def rotate_sat(sat, tau):
    """cuchanmgr.cu:383-404 / :895-916: rotate about z by -OEDot*tau (+ w x r)."""
    c = math.cos(-CONST_OEDot * tau)
    s = math.sin(-CONST_OEDot * tau)
    out = np.empty(8)
    out[0] = c * sat[0] - s * sat[1]
    out[1] = s * sat[0] + c * sat[1]
    out[2] = sat[2]
    out[3] = sat[3]
    out[4] = c * sat[4] - s * sat[5] - CONST_OEDot * s * sat[0] - CONST_OEDot * c * sat[1]
    out[5] = s * sat[4] + c * sat[5] + CONST_OEDot * c * sat[0] - CONST_OEDot * s * sat[1]
    out[6] = sat[6]
    out[7] = sat[7]
    return out


def ecef2ll_rad(p):
    """cuchanmgr.cu:37-50 (closed-form latitude, atan2 longitude)."""
    pn = math.sqrt(p[0] * p[0] + p[1] * p[1])
    theta = math.atan2(p[2] * CONST_WGS84_A, pn * CONST_WGS84_B)
    lat = math.atan2(p[2] + CONST_WGS84_EP ** 2 * CONST_WGS84_B * math.sin(theta) ** 3,
                     pn - CONST_WGS84_E ** 2 * CONST_WGS84_A * math.cos(theta) ** 3)
    lon = math.atan2(p[1], p[0])
    return lat, lon


def enu2ecef_mat(lat, lon):
    """cuchanmgr.cu:54-73 (row-major 3x3)."""
    sl, so, cl, co = math.sin(lat), math.sin(lon), math.cos(lat), math.cos(lon)
    return np.array([-so, -sl * co, cl * co,
                     co, -sl * so, cl * so,
                     0.0, cl, sl])


@dataclass
class ChanState:
    """Per-channel arrays owned by cuChanMgr (cuchanmgr.cu:1046-1078)."""
    prn: np.ndarray
    rc_start: np.ndarray
    rc_end: np.ndarray
    ri_start: np.ndarray
    ri_end: np.ndarray
    fc: np.ndarray
    fi: np.ndarray
    cp_start: np.ndarray
    cp_end: np.ndarray
    cp_ref: np.ndarray
    cp_ref_tow: np.ndarray
    tx_time: np.ndarray
    sat: np.ndarray          # [C][8] satStates_d (un-rotated, at tx_time)
    rx_time: float
    T: float
    doppler_sign: int = 1


def _time_update(ch: ChanState, nav, center, i, rx_time):
    """Shared tail of CHM_PropagateChannels (:451-602) and
    CHM_TimeUpdateChannels (:675-823): enhanced time update of channel i."""
    T = ch.T
    temp1 = math.floor((ch.fc[i] * T + ch.rc_end[i]) / CONST_L_CA)
    cp_pred = ch.cp_end[i] + temp1
    temp2 = math.fmod(ch.fc[i] * T + ch.rc_end[i], float(CONST_L_CA))
    if temp2 < 0.0:
        temp2 += CONST_L_CA
    tx_pred = ch.cp_ref_tow[i] + ((cp_pred - ch.cp_ref[i]) * CONST_T_CA) + (temp2 / CONST_F_CA)
    eph = select_eph(nav, int(ch.prn[i]), tx_pred)
    sat_pred = get_sat_pos(eph, tx_pred)
    tau = rx_time + T - (tx_pred + (center[3] / CONST_C)) + sat_pred[3]
    rot = rotate_sat(sat_pred, tau)
    lx, ly, lz = rot[0] - center[0], rot[1] - center[1], rot[2] - center[2]
    rng = math.sqrt(lx * lx + ly * ly + lz * lz)
    pr = rng - CONST_C * rot[3] + center[3]
    bc_tx = rx_time + T - pr / CONST_C
    frac = bc_tx - ch.cp_ref_tow[i] - ((int(ch.cp_end[i]) - int(ch.cp_ref[i])) * CONST_T_CA)
    bc_rc = frac * CONST_F_CA
    ch.cp_start[i] = ch.cp_end[i]
    t1 = math.floor(bc_rc / CONST_L_CA)
    ch.rc_start[i] = ch.rc_end[i]
    t2 = math.fmod(bc_rc, float(CONST_L_CA))
    if t2 < 0.0:
        t2 += CONST_L_CA
    ch.cp_end[i] += int(t1)
    ch.rc_end[i] = t2
    ch.ri_start[i] = ch.ri_end[i]
    t3 = math.fmod(ch.fi[i] * T + ch.ri_end[i], 1.0)
    if t3 < 0.0:
        t3 += 1.0
    ch.ri_end[i] = t3
    ch.tx_time[i] = tx_time_of(ch.cp_ref_tow[i], int(ch.cp_end[i]), int(ch.cp_ref[i]), ch.rc_end[i])
    ch.sat[i] = get_sat_pos(eph, ch.tx_time[i])


def chanmgr_start(nav, handoff, T, center, doppler_sign=1) -> ChanState:
    """cuChanMgr::Start (cuchanmgr.cu:1003-1185): load handoff (end-referenced
    slots, :1053-1069), CHM_ComputeSatStates (:240-306), CHM_TimeUpdateChannels
    (:641-829), rxTime += T (:1121)."""
    C = len(handoff["prn_list"])
    T = round(T * 1.0e6) / 1.0e6                                     # :1037
    ch = ChanState(
        prn=np.array(handoff["prn_list"], dtype=np.uint8),
        rc_start=np.zeros(C), rc_end=np.array(handoff["rc"], dtype=np.float64),
        ri_start=np.zeros(C), ri_end=np.array(handoff["ri"], dtype=np.float64),
        fc=np.array(handoff["fc"], dtype=np.float64), fi=np.array(handoff["fi"], dtype=np.float64),
        cp_start=np.zeros(C, dtype=np.int32), cp_end=np.array(handoff["cp"], dtype=np.int32),
        cp_ref=np.array(handoff["cp_timestamp"], dtype=np.int32),
        cp_ref_tow=np.array(handoff["TOW"], dtype=np.int32),
        tx_time=np.zeros(C), sat=np.zeros((C, 8)),
        rx_time=float(handoff["rxTime"]), T=T, doppler_sign=doppler_sign)
    for i in range(C):
        ch.tx_time[i] = tx_time_of(ch.cp_ref_tow[i], int(ch.cp_end[i]), int(ch.cp_ref[i]), ch.rc_end[i])
        eph = select_eph(nav, int(ch.prn[i]), ch.tx_time[i])
        ch.sat[i] = get_sat_pos(eph, ch.tx_time[i])
    for i in range(C):
        _time_update(ch, nav, center, i, ch.rx_time)
    ch.rx_time += T
    return ch


def chanmgr_update(ch: ChanState, nav, center):
    """cuChanMgr::Update (cuchanmgr.cu:1224-1268): CHM_PropagateChannels
    (:338-608) with center = x_{k|k}, then rxTime += T."""
    T = ch.T
    for i in range(len(ch.prn)):
        tau = ch.rx_time - (ch.tx_time[i] + (center[3] / CONST_C)) + ch.sat[i][3]    # :380
        rot = rotate_sat(ch.sat[i], tau)
        ve = [center[4] - CONST_OEDot * center[1], center[5] + CONST_OEDot * center[0],
              center[6], center[7]]                                                  # :408-411
        lx, ly, lz = rot[0] - center[0], rot[1] - center[1], rot[2] - center[2]
        rng = math.sqrt(lx * lx + ly * ly + lz * lz)
        rate = ((lx / rng) * (ve[0] - rot[4])) + ((ly / rng) * (ve[1] - rot[5])) + \
               ((lz / rng) * (ve[2] - rot[6]))
        bc_fi = CONST_F_L1 * ((rate - ve[3]) / CONST_C + rot[7]) / ch.doppler_sign   # :425
        pr = rng - CONST_C * rot[3] + center[3]
        bc_tx = ch.rx_time - pr / CONST_C
        frac = bc_tx - ch.cp_ref_tow[i] - ((int(ch.cp_end[i]) - int(ch.cp_ref[i])) * CONST_T_CA)
        bc_rc = frac * CONST_F_CA
        bc_fc = CONST_F_CA + (ch.doppler_sign * CONST_F_CA / CONST_F_L1) * bc_fi + \
            (bc_rc - ch.rc_end[i]) / T                                               # :434
        ch.fi[i] = bc_fi
        ch.fc[i] = bc_fc
        _time_update(ch, nav, center, i, ch.rx_time)
    ch.rx_time += T


def grid_prep(ch: ChanState, center_kk1, time_grid):
    """CHM_GridPrep (cuchanmgr.cu:853-923).  Returns (batchSatStates [C*T][8],
    enu2ecef[9]); uses ch.rx_time already advanced by T (:1121,:1249)."""
    C = len(ch.prn)
    Tn = len(time_grid)
    out = np.empty((C * Tn, 8))
    for c in range(C):
        for it in range(Tn):
            tau = ch.rx_time - (ch.tx_time[c] + ((time_grid[it] + center_kk1[3]) / CONST_C)) + ch.sat[c][3]
            out[c * Tn + it] = rotate_sat(ch.sat[c], tau)
    lat, lon = ecef2ll_rad(center_kk1)
    return out, enu2ecef_mat(lat, lon)
