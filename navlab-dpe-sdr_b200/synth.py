"""Synthetic GPS L1 C/A scenario generator (host, NumPy).

The reference's 450 MB demo capture and ``rngrid3.csv`` are missing from its
checkout, so every workload runs on a synthetic stand-in built here: satellites
from a broadcast-ephemeris (RINEX 2.x) file, a static receiver with a clock
bias and drift, and per-PRN signals

    A * d(t) * code(floor(N_c(t)) mod 1023) * exp(+j 2 pi phi_c(t))  + noise

with code phase ``N_c`` and carrier phase ``phi_c`` derived from the same
pseudorange model the receiver back-calculates
(``modules/src/batchcorrmanifold.cu:1779-1791``), so a handoff taken from the
truth is self-consistent.  Samples are complex int16, interleaved I,Q
(``README.md:56`` of the reference), seeds fixed.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import numpy as np

from . import gpsmath as gm

_HERE = os.path.dirname(os.path.abspath(__file__))
# broadcast ephemerides of the TOE = 417600 s set (GPS week 2008, day 186 of 2018) in RINEX 2.10 nav format:
# the records of the reference's demofiles/nist1860.18n the demo epoch uses
DEFAULT_RINEX = os.path.join(_HERE, "data", "brdc_toe417600.18n")

PRNS_8 = (2, 3, 6, 12, 17, 19, 24, 28)          # demofiles/handoff_params_usrp6.csv:5
PRNS_12 = PRNS_8 + (1, 5, 22, 30)               # the other TOE=417600 satellites


@dataclass
class ScenarioConfig:
    fs: float = 2.5e6
    T: float = 0.02
    prns: tuple = PRNS_8
    cn0_dbhz: float = 45.0
    noise_sigma: float = 1000.0                 # per I/Q component, LSB
    seed: int = 20180704
    rinex: str = DEFAULT_RINEX
    truth_ecef: tuple = (151158.46510991786, -4885422.338576897, 4090087.0543405097)
    truth_clock_bias_m: float = 175068.5560268988
    truth_clock_drift_mps: float = -0.11462027018250964
    truth_vel_enu: tuple = (0.0, 0.0, 0.0)      # receiver velocity, east / north / up m/s (0 = the static demo receiver)
    rx_time0: float = 414006.0680031631
    tow_ref: int = 414006
    cp_ref_base: int = 1000
    doppler_sign: int = 1


def spread_axis():
    """The PyGNSS 'spread grid' axis (pygnss receiver.py:995-1006)."""
    return np.array([-22, -19, -16, -13, -10, -7, -6, -5, -4, -3, -2, -1, 0,
                     1, 2, 3, 4, 5, 6, 7, 10, 13, 16, 19, 22], dtype=np.float64)


def spread_grid():
    """25^4 rngrid3-style position-clock grid: ENU 5 m * axis, clock 6 m * axis,
    x slowest, t fastest."""
    a = spread_axis()
    gx, gy, gz, gt = np.meshgrid(5 * a, 5 * a, 5 * a, 6 * a, indexing="ij")
    return np.ascontiguousarray(np.stack([gx.ravel(), gy.ravel(), gz.ravel(), gt.ravel()], axis=1))


def uniform_grid(n, spacing):
    """Uniform n^4 grid, offsets spacing*(i - (n-1)//2), t fastest
    (batchcorrmanifold.cu:164-187, :2331).  ``spacing``: scalar or 4-tuple."""
    sp = np.broadcast_to(np.asarray(spacing, dtype=np.float64), (4,))
    h = (n - 1) // 2
    ax = [sp[k] * (np.arange(n) - h) for k in range(4)]
    gx, gy, gz, gt = np.meshgrid(*ax, indexing="ij")
    grid = np.ascontiguousarray(np.stack([gx.ravel(), gy.ravel(), gz.ravel(), gt.ravel()], axis=1))
    return grid, ax[3].astype(np.float64)


class Scenario:
    def __init__(self, cfg: ScenarioConfig | None = None):
        self.cfg = cfg or ScenarioConfig()
        c = self.cfg
        self.S = int(round(c.fs * c.T))
        self.C = len(c.prns)
        self.recs = gm.read_rinex_nav(c.rinex)
        self.eph = [gm.pick_eph(self.recs, p, c.rx_time0) for p in c.prns]
        self.codes = np.stack([gm.ca_code(p) for p in c.prns]).astype(np.float64)
        self.cp_ref = np.array([c.cp_ref_base + (p % 12) for p in c.prns], dtype=np.int32)
        self.amp = math.sqrt(2.0) * c.noise_sigma * math.sqrt(10 ** (c.cn0_dbhz / 10.0) / c.fs)
        rng = np.random.default_rng(c.seed)
        self.phi0 = rng.random(self.C)
        self.bits = rng.integers(0, 2, size=(self.C, 4096)) * 2 - 1
        self._noise_seed = c.seed + 1
        lat, lon = gm.ecef_to_latlon(np.asarray(c.truth_ecef, dtype=np.float64))
        self.vel_ecef = np.asarray(gm.enu_to_ecef_matrix(lat, lon), dtype=np.float64).reshape(3, 3) @ \
            np.asarray(c.truth_vel_enu, dtype=np.float64)
        self._pr0 = np.array([self.link(k, c.rx_time0)["pr"] for k in range(self.C)])
        self._edge_cache = {}

    # ---- physics -----------------------------------------------------------
    def rx_state(self, rx_time):
        c = self.cfg
        dt = c.truth_clock_bias_m + c.truth_clock_drift_mps * (rx_time - c.rx_time0)
        p = np.asarray(c.truth_ecef, dtype=np.float64) + self.vel_ecef * (rx_time - c.rx_time0)
        return np.array([p[0], p[1], p[2], dt, self.vel_ecef[0], self.vel_ecef[1], self.vel_ecef[2],
                         c.truth_clock_drift_mps])

    def link(self, k, rx_time, state=None):
        """Solve the light-time equation for PRN index k at receiver clock time
        ``rx_time`` (receiver state ``state``, default truth).  Returns dict with
        tx_rel (= SV-clock transmit time minus tow_ref), pr, sat pos/vel/clk."""
        x = self.rx_state(rx_time) if state is None else state
        e = self.eph[k]
        rel = rx_time - self.cfg.tow_ref
        pr = 0.072 * gm.C
        for _ in range(6):
            tx_sv = self.cfg.tow_ref + (rel - pr / gm.C)
            pos, vel, clkb, clkd = gm.sat_state(e, tx_sv)
            tau = (rx_time - x[3] / gm.C) - (tx_sv - clkb)
            p, v = gm.rotate_z(pos, vel, tau)
            rho = math.sqrt((p[0] - x[0]) ** 2 + (p[1] - x[1]) ** 2 + (p[2] - x[2]) ** 2)
            pr = rho - gm.C * clkb + x[3]
        los = (p - x[:3]) / rho
        ve = np.array([x[4] - gm.OMEGA_E * x[1], x[5] + gm.OMEGA_E * x[0], x[6]])
        rate = float(los @ (ve - v))
        fi = gm.F_L1 * ((rate - x[7]) / gm.C + clkd) / self.cfg.doppler_sign
        return dict(tx_rel=rel - pr / gm.C, pr=pr, rho=rho, pos=pos, vel=vel, clkb=clkb,
                    clkd=clkd, fi=fi, rot_pos=p, rot_vel=v)

    def _edge(self, b):
        """Per-PRN total code chips and carrier cycles at the start of block b."""
        if b not in self._edge_cache:
            t = self.cfg.rx_time0 + b * self.cfg.T
            L = [self.link(k, t) for k in range(self.C)]
            chips = np.array([l["tx_rel"] * gm.F_CA for l in L])
            cyc = np.array([self.phi0[k] - (gm.F_L1 / gm.C) * (L[k]["pr"] - self._pr0[k]) *
                            self.cfg.doppler_sign for k in range(self.C)])
            fi = np.array([l["fi"] for l in L])
            self._edge_cache[b] = (chips, cyc, fi)
        return self._edge_cache[b]

    # ---- samples -----------------------------------------------------------
    def block(self, b):
        """int16[2*S] interleaved I,Q for block b (receiver time
        [rx_time0 + b T, rx_time0 + (b+1) T))."""
        S, C = self.S, self.C
        c0, p0, _ = self._edge(b)
        c1, p1, _ = self._edge(b + 1)
        u = np.arange(S, dtype=np.float64) / S
        sig = np.zeros(S, dtype=np.complex128)
        for k in range(C):
            chips = c0[k] + (c1[k] - c0[k]) * u
            fl = np.floor(chips)
            ci = np.mod(fl, gm.L_CA).astype(np.int64)
            bit = self.bits[k, np.mod(np.floor(fl / (gm.L_CA * 20)).astype(np.int64), 4096)]
            ph = p0[k] + (p1[k] - p0[k]) * u
            ph = ph - np.floor(ph)
            sig += self.amp * bit * self.codes[k, ci] * np.exp(2j * np.pi * ph)
        rng = np.random.default_rng(self._noise_seed + 7919 * b)
        sig += self.cfg.noise_sigma * (rng.standard_normal(S) + 1j * rng.standard_normal(S))
        iq = np.empty(2 * S, dtype=np.float64)
        iq[0::2] = sig.real
        iq[1::2] = sig.imag
        return np.clip(np.rint(iq), -32768, 32767).astype(np.int16)

    # ---- receiver-side parameters taken from the truth -----------------------
    def channels(self, b):
        """Truth channel parameters of block b: start-referenced (what
        BatchCorrScores consumes) and end-referenced (BatchCorrManifold)."""
        c0, p0, f0 = self._edge(b)
        c1, p1, f1 = self._edge(b + 1)
        T = self.cfg.T
        cp0 = np.floor(c0 / gm.L_CA).astype(np.int64)
        cp1 = np.floor(c1 / gm.L_CA).astype(np.int64)
        return dict(
            prn=np.array(self.cfg.prns, dtype=np.uint8),
            rc_start=c0 - cp0 * gm.L_CA, cp_start=(self.cp_ref + cp0).astype(np.int32),
            rc_end=c1 - cp1 * gm.L_CA, cp_end=(self.cp_ref + cp1).astype(np.int32),
            ri_start=p0 - np.floor(p0), ri_end=p1 - np.floor(p1),
            fc=(c1 - c0) / T, fi=(p1 - p0) / T,
            cp_ref=self.cp_ref.copy(),
            cp_ref_tow=np.full(self.C, self.cfg.tow_ref, dtype=np.int32))

    def handoff(self, b=0):
        """Fields of the reference's handoff CSV (dpinit.cpp:247-400), referenced
        to the start of block b."""
        ch = self.channels(b)
        t = self.cfg.rx_time0 + b * self.cfg.T
        return dict(rxTime=t, X_ECEF=self.rx_state(t), bytes_read=4 * self.S * b,
                    prn_list=ch["prn"], rc=ch["rc_start"], ri=ch["ri_start"], fc=ch["fc"],
                    fi=ch["fi"], cp=ch["cp_start"], cp_timestamp=ch["cp_ref"], TOW=ch["cp_ref_tow"])

    def epoch_inputs(self, b, center=None, time_grid=None, chan_noise=None):
        """Everything one BatchCorrScores + BatchCorrManifold epoch consumes for
        block b, computed from the truth (a perfectly tracking channel manager).

        center: receiver state x_{k|k-1}[8] the grid is centred on (default: the
        truth at the end of the block).  time_grid: clock offsets (m) of the grid's
        time axis (satellite states are rotated per time-grid point,
        cuchanmgr.cu:853-923).  chan_noise: optional dict of offsets added to
        rc/fc/fi (tracking error).
        """
        cfg = self.cfg
        ch = self.channels(b)
        if chan_noise:
            for key, val in chan_noise.items():
                ch[key] = ch[key] + val
        rx_time = cfg.rx_time0 + (b + 1) * cfg.T
        if center is None:
            center = self.rx_state(rx_time)
        center = np.asarray(center, dtype=np.float64)
        if time_grid is None:
            time_grid = np.zeros(1)
        Tn = len(time_grid)
        C = self.C
        tx_time = np.empty(C)
        sat = np.empty((C, 8))
        batch = np.empty((C * Tn, 8))
        for k in range(C):
            tx_time[k] = cfg.tow_ref + ((int(ch["cp_end"][k]) - int(ch["cp_ref"][k])) * gm.T_CA) + \
                (ch["rc_end"][k] / gm.F_CA)
            pos, vel, clkb, clkd = gm.sat_state(self.eph[k], tx_time[k])
            sat[k] = [pos[0], pos[1], pos[2], clkb, vel[0], vel[1], vel[2], clkd]
            for it in range(Tn):
                tau = rx_time - (tx_time[k] + ((time_grid[it] + center[3]) / gm.C)) + clkb
                p, v = gm.rotate_z(pos, vel, tau)
                batch[k * Tn + it] = [p[0], p[1], p[2], clkb, v[0], v[1], v[2], clkd]
        lat, lon = gm.ecef_to_latlon(center)
        return dict(S=self.S, fs=cfg.fs, T=cfg.T, rx_time=rx_time, center=center,
                    enu2ecef=gm.enu_to_ecef_matrix(lat, lon), sat_states=batch, sat_raw=sat,
                    tx_time=tx_time, time_dim=Tn, time_grid=np.asarray(time_grid, np.float64),
                    doppler_sign=cfg.doppler_sign, **ch)

    # ---- files in the reference's formats ------------------------------------
    def write_files(self, out_dir, n_blocks, grid=None, handoff_block=0):
        """samples .dat (int16 I,Q), handoff CSV, grid CSV ('x,y,z,delta_t').

        ``handoff_block``: block the handoff (and ``bytes_read``) refers to.  The
        reference treats ``lseek(fd, StartByte) == 0`` as a failure
        (sampleblock.cu:123-128), so its own runs need handoff_block >= 1."""
        os.makedirs(out_dir, exist_ok=True)
        dat = os.path.join(out_dir, "synthetic_l1ca_%dkHz.dat" % int(self.cfg.fs / 1e3))
        with open(dat, "wb") as f:
            for b in range(n_blocks):
                f.write(self.block(b).tobytes())
        h = self.handoff(handoff_block)
        csv = os.path.join(out_dir, "handoff_params_synth.csv")
        with open(csv, "w") as f:
            f.write("rxTime,%r\n" % float(h["rxTime"]))
            f.write("X_ECEF," + ",".join(repr(float(v)) for v in h["X_ECEF"]) + "\n")
            f.write("bytes_read,%d\n" % h["bytes_read"])
            for key in ("prn_list", "cp", "cp_timestamp", "TOW"):
                f.write(key + "," + ",".join(str(int(v)) for v in h[key]) + "\n")
            for key in ("rc", "ri", "fc", "fi"):
                f.write(key + "," + ",".join(repr(float(v)) for v in h[key]) + "\n")
            f.write("t_oe," + ",".join(str(int(e["toe"])) for e in self.eph) + "\n")
        gpath = None
        if grid is not None:
            gpath = os.path.join(out_dir, "rngrid_synth.csv")
            np.savetxt(gpath, grid, fmt="%.10g", delimiter=",")
        return dict(dat=dat, handoff=csv, grid=gpath, rinex=self.cfg.rinex)
