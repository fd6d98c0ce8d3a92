"""CPU tests: the oracle against independent known answers and the reference's
checked-in demo inputs (the reference ships no golden vectors)."""
import math
import os

import numpy as np
import pytest

import helpers as H
from helpers import orc, synth
from oracle import chanmgr_oracle as chm

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# IS-GPS-200 table 3-I: first 10 C/A chips of PRN 1..32, octal
FIRST10 = [0o1440, 0o1620, 0o1710, 0o1744, 0o1133, 0o1455, 0o1131, 0o1454, 0o1626, 0o1504, 0o1642, 0o1750,
           0o1764, 0o1772, 0o1775, 0o1776, 0o1156, 0o1467, 0o1633, 0o1715, 0o1746, 0o1763, 0o1063, 0o1706,
           0o1743, 0o1761, 0o1770, 0o1774, 0o1127, 0o1453, 0o1625, 0o1712]


def test_ca_code_first_ten_chips_is_gps_200():
    for prn in range(1, 33):
        code = orc.gen_ca_code(prn)
        bits = 0
        for c in code[:10]:
            bits = (bits << 1) | (1 if c > 0 else 0)
        assert bits == FIRST10[prn - 1], "PRN %d" % prn


def test_ca_code_two_constructions_agree_and_balance():
    gm = H.dpe_pkg.submodule("gpsmath")
    for prn in range(1, 38):
        a, b = orc.gen_ca_code(prn), gm.ca_code(prn)
        assert np.array_equal(a, b), "PRN %d" % prn
        assert abs(int(a.sum())) == 1                      # Gold code balance: 512 vs 511
        ac = np.array([np.dot(a, np.roll(a, k)) for k in (1, 7, 100, 511)])
        assert set(ac.tolist()) <= {-1, 63, -65}            # three-valued autocorrelation


def test_time_index_rounding_is_a_noop_at_2p5_and_10_mhz():
    for fs, S in ((2.5e6, 50000), (10e6, 200000)):
        t = orc.time_idcs(S, fs)
        assert np.array_equal(t, np.arange(S) / fs)


def test_nav_bit_boundary_and_flipped_replica():
    idx = orc.nav_bit_boundary([1003], [1000], [100.25], [1.023e6], 2.5e6)
    # 17 periods to the next bit: floor((1023*17 - 100.25) * 2.5e6/1.023e6) + 1
    assert idx[0] == math.floor((1023 * 17 - 100.25) * (2.5e6 / 1.023e6)) + 1
    t = orc.time_idcs(1000, 2.5e6)
    ci, nf, fl = orc.code_replica(5, t, 1.023e6, 10.5, 400, 1000)
    assert np.array_equal(fl[:400], nf[:400]) and np.array_equal(fl[400:], -nf[400:])
    _, _, fl0 = orc.code_replica(5, t, 1.023e6, 10.5, 1000, 1000)     # edge outside the block
    assert not fl0.any()
    assert ci[0] == 10 and ci.min() >= 0 and ci.max() <= 1022


def test_fft_correlogram_equals_direct_circular_correlation():
    rng = np.random.default_rng(1)
    S = 512
    xw = rng.standard_normal(S) + 1j * rng.standard_normal(S)
    r = rng.integers(0, 2, S) * 2.0 - 1.0
    c_fft = np.fft.ifft(np.fft.fft(xw) * np.conj(np.fft.fft(r)))
    for k in (0, 1, 5, -3, 200):
        direct = np.sum(np.roll(xw, -k) * r)                 # sum_n xw[(n+k) mod S] r[n]
        assert abs(c_fft[k % S] - direct) < 1e-9
        # SURVEY 8 a': the lerp of two bins is a correlation against the blended replica
        a = 0.37
        v = (1 - a) * c_fft[k % S] + a * c_fft[(k + 1) % S]
        assert abs(v - orc.blended_correlation(xw, r, k, a)) < 1e-9


def test_uniform_and_arthur_grids():
    g, tg = orc.init_pos_grid([3, 3, 3, 5], [1.0, 2.0, 3.0, 4.0])
    assert g.shape == (135, 4) and np.array_equal(tg, [-8, -4, 0, 4, 8])
    assert np.array_equal(g[0], [-1, -2, -3, -8]) and np.array_equal(g[1], [-1, -2, -3, -4])   # t fastest
    assert np.array_equal(g[67], [0, 0, 0, 0])
    gs, _ = synth.uniform_grid(3, (1.0, 2.0, 3.0, 4.0))
    assert np.array_equal(gs[:, :3], orc.init_pos_grid([3, 3, 3, 3], [1.0, 2.0, 3.0, 4.0])[0][:, :3])
    ga, _ = orc.init_pos_grid([9, 9, 9, 9], [1.0] * 4, orc.GRID_ARTHURBASIS)
    ax = np.unique(ga[:, 0])
    assert len(ax) == 9 and ax[4] == 0 and np.allclose(ax, -ax[::-1])


def test_geodetic_known_answers_round_trip():
    # pygnss utils.py:23-26 docstring values (ECE building, Mount Everest)
    for lat, lon, alt in ((40.11497089608554, -88.22793631642435, 203.9925799164921),
                          (27.98805865809616, 86.92527453636706, 8847.923165871762)):
        la, lo = math.radians(lat), math.radians(lon)
        a, e2 = 6378137.0, orc.CONST_WGS84_E ** 2
        N = a / math.sqrt(1 - e2 * math.sin(la) ** 2)
        p = [(N + alt) * math.cos(la) * math.cos(lo), (N + alt) * math.cos(la) * math.sin(lo),
             (N * (1 - e2) + alt) * math.sin(la)]
        la2, lo2 = chm.ecef2ll_rad(p)
        assert abs(la2 - la) < 1e-9 and abs(lo2 - lo) < 1e-12
        R = chm.enu2ecef_mat(la2, lo2).reshape(3, 3)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-14)
        up = np.array(p) / np.linalg.norm(p)
        assert R[:, 2] @ up > 0.999


def test_demo_handoff_and_rinex_are_mutually_consistent():
    """The reference's two checked-in demo inputs: the code phase back-calculated from
    ephemeris + handoff state (cuchanmgr.cu:85-210, batchcorrmanifold.cu:1779-1790) lands
    within 0.006 chip of the handed-off code phase for all 8 PRNs."""
    nav = chm.read_rinex_nav(synth.DEFAULT_RINEX)
    h = chm.read_handoff(os.path.join(GOLDEN, "handoff_params_usrp6.csv"))
    assert list(h["prn_list"]) == [2, 3, 6, 12, 17, 19, 24, 28]
    x = h["X_ECEF"]
    for i, prn in enumerate(h["prn_list"]):
        tx = chm.tx_time_of(int(h["TOW"][i]), int(h["cp"][i]), int(h["cp_timestamp"][i]), h["rc"][i])
        eph = chm.select_eph(nav, int(prn), tx)
        assert eph is not None and eph.toes == 417600.0
        sat = chm.get_sat_pos(eph, tx)
        tau = h["rxTime"] - (tx + x[3] / orc.CONST_C) + sat[3]
        rot = chm.rotate_sat(sat, tau)
        rng = math.dist(rot[:3], x[:3])
        assert 2.0e7 < rng < 2.4e7
        pr = rng - orc.CONST_C * rot[3] + x[3]
        bc_rc = (h["rxTime"] - pr / orc.CONST_C - h["TOW"][i] -
                 (int(h["cp"][i]) - int(h["cp_timestamp"][i])) * 1e-3) * 1.023e6
        assert abs(bc_rc - h["rc"][i]) < 0.006, "PRN %d: %.4f chip" % (prn, bc_rc - h["rc"][i])
        assert abs((h["fc"][i] - 1.023e6) - h["fi"][i] * 1.023e6 / 1.57542e9) < 1.0


def test_scenario_is_self_consistent_and_oracle_recovers_truth():
    sc, iq, grid, ep = H.epoch_case(n=7, center_offset=(10.0, -5.0, 5.0, 6.0))
    bcs = H.oracle_bcs()
    S = ep["S"]
    # every channel's prompt lag carries a clear peak
    for c in range(sc.C):
        mag = np.abs(bcs["code_scores"][c])
        assert mag[S // 2] > 8 * np.median(mag)
    r = H.oracle_pos(bcs, grid, ep)
    assert r["valid"].all()
    truth = sc.rx_state(ep["rx_time"])
    assert np.linalg.norm(r["z"][:3] - truth[:3]) < 9.0 and abs(r["z"][3] - truth[3]) < 12.0
    w = H.oracle_pos(bcs, grid, ep, weighted=True, per_time=True)
    assert np.linalg.norm(w["z"][:3] - ep["center"][:3]) < 10.0


def test_channel_manager_oracle_tracks_truth_over_epochs():
    """cuChanMgr restatement: start from a truth handoff, propagate 3 epochs with the
    true state as the fix; code phases must stay on the scenario's truth."""
    sc = H.scenario()
    nav = chm.read_rinex_nav(sc.cfg.rinex)
    h = sc.handoff(0)
    ch = chm.chanmgr_start(nav, h, sc.cfg.T, h["X_ECEF"])
    for b in range(3):
        truth = sc.channels(b)
        assert np.max(np.abs(ch.rc_start - truth["rc_start"])) < 2e-3
        assert np.max(np.abs(ch.rc_end - truth["rc_end"])) < 0.02
        assert np.array_equal(ch.cp_end, truth["cp_end"])
        sat, R = chm.grid_prep(ch, sc.rx_state(ch.rx_time), np.zeros(1))
        assert sat.shape == (sc.C, 8) and R.shape == (9,)
        chm.chanmgr_update(ch, nav, sc.rx_state(ch.rx_time))
