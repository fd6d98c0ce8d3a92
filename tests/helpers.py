"""Shared scenario / oracle plumbing for the parity tests."""
import functools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import dpe_pkg  # noqa: E402
from oracle import dpe_oracle as orc  # noqa: E402

synth = dpe_pkg.submodule("synth")


@functools.lru_cache(maxsize=8)
def scenario(fs=2.5e6, prns=synth.PRNS_8, seed=20180704, cn0=45.0):
    return synth.Scenario(synth.ScenarioConfig(fs=fs, prns=tuple(prns), seed=seed, cn0_dbhz=cn0))


@functools.lru_cache(maxsize=16)
def epoch_case(fs=2.5e6, prns=synth.PRNS_8, block=0, n=9, spacing=(5.0, 5.0, 5.0, 6.0), seed=20180704,
               center_offset=(0.0, 0.0, 0.0, 0.0)):
    """(scenario, iq, grid, epoch dict) for one block with an n^4 uniform grid."""
    sc = scenario(fs, prns, seed)
    grid, tg = synth.uniform_grid(n, spacing)
    rx_time = sc.cfg.rx_time0 + (block + 1) * sc.cfg.T
    center = sc.rx_state(rx_time).copy()
    center[:4] += np.asarray(center_offset)
    ep = sc.epoch_inputs(block, center=center, time_grid=tg)
    iq = sc.block(block)
    return sc, iq, grid, ep


@functools.lru_cache(maxsize=16)
def oracle_bcs(fs=2.5e6, prns=synth.PRNS_8, block=0, seed=20180704):
    sc = scenario(fs, prns, seed)
    ep = sc.epoch_inputs(block)
    iq = sc.block(block)
    return orc.batch_corr_scores(iq, ep["prn"], ep["rc_start"], ep["ri_start"], ep["fc"], ep["fi"],
                                 ep["cp_start"], ep["cp_ref"], ep["fs"])


def oracle_pos(bcs, grid, ep, lpower=1, weighted=False, per_time=False):
    args = (bcs["code_scores"], grid, ep["center"], ep["enu2ecef"], ep["sat_states"], ep["time_dim"],
            ep["fc"], ep["rc_end"], ep["cp_ref_tow"], ep["cp_end"], ep["cp_ref"], ep["rx_time"], ep["fs"],
            ep["S"])
    if weighted:
        return orc.pos_meas_weighted(*args, lpower=lpower, per_time_sat=per_time)
    return orc.pos_meas_ml(*args, lpower=lpower)
