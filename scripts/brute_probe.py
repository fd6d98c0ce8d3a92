#!/usr/bin/env python
"""Quick device-side probe of the brute-force path: per-stage milliseconds and the k_brute
FP32 rate on a workload (used while tuning; bench.py is the reported number)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import dpe_pkg

capi = dpe_pkg.submodule("capi")
names = sys.argv[1:] or ["c3", "demo"]
peak = capi.microbench_fp32(0, True)
print("FFMA2 peak %.2f TFLOP/s" % peak)
for name in names:
    sc, grid, tg = bench.build_workload(name)
    ep = bench.epoch_for_block(sc, 0, tg)
    iq = sc.block(0)
    ctx = capi.Context(fs=sc.cfg.fs, S=sc.S, max_chan=sc.C, G=grid.shape[0], time_dim=len(tg), lag_halfwidth=16,
                       flags=capi.FLAG_BRUTE_TILES)
    ctx.grid_set(grid)
    for _ in range(2):
        r = ctx.epoch_run(iq, ep, score_mode=capi.SCORE_BRUTE)
    ctx.profile_enable(True)
    n = 5
    for _ in range(n):
        r = ctx.epoch_run(iq, ep, score_mode=capi.SCORE_BRUTE)
    ms, cnt = ctx.profile_read()
    pairs = ctx.brute_pairs()
    kb = ms[capi.STAGE_BRUTE_CORR] / n
    tf = 6.0 * sc.S * pairs / (kb * 1e-3) / 1e12
    print("%s: pairs %d  stages(ms/epoch) prepare %.3f corr %.3f bins %.3f k_brute %.3f score %.3f est %.3f | "
          "k_brute %.2f TFLOP/s = %.1f%% of FFMA2 peak | argmax %d" %
          (name, pairs, ms[0] / n, ms[1] / n, ms[3] / n, kb, ms[5] / n, ms[6] / n, tf, 100 * tf / peak, r.argmax))
    ctx.close()
