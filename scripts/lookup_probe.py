#!/usr/bin/env python
"""Per-stage device milliseconds of the lookup (reference-formulation) path incl. the velocity manifold."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import dpe_pkg
capi = dpe_pkg.submodule("capi"); synth = dpe_pkg.submodule("synth")
for name in (sys.argv[1:] or ["demo"]):
    sc, grid, tg = bench.build_workload(name)
    vgrid, _ = synth.uniform_grid(25, 0.5)
    ep = bench.epoch_for_block(sc, 0, tg); iq = sc.block(0)
    ctx = capi.Context(fs=sc.cfg.fs, S=sc.S, max_chan=sc.C, G=grid.shape[0], time_dim=len(tg), lag_halfwidth=16,
                       Gv=vgrid.shape[0], dopp_halfwidth=64)
    ctx.grid_set(grid); ctx.vel_grid_set(vgrid)
    for _ in range(3): ctx.epoch_run(iq, ep, with_vel=1)
    ctx.profile_enable(True)
    n = 50
    t0 = time.perf_counter()
    for _ in range(n): ctx.epoch_run(iq, ep, with_vel=1)
    wall = (time.perf_counter() - t0) / n
    ms, cnt = ctx.profile_read()
    names = ("prepare", "correlogram", "lookup", "bins", "brute", "bscore", "estimate", "velocity")
    print(name, "wall/epoch %.1f us |" % (wall * 1e6), " ".join("%s %.1f" % (k, 1e3 * ms[i] / n) for i, k in enumerate(names) if ms[i] > 0), "(us)")
