#!/usr/bin/env python
"""Wall time per lookup epoch, one epoch at a time, through the two call paths (no event brackets, so that nothing sits between
the kernels): dpe_epoch_run (kernel by kernel on the caller's stream, like the console flow's stages) and
dpe_epoch_submit / dpe_epoch_collect (one CUDA-graph launch).  Position manifold alone and with the 25^4 velocity manifold."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import dpe_pkg
capi = dpe_pkg.submodule("capi"); synth = dpe_pkg.submodule("synth")
name = sys.argv[1] if len(sys.argv) > 1 else "demo"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
sc, grid, tg = bench.build_workload(name)
vgrid, _ = synth.uniform_grid(25, 0.5)
ep = bench.epoch_for_block(sc, 0, tg); iq = sc.block(0)
ctx = capi.Context(fs=sc.cfg.fs, S=sc.S, max_chan=sc.C, G=grid.shape[0], time_dim=len(tg), lag_halfwidth=16,
                   Gv=vgrid.shape[0], dopp_halfwidth=64)
ctx.grid_set(grid); ctx.vel_grid_set(vgrid)
out = []
for with_vel in (0, 1):
    for path in ("run", "graph"):
        f = (lambda: ctx.epoch_run(iq, ep, with_vel=with_vel)) if path == "run" else (lambda: ctx.epoch_run_dist(iq, ep, with_vel=with_vel))
        for _ in range(20): r = f()
        best = 1e9
        for rep in range(3):
            t0 = time.perf_counter()
            for _ in range(n): r = f()
            best = min(best, (time.perf_counter() - t0) / n)
        out.append("%s%s %.1f" % (path, "+vel" if with_vel else "", best * 1e6))
print(name, "us/epoch:", " | ".join(out), "| argmax", r.argmax, r.vel_argmax, "env", {k: v for k, v in os.environ.items() if k.startswith("DPE_") and k != "DPE_B200_LIB"})
