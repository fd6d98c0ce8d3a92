#!/usr/bin/env python
"""Device-side probe of the brute-force velocity manifold (k_brute_vel) on the 25^4 velocity grid the console flow
uses: velocity-stage milliseconds for the lookup and the brute-force formulation, FP32 rate, same fix.  (Tuning / ncu
target; bench.py's `velocity_brute` record is the reported number.)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import dpe_pkg

capi = dpe_pkg.submodule("capi")
synth = dpe_pkg.submodule("synth")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
peak = capi.microbench_fp32(0, True)
sc, grid, tg = bench.build_workload("demo")
vgrid, _ = synth.uniform_grid(n, (0.5, 0.5, 0.5, 0.25))
grid = np.ascontiguousarray(grid[:4096])
ep = bench.epoch_for_block(sc, 0, tg)
iq = sc.block(0)
ctx = capi.Context(fs=sc.cfg.fs, S=sc.S, max_chan=sc.C, G=grid.shape[0], time_dim=len(tg), lag_halfwidth=16,
                   Gv=vgrid.shape[0], dopp_halfwidth=64, flags=capi.FLAG_BRUTE_VEL)
ctx.grid_set(grid)
ctx.vel_grid_set(vgrid)
fix = {}
for mode, name in ((1, "lookup"), (2, "brute")):
    r = ctx.epoch_run(iq, ep, with_vel=mode)
    ctx.profile_enable(True)
    for _ in range(reps):
        r = ctx.epoch_run(iq, ep, with_vel=mode)
    ms, cnt = ctx.profile_read()
    ctx.profile_enable(False)
    v = ms[capi.STAGE_VELOCITY] / reps
    pairs = vgrid.shape[0] * sc.C - r.vel_out_of_window
    tf = 12.0 * sc.S * pairs / (v * 1e-3) / 1e12
    fix[name] = (r.vel_argmax, r.vel_max_score)
    print("%s: %d velocity candidates x %d PRNs, velocity stage %.3f ms, oow %d%s, argmax %d, max %.6e"
          % (name, vgrid.shape[0], sc.C, v, r.vel_out_of_window,
             (", %.2f TFLOP/s = %.1f%% of the FFMA2 peak %.1f" % (tf, 100 * tf / peak, peak)) if mode == 2 else "",
             r.vel_argmax, r.vel_max_score))
assert fix["lookup"][0] == fix["brute"][0]
print("rel diff of the max score %.3g" % (abs(fix["lookup"][1] - fix["brute"][1]) / fix["lookup"][1]))
ctx.close()
