#!/usr/bin/env python
"""Phase stamps of the two lookup-path kernels (probe build: make -C navlab-dpe-sdr_b200/csrc VARIANT=phase EXTRA=-DDPE_PHASE_TIMING,
run with DPE_B200_LIB=navlab-dpe-sdr_b200/lib/libdpe_b200_phase.so).  Prints the kernels' "PT ..." lines of the last epoch
(ns since the CTA's own start, %globaltimer) and the event-bracketed stage times of the same build."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import dpe_pkg
capi = dpe_pkg.submodule("capi"); synth = dpe_pkg.submodule("synth")
name = sys.argv[1] if len(sys.argv) > 1 else "demo"
sc, grid, tg = bench.build_workload(name)
ep = bench.epoch_for_block(sc, 0, tg); iq = sc.block(0)
ctx = capi.Context(fs=sc.cfg.fs, S=sc.S, max_chan=sc.C, G=grid.shape[0], time_dim=len(tg), lag_halfwidth=16)
ctx.grid_set(grid)
for _ in range(3): ctx.epoch_run(iq, ep)
sys.stdout.flush()
print("==== last epoch ====", flush=True)
ctx.profile_enable(True)
ctx.epoch_run(iq, ep)
ms, cnt = ctx.profile_read()
print("stage us: prepare %.1f lookup %.1f" % (1e3 * ms[0], 1e3 * ms[2]), flush=True)
ctx.close()
