#!/bin/bash
# ncu --set full over every kernel of one brute-force epoch and one lookup epoch except k_brute (captured separately)
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/o
timeout 900 ncu --set full --clock-control none -k regex:"k_(prepare|corr_partial|corr_finalize|sample_planes|replica_rd|pair_bins|block_scan|bucket_scan|group_headers|scatter|score_pairs|finalize)" -s 13 -c 13 \
   -o gpurun_out/o/minor_brute python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-both --flow-epochs 0 > gpurun_out/o/minor_brute.out 2>&1
tail -2 gpurun_out/o/minor_brute.out | cut -c1-200
timeout 900 ncu --set full --clock-control none -k regex:"k_(score_lookup|dc_sum|carr_partial|carr_finalize|score_vel|vel_finalize)" -c 8 \
   -o gpurun_out/o/minor_lookup python - > gpurun_out/o/minor_lookup.out 2>&1 <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, dpe_pkg
capi = dpe_pkg.submodule("capi"); synth = dpe_pkg.submodule("synth")
sc = synth.Scenario()
grid = synth.spread_grid(); tg = 6.0 * synth.spread_axis()
vgrid, _ = synth.uniform_grid(25, 0.5)
center = sc.rx_state(sc.cfg.rx_time0 + sc.cfg.T).copy(); center[:4] += (4.0, -3.0, 2.0, 5.0)
ep = sc.epoch_inputs(0, center=center, time_grid=tg); iq = sc.block(0)
ctx = capi.Context(fs=ep["fs"], S=ep["S"], max_chan=sc.C, G=grid.shape[0], time_dim=len(tg), lag_halfwidth=16, Gv=vgrid.shape[0])
ctx.grid_set(grid); ctx.vel_grid_set(vgrid)
for _ in range(2):
    r = ctx.epoch_run(iq, ep, score_mode=capi.SCORE_LOOKUP, with_vel=True)
print(r.z[:8], r.argmax, r.vel_argmax)
PY
tail -3 gpurun_out/o/minor_lookup.out | cut -c1-300
ls -la gpurun_out/o
