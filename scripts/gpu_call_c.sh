#!/bin/bash
# GPU call C: golden vectors from the unmodified reference + the reference's own epoch time on B200.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/golden
timeout 900 python oracle/make_golden_ref.py --out gpurun_out/golden > gpurun_out/golden.log 2>&1
tail -25 gpurun_out/golden.log
# reference kernels on B200 at the demo shape (25^4 + 25^4 uniform grids, 8 PRNs), no dumps
R=navlab-dpe-sdr_b200/data/brdc_toe417600.18n
timeout 600 oracle/_ref/ref_dpe /tmp/refrun/synthetic_l1ca_2500kHz.dat /tmp/refrun/handoff_params_synth.csv $R none \
   25 25 30 /tmp/refrun/dump25 32 2.5e6 0 > gpurun_out/ref_timing_demo.log 2>&1
tail -8 gpurun_out/ref_timing_demo.log
