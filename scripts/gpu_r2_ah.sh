#!/bin/bash
# round 2, call AH (1 GPU): A/B of 6 vs 7 candidates per thread in the two scoring kernels and of programmatic dependent launch
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2ah; mkdir -p $O
for e in "DPE_PDL=0 DPE_LK_CAND=6 DPE_VEL_CAND=6" "DPE_PDL=0" "DPE_PDL=1 DPE_LK_CAND=6 DPE_VEL_CAND=6" "DPE_PDL=1"; do
  env $e timeout 60 python scripts/lookup_wall_probe.py demo 300 >> $O/wall_probe.log 2>&1
  env $e timeout 60 python scripts/lookup_probe.py demo >> $O/stage_probe.log 2>&1
done
cat $O/wall_probe.log; cat $O/stage_probe.log
timeout 100 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -q -x -k "velocity or lookup_scores or full_demo_grid or fold_estimate or submit_collect" > $O/pytest_sub.log 2>&1; tail -3 $O/pytest_sub.log
