#!/bin/bash
# round 2, call AG (1 GPU): A/B of the correlogram chunk (1024 vs 512 samples per CTA, now that the chunk partials meet in
# integer atomics) + the batched reduction of the block partials; phase stamps of the 512 build
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2ag; mkdir -p $O
L=$PWD/navlab-dpe-sdr_b200/lib
for v in "" _c512; do
  for i in 1 2; do DPE_B200_LIB=$L/libdpe_b200$v.so timeout 60 python scripts/lookup_probe.py demo >> $O/lookup_probe$v.log 2>&1; done
  tail -2 $O/lookup_probe$v.log
done
DPE_B200_LIB=$L/libdpe_b200_c512.so timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "correlogram or velocity or lookup_scores or chip_index" > $O/pytest_c512.log 2>&1; tail -3 $O/pytest_c512.log
DPE_B200_LIB=$L/libdpe_b200_phase.so timeout 100 python scripts/phase_probe.py demo > $O/phase_demo_c512.log 2>&1
