#!/bin/bash
# round 2, call AE (1 GPU): where the time of the two single-wave lookup kernels goes -- phase stamps of the probe build
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2ae; mkdir -p $O
DPE_B200_LIB=$PWD/navlab-dpe-sdr_b200/lib/libdpe_b200_phase.so timeout 100 python scripts/phase_probe.py demo > $O/phase_demo.log 2>&1
grep -c "^PT" $O/phase_demo.log; tail -3 $O/phase_demo.log
timeout 60 python scripts/lookup_probe.py demo > $O/lookup_probe.log 2>&1; tail -2 $O/lookup_probe.log
