#!/bin/bash
# round 2, call AF (1 GPU): chunk partials through fixed-point integer atomics (k_prep_corr, k_carr_partial; no k_carr_finalize):
# full GPU tests, lookup probe, phase stamps
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2af; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; tail -15 $O/pytest_gpu.log
timeout 60 python scripts/lookup_probe.py demo > $O/lookup_probe.log 2>&1; tail -2 $O/lookup_probe.log
DPE_B200_LIB=$PWD/navlab-dpe-sdr_b200/lib/libdpe_b200_phase.so timeout 100 python scripts/phase_probe.py demo > $O/phase_demo.log 2>&1
