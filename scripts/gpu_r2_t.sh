#!/bin/bash
# round 2, call T (1 GPU): ncu --set full with source of the lookup-path kernels
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2t; mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_prep_corr|k_score_lookup|k_score_vel|k_carr_partial" -s 8 -c 4 -o $O/lookup_kernels -f \
   python scripts/lookup_probe.py demo > $O/ncu.log 2>&1; tail -2 $O/ncu.log
python scripts/ncu_summary.py $O/lookup_kernels.ncu-rep > $O/lookup_kernels_ncu_summary.txt; cat $O/lookup_kernels_ncu_summary.txt
