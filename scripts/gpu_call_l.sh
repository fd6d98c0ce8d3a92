#!/bin/bash
# ncu evidence for the current k_brute: full set on one launch + launch list of a short bench run
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/l
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_brute -s 1 -c 1 \
   -o gpurun_out/l/k_brute_v4 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-both --flow-epochs 0 \
   > gpurun_out/l/k_brute_ncu.out 2>&1
tail -3 gpurun_out/l/k_brute_ncu.out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/l/launches_demo.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-both --flow-epochs 0 \
   > gpurun_out/l/launches_demo.out 2>&1
tail -2 gpurun_out/l/launches_demo.out
ls -la gpurun_out/l
