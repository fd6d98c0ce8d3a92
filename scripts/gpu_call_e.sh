#!/bin/bash
# bench (default invocation) + ncu evidence for the current kernels
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/prof
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 3000 gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -c 800 gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
   --log-file gpurun_out/prof/launches_demo_v3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-both --flow-epochs 0 \
   > gpurun_out/prof/launches_demo_v3.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_brute -s 1 -c 1 \
   -o gpurun_out/prof/k_brute_v3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-both --flow-epochs 0 \
   > gpurun_out/prof/k_brute_v3.out 2>&1
tail -2 gpurun_out/prof/k_brute_v3.out
