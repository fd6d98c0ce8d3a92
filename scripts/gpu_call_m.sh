#!/bin/bash
# Final round-1 evidence: default bench line, reference arm, ncu full capture of k_brute, launch list.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/m
timeout 900 python bench.py > gpurun_out/m/bench_default.json 2> gpurun_out/m/bench_default.err
tail -c 1500 gpurun_out/m/bench_default.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/m/bench_reference.json 2> gpurun_out/m/bench_reference.err
tail -c 600 gpurun_out/m/bench_reference.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_brute -s 1 -c 1 \
   -o gpurun_out/m/k_brute_r01h python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-both --flow-epochs 0 \
   > gpurun_out/m/k_brute_ncu.out 2>&1
tail -3 gpurun_out/m/k_brute_ncu.out | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/m/launches_demo.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-both --flow-epochs 0 \
   > gpurun_out/m/launches_demo.out 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/m/launches_lookup.csv python bench.py --path lookup --steps 2 --warmup 1 --no-cpu-baseline --no-both --flow-epochs 0 \
   > gpurun_out/m/launches_lookup.out 2>&1
ls -la gpurun_out/m
