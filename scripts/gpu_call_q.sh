#!/bin/bash
# Final round-1 evidence: default bench line, reference arm, ncu full capture of k_brute, launch list.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/q
timeout 900 python bench.py > gpurun_out/q/bench_default.json 2> gpurun_out/q/bench_default.err
tail -c 1500 gpurun_out/q/bench_default.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/q/bench_reference.json 2> gpurun_out/q/bench_reference.err
tail -c 600 gpurun_out/q/bench_reference.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_brute -s 1 -c 1 \
   -o gpurun_out/q/k_brute_r01j python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-both --flow-epochs 0 \
   > gpurun_out/q/k_brute_ncu.out 2>&1
tail -3 gpurun_out/q/k_brute_ncu.out | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/q/launches_demo.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-both --flow-epochs 0 \
   > gpurun_out/q/launches_demo.out 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/q/launches_lookup.csv python bench.py --path lookup --steps 2 --warmup 1 --no-cpu-baseline --no-both --flow-epochs 0 \
   > gpurun_out/q/launches_lookup.out 2>&1
ls -la gpurun_out/q
for w in c3 c4; do
timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-both --flow-epochs 0 > gpurun_out/q/bench_${w}_n1.json 2> gpurun_out/q/bench_${w}_n1.err
python -c "
import json;d=json.load(open('gpurun_out/q/bench_${w}_n1.json'));print('$w',d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['stage_ms_per_step'],d['roofline']['achieved'],d['roofline']['frac'])"
done
