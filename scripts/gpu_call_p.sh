#!/bin/bash
# quick: parity tests + brute and lookup bench lines (stage times)
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/p
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/p/pytest_gpu.log 2>&1; tail -3 gpurun_out/p/pytest_gpu.log
for path in brute lookup; do
timeout 300 python bench.py --path $path --steps 20 --warmup 5 --no-cpu-baseline --no-both --flow-epochs 0 > gpurun_out/p/bench_$path.json 2> gpurun_out/p/bench_$path.err
python -c "
import json;d=json.load(open('gpurun_out/p/bench_$path.json'));print('$path',d['ms_per_step'],d['e2e']['ms_per_step'],d['stage_ms_per_step'],d['roofline']['achieved'],d['roofline']['frac'])"
done
