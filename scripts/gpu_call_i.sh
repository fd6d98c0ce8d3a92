#!/bin/bash
# GPU call I: parity tests with the split tail slots, demo + c3 bench lines.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/i
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/i/pytest_gpu.log 2>&1; tail -15 gpurun_out/i/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-both --flow-epochs 0 \
     > gpurun_out/i/bench_demo.json 2> gpurun_out/i/bench_demo.err
python -c "
import json;d=json.load(open('gpurun_out/i/bench_demo.json'));print(d['ms_per_step'],d['value'],d['roofline'],d['clocks'])"
