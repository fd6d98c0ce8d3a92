#!/bin/bash
# GPU call J: quick brute parity + demo bench (kernel iteration loop)
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/j
timeout 600 python -m pytest tests/test_gpu_properties.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/j/pytest_gpu.log 2>&1; tail -3 gpurun_out/j/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-both --flow-epochs 0 \
     > gpurun_out/j/bench_demo.json 2> gpurun_out/j/bench_demo.err
python -c "
import json;d=json.load(open('gpurun_out/j/bench_demo.json'));print(d['ms_per_step'],d['value'],d['roofline']['kernel_ms'],d['roofline']['achieved'],d['roofline']['frac'],d['clocks'])"
