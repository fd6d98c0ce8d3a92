#!/bin/bash
# BASELINE config 2 at full length: 45 s (2250 epochs) of the synthetic demo capture through the dpe_console flow,
# brute-force and lookup BCM; plus the GPU parity tests and smoke() on the final build.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/r
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r/pytest_gpu.log 2>&1; tail -2 gpurun_out/r/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r/smoke.log 2>&1; tail -2 gpurun_out/r/smoke.log
timeout 1200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-both --flow-epochs 2250 > gpurun_out/r/bench_flow45s.json 2> gpurun_out/r/bench_flow45s.err
python -c "
import json
for l in open('gpurun_out/r/bench_flow45s.json'):
    if l.startswith('{'):
        d=json.loads(l); print(json.dumps(d['flow']))"
tail -3 gpurun_out/r/bench_flow45s.err
