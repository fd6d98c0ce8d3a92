#!/bin/bash
# GPU call H: A/B of the k_brute blend (all-FMA vs LOP3 on the ALU pipe), parity tests on the new default.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/h
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/h/pytest_gpu.log 2>&1; tail -3 gpurun_out/h/pytest_gpu.log
for b in fma lop; do
  DPE_BRUTE_BLEND=$b timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-both --flow-epochs 0 \
     > gpurun_out/h/bench_demo_$b.json 2> gpurun_out/h/bench_demo_$b.err
  python -c "
import json;d=json.load(open('gpurun_out/h/bench_demo_$b.json'));print('$b',d['ms_per_step'],d['value'],d['roofline']['kernel_ms'],d['roofline']['achieved'],d['clocks'])"
done
DPE_BRUTE_BLEND=lop timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-both --flow-epochs 0 --workload c3 \
     > gpurun_out/h/bench_c3_lop.json 2> gpurun_out/h/bench_c3_lop.err
tail -c 600 gpurun_out/h/bench_c3_lop.json
