#!/bin/bash
# 8-GPU box, final build (presort stream): demo at 1 and 8 GPUs + the GPU parity tests
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/t
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t/pytest_gpu.log 2>&1; tail -2 gpurun_out/t/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-both --flow-epochs 0 > gpurun_out/t/bench_demo_n1.json 2> gpurun_out/t/bench_demo_n1.err
for n in 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2963$n \
   bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --no-both --flow-epochs 0 > gpurun_out/t/bench_demo_n$n.json 2> gpurun_out/t/bench_demo_n$n.err
done
for n in 1 4 8; do python -c "
import json
for l in open('gpurun_out/t/bench_demo_n$n.json'):
    if l.startswith('{'):
        d=json.loads(l); print('demo n$n', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['stage_ms_per_step'])"; done
