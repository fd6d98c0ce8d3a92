#!/bin/bash
# 2-GPU bench (sharded grid, NCCL broadcast + all-gather) and the 1-GPU bench lines.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_demo_n1.json 2> gpurun_out/bench_demo_n1.err
tail -c 600 gpurun_out/bench_demo_n1.json; tail -3 gpurun_out/bench_demo_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_demo_n2.json 2> gpurun_out/bench_demo_n2.err
tail -c 1500 gpurun_out/bench_demo_n2.json; tail -5 gpurun_out/bench_demo_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --workload c4 > gpurun_out/bench_c4_n2.json 2> gpurun_out/bench_c4_n2.err
tail -c 1500 gpurun_out/bench_c4_n2.json; tail -5 gpurun_out/bench_c4_n2.err
