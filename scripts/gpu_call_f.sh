#!/bin/bash
# 8-GPU box: scaling of demo (strong), c4 (51^4 @ 10 MHz sharded over 8) and c5 (256 streams, 32 per GPU)
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/scale
run() {  # n workload steps warmup port
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $5 \
     bench.py --gpus $1 --steps $3 --warmup $4 --no-cpu-baseline --workload $2 --flow-epochs 0 \
     > gpurun_out/scale/bench_$2_n$1.json 2> gpurun_out/scale/bench_$2_n$1.err
  tail -c 400 gpurun_out/scale/bench_$2_n$1.json; grep -i "error" gpurun_out/scale/bench_$2_n$1.err | tail -3
}
run 4 demo 10 3 29601
run 8 demo 10 3 29602
run 8 c4 5 3 29603
run 8 c5 2 1 29604
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv > gpurun_out/scale/smi.txt
