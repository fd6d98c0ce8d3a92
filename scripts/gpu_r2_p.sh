#!/bin/bash
# round 2, call P (8 GPUs): contexts in flight per rank, 2 vs 3 (demo, same box)
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2p; mkdir -p $O
for d in 2 3 2 3; do
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2970$d \
   bench.py --gpus 8 --steps 40 --warmup 4 --depth $d --no-cpu-baseline --flow-epochs 0 --configs none --no-both 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('depth', $d, 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['ms_per_step'], 4), 'k_brute', round(d['roofline']['kernel_ms'], 4), 'lat', round(d['latency']['ms_per_epoch'], 4))"
done
