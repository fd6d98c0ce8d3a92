#!/bin/bash
# round 2, call N (1 GPU): A/B of the padding-group skip in k_brute, whole grid and a 1/8 shard, same box
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2n; mkdir -p $O
for sh in 0 8; do for sp in 0 1 0 1; do
DPE_BENCH_SHARD_OF=$sh DPE_BRUTE_SKIP_PAD=$sp timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --flow-epochs 0 --configs none --no-both 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('shard_of', $sh, 'skip_pad', $sp, 'ms/step', round(d['ms_per_step'], 4), 'k_brute', round(d['roofline']['kernel_ms'], 4), 'lat', round(d['latency']['ms_per_epoch'], 4))"
done; done
