#!/bin/bash
# 2-GPU check of the final build (presort on a second stream, tile-granular split): demo and c3
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/s
for w in demo c3; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 \
   bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-both --workload $w --flow-epochs 0 \
   > gpurun_out/s/bench_${w}_n2.json 2> gpurun_out/s/bench_${w}_n2.err
python -c "
import json
for l in open('gpurun_out/s/bench_${w}_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$w n2', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['stage_ms_per_step'], d['fix'])"
grep -i "error" gpurun_out/s/bench_${w}_n2.err | tail -3
done
timeout 300 python bench.py --impl reference --gpus 2 --steps 2 --warmup 1 | cut -c1-300
