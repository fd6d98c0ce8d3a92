#!/bin/bash
# 8-GPU box: demo strong scaling 1/2/4/8 with the split tail slots, c4 (51^4 @ 10 MHz) on 8
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/scale4
run() {  # n workload steps warmup port
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $5 \
     bench.py --gpus $1 --steps $3 --warmup $4 --no-cpu-baseline --no-both --workload $2 --flow-epochs 0 \
     > gpurun_out/scale4/bench_$2_n$1.json 2> gpurun_out/scale4/bench_$2_n$1.err
  python -c "
import json
for l in open('gpurun_out/scale4/bench_$2_n$1.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$2 n$1', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['stage_ms_per_step'], d['roofline']['achieved'])"
  grep -i "error" gpurun_out/scale4/bench_$2_n$1.err | tail -3
}
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-both --flow-epochs 0 > gpurun_out/scale4/bench_demo_n1.json 2> gpurun_out/scale4/bench_demo_n1.err
python -c "
import json;d=json.load(open('gpurun_out/scale4/bench_demo_n1.json'));print('demo n1',d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['stage_ms_per_step'])"
run 2 demo 10 3 29611
run 4 demo 10 3 29612
run 8 demo 10 3 29613
run 8 c4 5 3 29614
run 8 c5 2 1 29615
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv > gpurun_out/scale4/smi.txt
