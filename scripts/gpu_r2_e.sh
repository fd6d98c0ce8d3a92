#!/bin/bash
# round 2, call E (2 GPUs): SMs reserved for the NCCL kernels under k_brute -- 0 / 1 / 2
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2e; mkdir -p $O
for r in 0 1 2; do
DPE_COMM_RESERVE_SMS=$r timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2964$r \
   bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline --flow-epochs 0 --configs none --no-both > $O/bench_n2_res$r.json 2> $O/bench_n2_res$r.err
done
python - <<'PY'
import json
for r in (0, 1, 2):
    for l in open("gpurun_out/r2e/bench_n2_res%d.json" % r):
        if l.startswith("{"):
            d = json.loads(l)
            print("reserve", r, "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), "lat", round(d["latency"]["ms_per_epoch"], 4), "k_brute", round(d["roofline"]["kernel_ms"], 4))
PY
