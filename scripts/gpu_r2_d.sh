#!/bin/bash
# round 2, call D (2 GPUs): long-run / bridge tests, NCCL in-library tests, 2-GPU flow, 2-GPU bench
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2d; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m pytest tests/test_longrun.py tests/test_bridge.py -m gpu -q -s > $O/pytest_long.log 2>&1; grep -v "^\[" $O/pytest_long.log | tail -25
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_host_flow.py -m gpu -q > $O/pytest_dist.log 2>&1; tail -25 $O/pytest_dist.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29632 \
   bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --flow-epochs 0 > $O/bench_n2.json 2> $O/bench_n2.err; tail -5 $O/bench_n2.err
python - <<'PY'
import json
for l in open("gpurun_out/r2d/bench_n2.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("N=2 ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], "lat", d["latency"])
        print("  roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["roofline"].get("kernel_share_pipelined"))
        print("  other", d.get("other_path"))
        print("  configs", {k: (v.get("ms_per_step"), (v.get("roofline") or {}).get("frac"), v.get("error")) for k, v in (d.get("configs") or {}).items()})
PY
