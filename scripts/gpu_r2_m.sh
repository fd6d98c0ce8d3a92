#!/bin/bash
# round 2, call M (2 GPUs): padding groups skip their FFMA work, NCCL collectives capped at one CTA
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2m; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_dist.py -m gpu -q -x > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --flow-epochs 0 --configs none --no-both > $O/bench_n1.json 2> $O/bench_n1.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29672 \
   bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --flow-epochs 0 --configs c3 --no-both > $O/bench_n2.json 2> $O/bench_n2.err; tail -3 $O/bench_n2.err
python - <<'PY'
import json
for f in ("bench_n1", "bench_n2"):
    try:
        for l in open("gpurun_out/r2m/%s.json" % f):
            if l.startswith("{"):
                d = json.loads(l)
                print(f, "ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], "lat", d["latency"]["ms_per_epoch"], "k_brute", d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"].get("kernel_share_pipelined"))
                for k, v in (d.get("configs") or {}).items():
                    print("   ", k, v.get("error") or (v["ms_per_step"], v["value"], v["roofline"]["frac"], v["roofline"]["kernel_ms"], (v.get("e2e") or {}).get("ms_per_step")))
    except Exception as e:
        print(f, "ERR", e)
PY
