#!/bin/bash
# round 2, call F (2 GPUs): everything after the fused pre-pass, fast geometry and CUDA-graph submit
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2f; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -25 $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; tail -3 $O/bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29652 \
   bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --flow-epochs 0 > $O/bench_n2.json 2> $O/bench_n2.err; tail -3 $O/bench_n2.err
python - <<'PY'
import json
for f in ("bench_n1", "bench_n2"):
    try:
        for l in open("gpurun_out/r2f/%s.json" % f):
            if l.startswith("{"):
                d = json.loads(l)
                print(f, "ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], "lat", d["latency"]["ms_per_epoch"], d["latency"]["stage_ms"], "launches", d["gpu_launches_per_epoch"])
                print("  roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], "other", d.get("other_path"))
                print("  configs", {k: (v.get("ms_per_step"), (v.get("roofline") or {}).get("frac"), v.get("error")) for k, v in (d.get("configs") or {}).items()})
                print("  flow", d.get("flow"))
    except Exception as e:
        print(f, "ERR", e)
PY
