#!/bin/bash
# round 2, call L (8 GPUs): in-library NCCL tests on 4 ranks, demo at 8 and 4 GPUs with the c3 / c4 / c5 sub-records
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2l; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 300 python -m pytest tests/test_gpu_dist.py tests/test_host_flow.py -m gpu -q -k "nccl or sharded or another_device" > $O/pytest_dist.log 2>&1; tail -4 $O/pytest_dist.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29688 \
   bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --flow-epochs 0 > $O/bench_n8.json 2> $O/bench_n8.err; tail -3 $O/bench_n8.err
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29684 \
   bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline --flow-epochs 0 --configs none --no-both > $O/bench_n4.json 2> $O/bench_n4.err; tail -3 $O/bench_n4.err
timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --flow-epochs 0 --configs none --no-both > $O/bench_n1.json 2> $O/bench_n1.err
python - <<'PY'
import json
for f in ("bench_n1", "bench_n4", "bench_n8"):
    try:
        for l in open("gpurun_out/r2l/%s.json" % f):
            if l.startswith("{"):
                d = json.loads(l)
                print(f, "ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], "lat", d["latency"]["ms_per_epoch"], "k_brute", d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"].get("kernel_share_pipelined"))
                for k, v in (d.get("configs") or {}).items():
                    print("   ", k, v.get("error") or (v["ms_per_step"], v["value"], v["roofline"]["frac"], v["roofline"]["kernel_ms"], (v.get("e2e") or {}).get("ms_per_step")))
    except Exception as e:
        print(f, "ERR", e)
PY
