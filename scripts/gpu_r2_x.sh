#!/bin/bash
# round 2, call X (8 GPUs): scaling 1 / 2 / 4 / 8 of the final tree, sub-records (c3, c4, c5) at 8
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2x; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --flow-epochs 0 --configs none --no-both --no-vel-brute > $O/bench_n1.json 2> $O/bench_n1.err
for n in 2 4; do
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2969$n \
   bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline --flow-epochs 0 --configs none --no-both > $O/bench_n$n.json 2> $O/bench_n$n.err
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29698 \
   bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --flow-epochs 0 > $O/bench_n8.json 2> $O/bench_n8.err; tail -3 $O/bench_n8.err
python - <<'PY'
import json
for f in ("bench_n1", "bench_n2", "bench_n4", "bench_n8"):
    try:
        for l in open("gpurun_out/r2x/%s.json" % f):
            if l.startswith("{"):
                d = json.loads(l)
                print(f, "ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], "lat", d["latency"]["ms_per_epoch"], "k_brute", d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"].get("kernel_share_pipelined"))
                for k, v in (d.get("configs") or {}).items():
                    print("   ", k, v.get("error") or (v["ms_per_step"], v["value"], v["roofline"]["frac"], v["roofline"]["kernel_ms"], (v.get("e2e") or {}).get("ms_per_step")))
                if d.get("other_path"): print("   other", d["other_path"]["ms_per_step"], d["other_path"]["latency_ms_per_epoch"])
    except Exception as e:
        print(f, "ERR", e)
PY
