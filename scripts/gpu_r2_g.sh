#!/bin/bash
# round 2, call G (2 GPUs): NCCL tests after the graph fix; lookup-path per-kernel times (ncu) and a full capture
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2g; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_dist.py tests/test_host_flow.py -m gpu -q -x > $O/pytest_dist.log 2>&1; tail -8 $O/pytest_dist.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29662 \
   bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --flow-epochs 0 --configs none > $O/bench_n2.json 2> $O/bench_n2.err; tail -3 $O/bench_n2.err
timeout 200 python bench.py --path lookup --steps 50 --warmup 5 --configs none --no-both --no-cpu-baseline --flow-epochs 0 > $O/bench_lookup.json 2> $O/bench_lookup.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_lookup.csv \
   python bench.py --path lookup --steps 3 --warmup 1 --depth 1 --configs none --no-both --no-cpu-baseline --flow-epochs 0 > $O/ncu_lookup.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_prep_corr|k_score_lookup" -s 4 -c 2 -o $O/lookup_kernels -f \
   python bench.py --path lookup --steps 3 --warmup 1 --depth 1 --configs none --no-both --no-cpu-baseline --flow-epochs 0 > $O/ncu_full.log 2>&1
python - <<'PY'
import json, csv
for f in ("bench_n2", "bench_lookup"):
    try:
        for l in open("gpurun_out/r2g/%s.json" % f):
            if l.startswith("{"):
                d = json.loads(l)
                print(f, "ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "lat", d["latency"]["ms_per_epoch"], d["latency"]["stage_ms"], "launches", d["gpu_launches_per_epoch"], d["roofline"]["frac"])
    except Exception as e:
        print(f, "ERR", e)
rows = [r for r in csv.reader(l for l in open("gpurun_out/r2g/launches_lookup.csv") if l.startswith('"'))]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
for r in rows[1:][-16:]:
    print(r[ki][:60], r[vi])
PY
