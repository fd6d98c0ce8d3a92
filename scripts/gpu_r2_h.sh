#!/bin/bash
# round 2, call H (1 GPU): 4-candidates-per-thread lookup kernel, fence fix
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2h; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; tail -8 $O/pytest_gpu.log
timeout 200 python bench.py --path lookup --steps 50 --warmup 5 --configs none --no-both --no-cpu-baseline --flow-epochs 0 > $O/bench_lookup.json 2> $O/bench_lookup.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_lookup.csv \
   python bench.py --path lookup --steps 3 --warmup 1 --depth 1 --configs none --no-both --no-cpu-baseline --flow-epochs 0 > $O/ncu_lookup.log 2>&1
python - <<'PY'
import json, csv
for f in ("bench_lookup",):
    for l in open("gpurun_out/r2h/%s.json" % f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, "ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "lat", d["latency"]["ms_per_epoch"], d["latency"]["stage_ms"], "launches", d["gpu_launches_per_epoch"], d["roofline"]["frac"])
rows = [r for r in csv.reader(l for l in open("gpurun_out/r2h/launches_lookup.csv") if l.startswith('"'))]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
for r in rows[1:][-12:-5]:
    print(r[ki][:60], r[vi])
PY
