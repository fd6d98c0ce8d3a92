#!/bin/bash
# round 2, call U (1 GPU): lookup-path rework, second pass (k_score_vel 6 candidates per thread, DC sums inside k_prep_corr)
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2u; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
python scripts/lookup_probe.py demo > $O/probe_default.log 2>&1; tail -1 $O/probe_default.log
DPE_VEL_FORK=0 python scripts/lookup_probe.py demo > $O/probe_nofork.log 2>&1; tail -1 $O/probe_nofork.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_lookup_vel.csv \
   python scripts/lookup_probe.py demo > $O/ncu_a.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2u/launches_lookup_vel.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); ig=hdr.index('Grid Size')
for r in rows[-10:]: print(r[ik][:44], r[iv], r[ig])
PY
timeout 200 python bench.py --path lookup --steps 50 --warmup 5 --no-cpu-baseline --flow-epochs 100 --configs none --no-both --no-vel-brute > $O/bench_lookup.json 2> $O/bench_lookup.err
python - <<'PY'
import json
for l in open('gpurun_out/r2u/bench_lookup.json'):
    if l.startswith('{'):
        d=json.loads(l); print('lookup ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'lat', d['latency']); print('flow', json.dumps(d.get('flow'))[:1500])
PY
