#!/bin/bash
# round 2, call R (1 GPU): lookup-path rework -- moment-based carrier spectrum, velocity branch on its own stream, 64-thread
# k_score_lookup CTAs, parallel tail of k_prep_corr, estimate folded into the scoring kernel: tests + A/B probes
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2r; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
python scripts/lookup_probe.py demo > $O/probe_default.log 2>&1; tail -1 $O/probe_default.log
DPE_VEL_FORK=0 python scripts/lookup_probe.py demo > $O/probe_nofork.log 2>&1; tail -1 $O/probe_nofork.log
DPE_VEL_FORK=0 DPE_CARR_DIRECT=1 python scripts/lookup_probe.py demo > $O/probe_nofork_direct.log 2>&1; tail -1 $O/probe_nofork_direct.log
for v in "3 128" "6 128" "6 64" "4 128"; do set -- $v
DPE_VEL_FORK=0 DPE_LK_CAND=$1 DPE_LK_BLOCK=$2 python scripts/lookup_probe.py demo 2>&1 | tail -1
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_lookup_vel.csv \
   python scripts/lookup_probe.py demo > $O/ncu_a.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2r/launches_lookup_vel.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); ig=hdr.index('Grid Size')
for r in rows[-14:]: print(r[ik][:44], r[iv], r[ig])
PY
