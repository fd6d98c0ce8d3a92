#!/bin/bash
# round 2, call Q (1 GPU): brute-force velocity manifold -- parity tests, probe, bench leg, ncu --set full of k_brute_vel
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2q; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "velocity" > $O/pytest_vel.log 2>&1; tail -15 $O/pytest_vel.log
timeout 120 python scripts/vel_brute_probe.py 25 3 > $O/vel_probe.log 2>&1; tail -5 $O/vel_probe.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --flow-epochs 0 --configs none --no-both > $O/bench.json 2> $O/bench.err; tail -2 $O/bench.err
python -c "
import json
for l in open('$O/bench.json'):
    if l.startswith('{'):
        d = json.loads(l); print(json.dumps(d.get('velocity_brute'), indent=1)); print(d['ms_per_step'], d['roofline']['frac'])
"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_brute_vel" -s 1 -c 1 -o $O/k_brute_vel -f \
   python scripts/vel_brute_probe.py 25 1 > $O/ncu_vel.log 2>&1; tail -3 $O/ncu_vel.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_vel|k_brute_vel|k_score_vpairs|k_block_scan|k_scatter|k_carr|k_score_vel" --csv --log-file $O/launches_vel.csv \
   python scripts/vel_brute_probe.py 25 1 > $O/ncu_vel2.log 2>&1
python scripts/ncu_summary.py $O/k_brute_vel.ncu-rep > $O/k_brute_vel_ncu_summary.txt; head -60 $O/k_brute_vel_ncu_summary.txt
