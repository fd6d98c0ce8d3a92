#!/bin/bash
# round 2, call Y (1 GPU): the evidence run of the final tree -- full GPU tests, smoke, default bench, reference arm, launch lists,
# ncu --set full of k_brute, the side kernels, the lookup-path kernels and the velocity kernels
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2y; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -5 $O/smoke.log
timeout 600 python bench.py > $O/bench_demo_n1.json 2> $O/bench_demo_n1.err; tail -2 $O/bench_demo_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_demo_steps2.csv \
   python bench.py --steps 2 --warmup 1 --configs none --no-both --no-cpu-baseline --flow-epochs 0 --no-vel-brute --no-ncu-traffic > $O/ncu_a.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file $O/launches_demo_lookup_steps2.csv \
   python bench.py --path lookup --steps 2 --warmup 1 --depth 1 --configs none --no-both --no-cpu-baseline --flow-epochs 0 --no-vel-brute --no-ncu-traffic > $O/ncu_b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_brute$|k_brute\(" -s 2 -c 1 -o $O/k_brute_demo -f \
   python bench.py --steps 2 --warmup 1 --depth 1 --configs none --no-both --no-cpu-baseline --flow-epochs 0 --no-vel-brute --no-ncu-traffic > $O/ncu_c.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"k_prep_corr|k_pair_bins|k_block_scan|k_scatter|k_sample_planes|k_replica_rd|k_score_pairs" -s 14 -c 7 -o $O/side_kernels -f \
   python bench.py --steps 2 --warmup 1 --depth 1 --configs none --no-both --no-cpu-baseline --flow-epochs 0 --no-vel-brute --no-ncu-traffic > $O/ncu_d.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"k_prep_corr|k_score_lookup|k_score_vel|k_carr_partial|k_carr_finalize" -s 10 -c 5 -o $O/lookup_kernels -f \
   python scripts/lookup_probe.py demo > $O/ncu_e.log 2>&1
for f in k_brute_demo side_kernels lookup_kernels; do python scripts/ncu_summary.py $O/$f.ncu-rep > $O/${f}_ncu_summary.txt; done
head -30 $O/k_brute_demo_ncu_summary.txt
python - <<'PY'
import json
for l in open('gpurun_out/r2y/bench_demo_n1.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'launches/epoch', d['gpu_launches_per_epoch'])
        print('other', json.dumps(d['other_path'])[:600])
        print('like', json.dumps(d['like_for_like'])[:600])
        print('flow', json.dumps(d['flow'])[:900])
        print('vel', json.dumps(d['velocity_brute'])[:600])
        print('cpu', json.dumps(d['cpu_baseline'])[:300])
        for k,v in d['configs'].items(): print(k, v.get('error') or (v['ms_per_step'], v['value'], v['roofline']['frac']))
PY
