#!/bin/bash
# round 2, call K (1 GPU): the evidence run -- full default bench, launch lists, ncu --set full of k_brute and the side kernels
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2k; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_demo_n1.json 2> $O/bench_demo_n1.err; tail -2 $O/bench_demo_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_demo_steps2.csv \
   python bench.py --steps 2 --warmup 1 --configs none --no-both --no-cpu-baseline --flow-epochs 0 > $O/ncu_a.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file $O/launches_demo_lookup_steps2.csv \
   python bench.py --path lookup --steps 2 --warmup 1 --depth 1 --configs none --no-both --no-cpu-baseline --flow-epochs 0 > $O/ncu_b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_brute" -s 2 -c 1 -o $O/k_brute_demo -f \
   python bench.py --steps 2 --warmup 1 --depth 1 --configs none --no-both --no-cpu-baseline --flow-epochs 0 > $O/ncu_c.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"k_prep_corr|k_pair_bins|k_block_scan|k_scatter|k_sample_planes|k_replica_rd|k_score_pairs|k_finalize" -s 16 -c 8 -o $O/side_kernels -f \
   python bench.py --steps 2 --warmup 1 --depth 1 --configs none --no-both --no-cpu-baseline --flow-epochs 0 > $O/ncu_d.log 2>&1
python scripts/ncu_summary.py $O/k_brute_demo.ncu-rep > $O/k_brute_demo_ncu_summary.txt
python scripts/ncu_summary.py $O/side_kernels.ncu-rep > $O/side_kernels_ncu_summary.txt
head -40 $O/k_brute_demo_ncu_summary.txt
