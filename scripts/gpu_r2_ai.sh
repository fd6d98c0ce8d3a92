#!/bin/bash
# round 2, call AI (1 GPU): final tree after the lookup-path rework (integer-atomic chunk totals, no k_carr_finalize, batched
# block-partial reduction, programmatic dependent launch of the scoring kernels): full GPU tests, smoke, default bench line, launch lists
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2ai; mkdir -p $O
timeout 120 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -4 $O/smoke.log
timeout 200 python bench.py > $O/bench_demo_n1.json 2> $O/bench_demo_n1.err; tail -2 $O/bench_demo_n1.err
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file $O/launches_demo_lookup_steps2.csv \
   python bench.py --path lookup --steps 2 --warmup 1 --depth 1 --configs none --no-both --no-cpu-baseline --flow-epochs 0 --no-vel-brute --no-ncu-traffic > $O/ncu_b.log 2>&1
timeout 60 python scripts/lookup_wall_probe.py demo 300 > $O/wall_probe.log 2>&1; cat $O/wall_probe.log
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_demo_steps2.csv \
   python bench.py --steps 2 --warmup 1 --configs none --no-both --no-cpu-baseline --flow-epochs 0 --no-vel-brute --no-ncu-traffic > $O/ncu_a.log 2>&1
