#!/bin/bash
# GPU call B: locate the reference's illegal access, first bench line, ncu evidence.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/prof
python - <<'PY'
import sys; sys.path.insert(0, '.')
import dpe_pkg
synth = dpe_pkg.submodule("synth")
sc = synth.Scenario()
grid, _ = synth.uniform_grid(9, (5.0, 5.0, 5.0, 6.0))
print(sc.write_files("/tmp/refrun", 6, grid=grid, handoff_block=1))
PY
R=navlab-dpe-sdr_b200/data/brdc_toe417600.18n
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 oracle/_ref/ref_dpe /tmp/refrun/synthetic_l1ca_2500kHz.dat \
   /tmp/refrun/handoff_params_synth.csv $R /tmp/refrun/rngrid_synth.csv 9 5 1 /tmp/refrun/dump 32 2.5e6 1 \
   > gpurun_out/ref_sanitizer.log 2>&1
grep -v "^\[" gpurun_out/ref_sanitizer.log | head -60
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_demo.json 2> gpurun_out/bench_demo.err
cat gpurun_out/bench_demo.json; tail -5 gpurun_out/bench_demo.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
   --log-file gpurun_out/prof/launches_demo.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-both \
   > gpurun_out/prof/launches_demo.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_brute -s 1 -c 1 \
   -o gpurun_out/prof/k_brute_demo python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-both --workload c3 \
   > gpurun_out/prof/k_brute_ncu.out 2>&1
tail -3 gpurun_out/prof/k_brute_ncu.out
ls -la gpurun_out/prof
