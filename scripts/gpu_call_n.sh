#!/bin/bash
# compute-sanitizer over the smoke epoch (lookup + brute) and a split-tail case
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/n
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/n/memcheck.log 2>&1
grep -v "^\[" gpurun_out/n/memcheck.log | tail -8
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/n/racecheck.log 2>&1
tail -6 gpurun_out/n/racecheck.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_properties.py -m gpu -x -q -k "split_tail_slots and 4700 or minimum_sizes or ten_mega" > gpurun_out/n/memcheck_tail.log 2>&1
tail -6 gpurun_out/n/memcheck_tail.log
