#!/bin/bash
# round 2, call AC (1 GPU): compute-sanitizer over the round-2 kernels (fused pre-pass, folded estimate, CUDA-graph submit,
# 3-launch pair sort, moment-based carrier spectrum, both velocity formulations) -- memcheck and racecheck, bounded
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2ac; mkdir -p $O
timeout 100 compute-sanitizer --tool memcheck --print-limit 10 python -c "import __graft_entry__ as g; g.smoke()" > $O/memcheck_smoke.log 2>&1
grep -v "^\[" $O/memcheck_smoke.log | tail -6
timeout 110 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
   -k "(velocity and prns0) or fold_estimate or test_weighted_estimate or ragged_groups" > $O/memcheck_tests.log 2>&1
tail -6 $O/memcheck_tests.log
timeout 80 compute-sanitizer --tool racecheck --print-limit 10 python -c "import __graft_entry__ as g; g.smoke()" > $O/racecheck_smoke.log 2>&1
tail -5 $O/racecheck_smoke.log
