#!/bin/bash
# round 2, call AD (1 GPU): racecheck over the kernels that do not use the TMA ring (k_brute floods racecheck's hazard limit
# with its known false positive): fused pre-pass, lookup scoring with the folded estimate, carrier spectrum, both velocity kernels
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2ad; mkdir -p $O
NV_COMPUTE_SANITIZER_MAX_RACECHECK_HAZARDS=200 timeout 150 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
   -k "(velocity and prns0) or (fold_estimate and LOOKUP) or test_weighted_estimate or lookup_scores" > $O/racecheck_tests.log 2>&1
grep -v "^\[" $O/racecheck_tests.log | tail -30
