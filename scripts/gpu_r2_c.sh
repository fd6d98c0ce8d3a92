#!/bin/bash
# round 2, call C (1 GPU): long-run golden (fast moving receiver), long-run / bridge / a13 tests
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2c; mkdir -p $O
timeout 600 python oracle/make_golden_ref.py --longrun 300 --out $O/golden --work /tmp/refl > $O/golden_longrun.log 2>&1; tail -4 $O/golden_longrun.log
cp $O/golden/ref_longrun_moving.npz tests/golden/
timeout 900 python -m pytest tests/test_longrun.py tests/test_bridge.py tests/test_golden_ref.py -m gpu -q -s > $O/pytest.log 2>&1; grep -v "^\[" $O/pytest.log | tail -40
