#!/bin/bash
# round 2, call AA (1 GPU): the reference's dormant velocity reduction kernels executed (ref_dpe_weighted) -> golden regenerated;
# weighted velocity estimate of the CUDA path against the oracle
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2aa; mkdir -p $O/golden
timeout 300 python oracle/make_golden_ref.py --weighted --out $O/golden --work /tmp/refw > $O/golden_weighted.log 2>&1; tail -4 $O/golden_weighted.log
cp $O/golden/ref_weighted_n9.npz tests/golden/ref_weighted_n9.npz
timeout 300 python -m pytest tests/test_golden_ref.py tests/test_gpu_parity.py -q -k "weighted or velocity" 2>&1 | tail -8
