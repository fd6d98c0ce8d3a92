#!/bin/bash
# First GPU call of round 1: parity tests, reference golden vectors, micro-benchmarks.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 600 python oracle/make_golden_ref.py --out gpurun_out/golden > gpurun_out/golden.log 2>&1
tail -30 gpurun_out/golden.log
timeout 120 python - <<'PY' 2>&1 | tee gpurun_out/microbench.log
import dpe_pkg
capi = dpe_pkg.submodule("capi")
print("fp32 FFMA  TFLOP/s", capi.microbench_fp32(0, False))
print("fp32 FFMA2 TFLOP/s", capi.microbench_fp32(0, True))
print("hbm copy GB/s", capi.microbench_hbm(0, 1 << 30))
PY
