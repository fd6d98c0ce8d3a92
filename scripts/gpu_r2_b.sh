#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2b; mkdir -p $O
timeout 300 python scripts/dbg_bridge.py ref_dpe > $O/dbg_ref.log 2>&1; tail -70 $O/dbg_ref.log
timeout 300 python scripts/dbg_bridge.py ref_dpe_bridge > $O/dbg_bridge.log 2>&1; tail -90 $O/dbg_bridge.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_bridge.py > $O/pytest_gpu.log 2>&1; tail -15 $O/pytest_gpu.log
