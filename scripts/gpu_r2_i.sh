#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2i; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_golden_ref.py -m gpu -q -x > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 200 python bench.py --path lookup --steps 50 --warmup 5 --configs none --no-both --no-cpu-baseline --flow-epochs 0 > $O/bench_lookup.json 2> $O/bench_lookup.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_prep_corr|k_score_lookup" -s 4 -c 2 -o $O/lookup_kernels -f \
   python bench.py --path lookup --steps 3 --warmup 1 --depth 1 --configs none --no-both --no-cpu-baseline --flow-epochs 0 > $O/ncu_full.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/r2i/bench_lookup.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "lat", d["latency"]["ms_per_epoch"], d["latency"]["stage_ms"])
PY
