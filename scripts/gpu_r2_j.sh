#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2j; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden_ref.py -m gpu -q -x > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
for nc in 0 3 4 6; do
DPE_LK_CAND=$nc timeout 200 python bench.py --path lookup --steps 50 --warmup 5 --configs none --no-both --no-cpu-baseline --flow-epochs 0 > $O/bench_lookup_nc$nc.json 2> $O/bench_lookup_nc$nc.err
done
python - <<'PY'
import json
for nc in (0, 3, 4, 6):
  for l in open("gpurun_out/r2j/bench_lookup_nc%d.json" % nc):
    if l.startswith("{"):
        d = json.loads(l)
        print("nc", nc, "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "lat", round(d["latency"]["ms_per_epoch"],4), d["latency"]["stage_ms"]["prepare"], d["latency"]["stage_ms"]["lookup"])
PY
