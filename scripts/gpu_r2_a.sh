#!/bin/bash
# round 2, call A (1 GPU): GPU parity tests, reference goldens (weighted estimator, long moving run),
# bridge, first pipelined bench + launch list
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_bridge.py > $O/pytest_gpu.log 2>&1; tail -15 $O/pytest_gpu.log
timeout 300 python -m pytest tests/test_bridge.py -m gpu -q > $O/pytest_bridge.log 2>&1; tail -30 $O/pytest_bridge.log
timeout 300 python oracle/make_golden_ref.py --weighted --out $O/golden --work /tmp/refw > $O/golden_weighted.log 2>&1; tail -5 $O/golden_weighted.log
timeout 600 python oracle/make_golden_ref.py --longrun 300 --out $O/golden --work /tmp/refl > $O/golden_longrun.log 2>&1; tail -5 $O/golden_longrun.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; tail -3 $O/bench_n1.err
timeout 300 python bench.py --steps 20 --warmup 3 --depth 1 --configs none --no-both --no-cpu-baseline --flow-epochs 0 > $O/bench_n1_depth1.json 2> $O/bench_n1_depth1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_demo_steps2.csv \
   python bench.py --steps 2 --warmup 1 --configs none --no-both --no-cpu-baseline --flow-epochs 0 > $O/ncu_bench.log 2>&1
python - <<'PY'
import json
for f in ("bench_n1", "bench_n1_depth1"):
    try:
        for l in open("gpurun_out/r2a/%s.json" % f):
            if l.startswith("{"):
                d = json.loads(l)
                print(f, "ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], "lat", d["latency"]["ms_per_epoch"], d["latency"]["stage_ms"])
                print("  roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], "other", d.get("other_path"))
                print("  configs", {k: (v.get("ms_per_step"), (v.get("roofline") or {}).get("frac")) for k, v in (d.get("configs") or {}).items()})
                print("  flow", d.get("flow"), "refgpu", d.get("reference_gpu"), "like", d.get("like_for_like"))
                print("  side", d.get("side_kernels"))
    except Exception as e:
        print(f, "ERR", e)
PY
