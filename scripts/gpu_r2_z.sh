#!/bin/bash
# round 2, call Z (1 GPU): narrow slots in k_brute (helper warps split the chunks of a padded slot, shares equal in cost) -- parity
# + A/B on the whole grid and on what rank 3 of 8 would hold.  The experiment is scripts/probe/k_brute_narrow_slots.patch (NOT in
# the tree: no gain); DPE_BRUTE_PAD_MODE only exists with that patch applied.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2z; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_dist.py -m gpu -q -x -k "brute or shard or split or ragged or random or maximum or c3 or c4 or c5 or side_kernels or two_contexts" > $O/pytest.log 2>&1; tail -4 $O/pytest.log
for pm in 1 2 1 2; do
DPE_BRUTE_PAD_MODE=$pm python scripts/brute_probe.py demo c3 2>&1 | grep -E "^demo|^c3" | sed "s/^/pad_mode $pm full: /"
done
for pm in 1 2 1 2; do
DPE_BRUTE_PAD_MODE=$pm DPE_BENCH_SHARD_OF=8 timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --flow-epochs 0 --configs none --no-both --no-vel-brute --no-ncu-traffic 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('pad_mode $pm shard 3/8: ms/step', round(d['ms_per_step'], 4), 'k_brute', round(d['roofline']['kernel_ms'], 4), 'lat', round(d['latency']['ms_per_epoch'], 4))"
done
