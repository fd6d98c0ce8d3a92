#!/bin/bash
# The reference's own kernels on B200 at the demo shape (25^4 spread grid from CSV), for the record.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/ref
python - <<'PY'
import sys; sys.path.insert(0, '.')
import dpe_pkg
synth = dpe_pkg.submodule("synth")
sc = synth.Scenario()
print(sc.write_files("/tmp/refrun", 110, grid=synth.spread_grid(), handoff_block=1))
PY
R=navlab-dpe-sdr_b200/data/brdc_toe417600.18n
for V in 5 25; do
  timeout 300 oracle/_ref/ref_dpe /tmp/refrun/synthetic_l1ca_2500kHz.dat /tmp/refrun/handoff_params_synth.csv $R \
     /tmp/refrun/rngrid_synth.csv 25 $V 60 /tmp/refrun/dump_v$V 32 2.5e6 0 > gpurun_out/ref/ref_timing_pos25_vel$V.log 2>&1
  grep "REF_\|error\|terminate\|what" gpurun_out/ref/ref_timing_pos25_vel$V.log | tail -5
done
timeout 300 compute-sanitizer --tool memcheck --print-limit 3 oracle/_ref/ref_dpe /tmp/refrun/synthetic_l1ca_2500kHz.dat \
   /tmp/refrun/handoff_params_synth.csv $R /tmp/refrun/rngrid_synth.csv 25 25 1 /tmp/refrun/dump_s 32 2.5e6 0 \
   > gpurun_out/ref/ref_sanitizer_vel25.log 2>&1
grep -v "^\[" gpurun_out/ref/ref_sanitizer_vel25.log | grep "=========" | head -20
