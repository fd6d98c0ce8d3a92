"""navlab-dpe-sdr_b200 -- B200-native DPE batch-correlation-manifold hot path.

The directory name carries a hyphen (it mirrors the reference's repository
name), so it is imported under the module name ``navlab_dpe_sdr_b200`` through
``dpe_pkg.load()`` at the repo root.

Contents (only what the hot path needs):
  csrc/    hand-written sm_100a CUDA kernels + the C-ABI (libdpe_b200.so)
  host/    C++ mirror of the reference's dsp module/flow interface over the C-ABI
  capi.py  ctypes binding of include/dpe_b200.h (what the tests and bench call)
  synth.py synthetic GPS L1 C/A scenarios (the reference's dataset is missing)
"""
__all__ = ["capi", "synth", "gpsmath"]
