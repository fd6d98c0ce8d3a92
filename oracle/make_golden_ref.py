#!/usr/bin/env python
"""Generate golden vectors from the UNMODIFIED reference (TEST INFRASTRUCTURE).

Runs on a GPU box: writes a synthetic L1 C/A capture + handoff CSV + grid CSV in
the reference's own file formats (synth.Scenario.write_files), runs the
reference's DPEFlow through oracle/_ref/ref_dpe (built by oracle/Makefile from
/root/reference/cudarecv, sm_100a), and packs what the reference consumed and
produced per epoch into one .npz:

  inputs  (cuChanMgr / cuEKF outputs before the epoch): prn, rc_start, ri_start,
          rc_end, fc, fi, cp_start, cp_end, cp_ref, cp_ref_tow, sat_states [C*T][8],
          enu2ecef, x_kk1, rx_time, iq (the int16 block SampleBlock handed over)
  outputs (after BatchCorrManifold::Update): code_scores_win [C][2W+2] complex,
          pos_scores [G], zval [8], and x_k1k1 after the epoch.

The .npz is committed under tests/golden/ (this script is how it was made); the
CPU tests pin oracle/dpe_oracle.py against it and the GPU tests pin the CUDA path.

usage: python oracle/make_golden_ref.py --out gpurun_out/golden [--epochs 3] [--n 9]
"""
import argparse
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dpe_pkg  # noqa: E402

synth = dpe_pkg.submodule("synth")

I32 = ("cp_ref", "cp_start", "cp_end", "cp_ref_tow")
F64 = ("rx_time", "tx_time", "rc_start", "ri_start", "rc_end", "ri_end", "fc", "fi", "sat_states", "sat_raw",
       "enu2ecef", "x_kk1", "x_k1k1", "code_scores_win", "carr_scores_win", "pos_scores", "zval", "time_grid")


def golden_files(work, epochs, n, offset):
    """Capture + handoff + RINEX + n^4 grid CSV of the golden runs (also used by tests/test_bridge.py)."""
    sc = synth.Scenario()
    grid, _ = synth.uniform_grid(n, (5.0, 5.0, 5.0, 6.0))
    # StartByte 0 is rejected by the reference (sampleblock.cu:123-128) -> hand off at block 1.  The file
    # must outlast the reader's 32-block read-ahead: at EOF the reference frees its buffers while the
    # flow may still be using the last one (sampleblock.cu:449-462).
    files = sc.write_files(work, epochs + 40, grid=grid, handoff_block=1)
    # hand the reference a state that is `offset` away from the truth so the arg-max is not the centre
    lines = open(files["handoff"]).read().splitlines()
    for i, l in enumerate(lines):
        if l.startswith("X_ECEF,"):
            v = [float(x) for x in l.split(",")[1:]]
            for k in range(4):
                v[k] += float(offset[k])
            lines[i] = "X_ECEF," + ",".join(repr(float(x)) for x in v)
    open(files["handoff"], "w").write("\n".join(lines) + "\n")
    return sc, grid, files


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/golden")
    ap.add_argument("--work", default="/tmp/refrun")
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--n", type=int, default=9)
    ap.add_argument("--W", type=int, default=32)
    ap.add_argument("--ekf", action="store_true", help="enable the reference's 8-state KF (cuEKF EnableEKF=true); "
                    "writes only the per-epoch states (ref_ekf_n<N>.npz)")
    ap.add_argument("--gen-grid", type=int, default=-1, help="let the reference GENERATE its position grid "
                    "(LoadPosGrid=false): ManifoldGridTypes value, 0 Uniform / 2 ArthurBasis; writes ref_grid_t<T>_n<N>.npz")
    ap.add_argument("--lpower", type=int, default=1, help="BatchCorrManifold LPower; != 1 writes ref_bcm_L<L>_n<N>.npz "
                    "(the BCM inputs and outputs only)")
    ap.add_argument("--weighted", action="store_true", help="run oracle/_ref/ref_dpe_weighted: after every epoch the "
                    "reference's DORMANT BCM_PosMeasReduction + BCM_ReduceAndPosMeas are launched on the module's own "
                    "buffers; writes ref_weighted_n<N>.npz (BCM inputs, CodeScores window, zval_weighted, block partials)")
    ap.add_argument("--longrun", type=int, default=0, help="closed-loop run of this many epochs on a MOVING receiver "
                    "(25^4 spread position grid from CSV + 25^4 velocity grid, 0.5 m/s): writes ref_longrun_moving.npz "
                    "(per-epoch fixes, channel parameters and CodeScores windows; the capture is regenerated from the "
                    "seed by the test)")
    ap.add_argument("--spacing", type=float, default=2.0, help="GridDimSpacing for --gen-grid (all 8 dimensions)")
    ap.add_argument("--offset", type=float, nargs=4, default=[7.0, -4.0, 3.0, 8.0],
                    help="ECEF x,y,z and clock (m) offset of the handed-off state from the truth")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_dpe_weighted" if a.weighted else "ref_dpe")
    if not os.path.exists(exe):
        sys.exit("%s missing: run `make -C oracle` where /root/reference exists" % exe)
    if a.longrun:
        return longrun(a, exe)

    sc, grid, files = golden_files(a.work, a.epochs, a.n, a.offset)

    dump = os.path.join(a.work, "dump")
    cmd = [exe, files["dat"], files["handoff"], files["rinex"], files["grid"] if a.gen_grid < 0 else "none", str(a.n), "5",
           str(a.epochs), dump, str(a.W), repr(sc.cfg.fs), "1", "1" if a.ekf else "0"]
    if a.gen_grid >= 0 or a.lpower != 1:
        cmd += [str(max(a.gen_grid, 0)), repr(a.spacing if a.gen_grid >= 0 else 1.0), str(a.lpower)]
    print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    open(os.path.join(a.out, "ref_run.log"), "w").write(r.stdout)
    print(r.stdout[-3000:])
    if r.returncode != 0:
        sys.exit("reference run failed (%d)" % r.returncode)

    meta = dict(l.split() for l in open(os.path.join(dump, "meta.txt")))
    C, CT = int(meta["C"]), int(meta["CT"])
    pack = dict(first_block=1, Wd=64, vel_dim=5, C=C, T=CT // C, S=sc.S, fs=sc.cfg.fs, W=a.W, n=a.n, epochs=a.epochs, grid=grid,
                offset=np.array(a.offset), truth0=sc.rx_state(sc.cfg.rx_time0 + sc.cfg.T))
    for e in range(a.epochs):
        def rd(name, dt):
            return np.fromfile(os.path.join(dump, "e%03d_%s.bin" % (e, name)), dtype=dt)
        for k in F64:
            pack["e%d_%s" % (e, k)] = rd(k, np.float64)
        for k in I32:
            pack["e%d_%s" % (e, k)] = rd(k, np.int32)
        pack["e%d_prn" % e] = rd("prn", np.uint8)
        pack["n_fft"] = int(rd("n_fft", np.int32)[0])
        pack["e%d_iq" % e] = rd("iq", np.int16)
    if a.ekf:
        keep = ("x_kk1", "x_k1k1", "zval", "rx_time")
        import re
        pack = {k: v for k, v in pack.items()
                if not re.match(r"^e\d+_", k) or re.sub(r"^e\d+_", "", k) in keep}
        pack["ekf"] = 1
    name = ("ref_ekf_n%d.npz" if a.ekf else "ref_epochs_n%d.npz") % a.n
    if a.gen_grid >= 0:                                    # keep what pins the generator: axis values, scores, fixes
        import re
        keep = ("x_kk1", "x_k1k1", "zval", "rx_time", "time_grid", "pos_scores")
        pack = {k: v for k, v in pack.items()
                if not re.match(r"^e\d+_", k) or re.sub(r"^e\d+_", "", k) in keep}
        del pack["grid"]
        pack["grid_type"], pack["spacing"] = a.gen_grid, a.spacing
        name = "ref_grid_t%d_n%d.npz" % (a.gen_grid, a.n)
    if a.weighted:                                         # BatchCorrManifold in and out + the dormant estimator's results
        import re
        drop = ("iq", "sat_raw", "ri_end", "pos_scores")
        pack = {k: v for k, v in pack.items() if not re.match(r"^e\d+_", k) or re.sub(r"^e\d+_", "", k) not in drop}
        for e in range(a.epochs):
            pack["e%d_zval_weighted" % e] = np.fromfile(os.path.join(dump, "e%03d_zval_weighted.bin" % e), dtype=np.float64)
            pack["e%d_weighted_parts" % e] = np.fromfile(os.path.join(dump, "e%03d_weighted_parts.bin" % e),
                                                         dtype=np.float64).reshape(8, 5)
            # the velocity twins (BCM_VelMeasReduction + BCM_ReduceAndVelMeas) on the module's CarrScores
            pack["e%d_zval_weighted_vel" % e] = np.fromfile(os.path.join(dump, "e%03d_zval_weighted_vel.bin" % e),
                                                            dtype=np.float64)
            pack["e%d_weighted_vel_parts" % e] = np.fromfile(os.path.join(dump, "e%03d_weighted_vel_parts.bin" % e),
                                                             dtype=np.float64).reshape(8, 5)
        name = "ref_weighted_n%d.npz" % a.n
    if a.lpower != 1:                                      # BatchCorrManifold in and out: enough to pin sum |v|^L
        import re
        drop = ("iq", "carr_scores_win", "sat_raw", "tx_time", "ri_end")
        pack = {k: v for k, v in pack.items() if not re.match(r"^e\d+_", k) or re.sub(r"^e\d+_", "", k) not in drop}
        pack["lpower"] = a.lpower
        name = "ref_bcm_L%d_n%d.npz" % (a.lpower, a.n)
    path = os.path.join(a.out, name)
    np.savez_compressed(path, **pack)
    print("wrote", path, os.path.getsize(path), "bytes")
    if os.path.exists(os.path.join(dump, "XFile.csv")):
        open(os.path.join(a.out, "ref_XFile.csv"), "w").write(open(os.path.join(dump, "XFile.csv")).read())


LONGRUN = dict(truth_vel_enu=(12.0, 8.0, 1.0), offset=(3.3, -2.1, 1.7, 4.4), vel_dim=25, vel_spacing=0.5, W=4)


def longrun_scenario():
    """The moving-receiver scenario of the long closed-loop run (tests regenerate the same capture)."""
    return synth.Scenario(synth.ScenarioConfig(truth_vel_enu=LONGRUN["truth_vel_enu"]))


def longrun_files(work, epochs):
    """Capture, handoff (block 1, state `offset` away from the truth), RINEX and 25^4 spread grid CSV."""
    sc = longrun_scenario()
    files = sc.write_files(work, epochs + 40, grid=synth.spread_grid(), handoff_block=1)
    lines = open(files["handoff"]).read().splitlines()
    for i, l in enumerate(lines):
        if l.startswith("X_ECEF,"):
            v = [float(x) for x in l.split(",")[1:]]
            for k in range(4):
                v[k] += float(LONGRUN["offset"][k])
            lines[i] = "X_ECEF," + ",".join(repr(float(x)) for x in v)
    open(files["handoff"], "w").write("\n".join(lines) + "\n")
    return sc, files


def longrun(a, exe):
    sc, files = longrun_files(a.work, a.longrun)
    dump = os.path.join(a.work, "dump_long")
    W = LONGRUN["W"]
    cmd = [exe, files["dat"], files["handoff"], files["rinex"], files["grid"], "25", str(LONGRUN["vel_dim"]),
           str(a.longrun), dump, str(W), repr(sc.cfg.fs), "2", "0", "0", repr(LONGRUN["vel_spacing"]), "1"]
    print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1200)
    open(os.path.join(a.out, "ref_longrun.log"), "w").write(r.stdout)
    print(r.stdout[-2000:])
    if r.returncode != 0:
        sys.exit("reference run failed (%d)" % r.returncode)
    n = a.longrun
    names_f = ("rx_time", "tx_time", "rc_start", "ri_start", "rc_end", "fc", "fi", "x_kk1", "x_k1k1", "zval",
               "code_scores_win", "sat_raw", "enu2ecef")
    names_i = ("cp_ref", "cp_start", "cp_end", "cp_ref_tow")
    pack = dict(epochs=n, W=W, S=sc.S, fs=sc.cfg.fs, first_block=1, truth_vel_enu=np.array(LONGRUN["truth_vel_enu"]),
                offset=np.array(LONGRUN["offset"]), vel_dim=LONGRUN["vel_dim"], vel_spacing=LONGRUN["vel_spacing"],
                truth=np.stack([sc.rx_state(sc.cfg.rx_time0 + (e + 2) * sc.cfg.T) for e in range(n)]))
    for k in names_f:
        pack[k] = np.stack([np.fromfile(os.path.join(dump, "e%03d_%s.bin" % (e, k)), dtype=np.float64) for e in range(n)])
    for k in names_i:
        pack[k] = np.stack([np.fromfile(os.path.join(dump, "e%03d_%s.bin" % (e, k)), dtype=np.int32) for e in range(n)])
    pack["prn"] = np.fromfile(os.path.join(dump, "e000_prn.bin"), dtype=np.uint8)
    xf = os.path.join(dump, "XFile.csv")
    if os.path.exists(xf):
        txt = open(xf).read()
        open(os.path.join(a.out, "ref_longrun_XFile.csv"), "w").write(txt)
        rows = [[float(v) for v in l.replace(" ", "").strip(",").split(",")] for l in txt.splitlines() if l.strip()]
        pack["xfile"] = np.array([r_ for r_ in rows if len(r_) == 8])
    path = os.path.join(a.out, "ref_longrun_moving.npz")
    np.savez_compressed(path, **pack)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
