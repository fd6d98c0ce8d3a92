"""debug: run the bridged reference on the golden files and diff every dumped array with the golden npz"""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import make_golden_ref as mg
exe = sys.argv[1] if len(sys.argv) > 1 else "ref_dpe_bridge"
g = np.load(os.path.join(ROOT, "tests/golden/ref_epochs_n9.npz"))
n, epochs, W = int(g["n"]), int(g["epochs"]), int(g["W"])
work = "/tmp/dbg_" + exe
sc, grid, files = mg.golden_files(work, epochs, n, list(g["offset"]))
dump = work + "/dump"
cmd = [os.path.join(ROOT, "oracle/_ref", exe), files["dat"], files["handoff"], files["rinex"], files["grid"], str(n), "5", str(epochs), dump, str(W), repr(sc.cfg.fs), "1", "0"]
r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
print(r.stdout[-2500:])
for e in range(epochs):
    for k in ("rx_time", "tx_time", "rc_start", "ri_start", "rc_end", "ri_end", "fc", "fi", "x_kk1", "enu2ecef", "sat_raw", "sat_states", "zval", "x_k1k1", "pos_scores", "time_grid"):
        f = os.path.join(dump, "e%03d_%s.bin" % (e, k))
        if not os.path.exists(f):
            print(e, k, "missing"); continue
        a = np.fromfile(f, dtype=np.float64); b = g["e%d_%s" % (e, k)]
        if a.shape != b.shape:
            print(e, k, "shape", a.shape, b.shape); continue
        d = np.abs(a - b)
        print(e, k, "max abs diff %.3e  (max |ref| %.3e)" % (d.max(), np.abs(b).max()), a[:4] if d.max() > 1e-3 * max(1, np.abs(b).max()) else "")
    for k in ("cp_start", "cp_end", "cp_ref", "cp_ref_tow"):
        a = np.fromfile(os.path.join(dump, "e%03d_%s.bin" % (e, k)), dtype=np.int32)
        print(e, k, "equal" if np.array_equal(a, g["e%d_%s" % (e, k)]) else (a, g["e%d_%s" % (e, k)]))
