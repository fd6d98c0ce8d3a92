"""Host-side mirror of the reference's dsp module / flow / console interface
(navlab-dpe-sdr_b200/host/, libdpe_flow.so).  CPU part: console grammar, host GPS code against
the oracle, N>1 partial combination over gloo.  GPU part: `newflow dpe / setparam / loadflow /
startflow` end to end on synthetic files against the oracle closed loop and the reference's own
epochs (tests/golden/ref_epochs_n9.npz)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers as H
from helpers import orc, synth
from oracle import chanmgr_oracle as chm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def flowapi():
    import dpe_pkg
    fa = dpe_pkg.submodule("flowapi")
    if not os.path.exists(fa.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return fa


def test_flow_library_exports(flowapi):
    import ctypes
    lib = ctypes.CDLL(flowapi.LIB_PATH)
    for name in flowapi.EXPORTS:
        assert hasattr(lib, name)


def test_console_grammar_and_param_typing(flowapi):
    sh = flowapi.Shell()
    assert sh.exec("bogus") == -1
    assert sh.exec("newf") == -1                       # missing option
    assert sh.exec("NEW dpe") == -1                    # mandatory part is NEWF
    assert sh.exec("newflow acq") == -1                # the stale dofile's flow type does not exist
    assert sh.exec("newf dpe rx") == 0                 # "DPE", 3 mandatory letters, case-insensitive
    assert sh.exec("loadflow rx") == 0
    assert sh.exec("loadflow rx") == -1                # already loaded
    assert sh.exec("setparam rx BatchCorrManifold PosGridDimSize 9") == 0
    assert sh.exec("setp rx BatchCorrManifold PosGridDimSize 9.0") == -1      # float literal into an INT param
    assert sh.exec("setp rx BatchCorrManifold GridDimSpacing 5.0") == 0       # FLOAT param
    assert sh.exec("setp rx SampleBlock SamplingFrequency 2.5e6") == -1       # reference quirk: no double literal
    assert sh.exec("setp rx SampleBlock SamplingFrequency 2.5e6d") == 0       # extension: d suffix = double
    assert sh.exec('setp rx SampleBlock Filename "/tmp/x.dat"') == 0
    assert sh.exec("setp rx cuEKF EnableEKF false") == 0
    assert sh.exec("setp rx NoSuchModule X 1") == -1
    assert sh.exec("printport rx BatchCorrManifold zVal") == 0
    assert sh.exec("printport rx BatchCorrManifold nope") == -1
    assert sh.exec("addalias second 0") == 0 and sh.exec("lsflow") == 0 and sh.exec("actalias") == 0
    z = sh.read_port("rx", "BatchCorrManifold", "zVal")
    assert z.shape == (8,) and not z.any()
    assert sh.exec("delflow rx") == 0 and sh.exec("loadflow rx") == -1
    assert sh.exec("quit -f") == 1
    sh.close()


def test_console_binary_runs_a_dofile(flowapi, tmp_path):
    do = tmp_path / "client.dofile"
    do.write_text("# comment\nnewflow dpe\nloadflow 0\nsetparam 0 BatchCorrManifold LPower 2\nlsflow\nquit -f\n")
    r = subprocess.run([flowapi.CONSOLE_PATH, "-f", str(do)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0
    assert "0: DPE" in r.stdout and "Completed LoadFlow" in r.stderr


def test_host_satellite_state_matches_oracle(flowapi):
    sc = H.scenario()
    nav = chm.read_rinex_nav(sc.cfg.rinex)
    for prn in synth.PRNS_12:
        for t in (414006.0, 414006.99, 417600.0):
            ref = chm.get_sat_pos(chm.select_eph(nav, prn, t), t)
            got = flowapi.sat_position(sc.cfg.rinex, prn, t)
            assert np.max(np.abs(got[:3] - ref[:3])) < 1e-6            # metres
            assert np.max(np.abs(got[4:7] - ref[4:7])) < 1e-9          # m/s
            assert abs(got[3] - ref[3]) < 1e-15 and abs(got[7] - ref[7]) < 1e-18


def test_host_grid_generation_matches_oracle(flowapi):
    for gt in (orc.GRID_UNIFORM, orc.GRID_ARTHURBASIS):
        for n in (5, 9, 12):
            ref, _ = orc.init_pos_grid([n] * 4, [1.0, 2.0, 0.5, 6.0], gt)
            got = flowapi.make_grid([n] * 4, [1.0, 2.0, 0.5, 6.0], gt)
            assert np.array_equal(got, ref)


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import dpe_pkg
    sharding = dpe_pkg.submodule("sharding")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc, iq, grid, ep = H.epoch_case(n=7, center_offset=(6.0, -4.0, 3.0, 7.0))
    if rank == 0:
        blk = torch.from_numpy(iq.copy()).view(torch.uint8)          # NCCL / gloo have no int16: ship the bytes
    else:
        blk = torch.zeros(2 * iq.shape[0], dtype=torch.uint8)
    dist.broadcast(blk, 0)                                           # the 20 ms block travels from rank 0
    iq_r = blk.view(torch.int16).numpy()
    lo, hi = sharding.shard_range(grid.shape[0], world, rank)
    bcs = orc.batch_corr_scores(iq_r, ep["prn"], ep["rc_start"], ep["ri_start"], ep["fc"], ep["fi"], ep["cp_start"],
                                ep["cp_ref"], ep["fs"])
    r = H.oracle_pos(bcs, grid[lo:hi], ep)                          # this rank's shard of the grid
    px, py, pz, pt = orc.candidate_ecef(grid[lo:hi], ep["center"], ep["enu2ecef"])
    part = torch.from_numpy(sharding.make_partial(r["scores"], np.stack([px, py, pz, pt], 1), lo))
    gathered = [torch.zeros_like(part) for _ in range(world)]
    dist.all_gather(gathered, part)
    res = {m: sharding.combine_partials(torch.stack(gathered).numpy(), m) for m in (0, 1)}
    if rank == 0:
        q.put((res, lo, hi))
    dist.destroy_process_group()


def test_sharded_estimate_over_gloo_world_size_2():
    """N>1 path on CPU: block broadcast, contiguous grid shards, all-gather of the per-rank
    partials, combination (lowest global index wins ties) == the single-rank answer."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 1000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    import queue as _queue
    out = None
    for _ in range(300):
        try:
            out = q.get(timeout=1)
            break
        except _queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out is not None
    res, lo, hi = out
    sc, iq, grid, ep = H.epoch_case(n=7, center_offset=(6.0, -4.0, 3.0, 7.0))
    bcs = H.oracle_bcs()
    one = H.oracle_pos(bcs, grid, ep)
    assert (lo, hi) == (0, (grid.shape[0] + 1) // 2)
    assert res[0]["argmax"] == one["argmax"]
    assert np.max(np.abs(res[0]["z"] - one["z"])) < 1e-9
    w = H.oracle_pos(bcs, grid, ep, weighted=True, per_time=False)
    assert np.max(np.abs(res[1]["z"] - w["z"])) < 1e-6
    assert abs(res[0]["sum_score"] - one["scores"].sum()) / one["scores"].sum() < 1e-12


def test_shard_ranges_cover_the_grid():
    import dpe_pkg
    sharding = dpe_pkg.submodule("sharding")
    for G in (1, 7, 6561, 390625, 6765201):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(G, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == G
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
    # ties: equal maxima on two ranks -> the lower global index
    a = sharding.make_partial([1.0, 5.0], [[1, 1, 1, 1], [2, 2, 2, 2]], 0)
    b = sharding.make_partial([5.0, 2.0], [[3, 3, 3, 3], [4, 4, 4, 4]], 2)
    assert sharding.combine_partials([b, a])["argmax"] == 1
    assert np.array_equal(sharding.combine_partials([a, b])["z"], [2, 2, 2, 2])


# ---------------------------------------------------------------------------------------------
def _write_scenario(tmp, n, epochs, first_block=1):
    sc = H.scenario()
    grid, _ = synth.uniform_grid(n, (5.0, 5.0, 5.0, 6.0))
    files = sc.write_files(str(tmp), epochs + 2, grid=grid, handoff_block=first_block)
    return sc, grid, files


def _drive(flowapi, files, n, extra=(), epochs=-1, xfile=None):
    sh = flowapi.Shell()
    cmds = ["newflow dpe rx", "loadflow rx",
            'setparam rx SampleBlock Filename "%s"' % files["dat"],
            'setparam rx DPInit HandoffFilename "%s"' % files["handoff"],
            'setparam rx DPInit RINEXFilename "%s"' % files["rinex"],
            'setparam rx BatchCorrManifold LoadPosGridFilename "%s"' % files["grid"],
            "setparam rx BatchCorrManifold LoadPosGrid true",
            "setparam rx BatchCorrManifold PosGridDimSize %d" % n,
            "setparam rx BatchCorrManifold VelGridDimSize 5",
            'setparam rx XECEFLogger Filename "%s"' % xfile] + list(extra)
    for c in cmds:
        assert sh.exec(c) == 0, c
    assert sh.run_blocking("rx", epochs) == 0
    return sh


@pytest.mark.gpu
@pytest.mark.parametrize("brute", [False, True])
def test_dpe_flow_end_to_end_against_oracle_closed_loop(flowapi, tmp_path, brute):
    n, epochs = 7, 6
    sc, grid, files = _write_scenario(tmp_path, n, epochs, first_block=0)
    xfile = str(tmp_path / "XFile.csv")
    extra = ["setparam rx DPInit InitDeltaX 6.0", "setparam rx DPInit InitDeltaY -4.0",
             "setparam rx DPInit InitDeltaZ 3.0", "setparam rx DPInit InitDeltaT 7.0"]
    if brute:
        extra.append("setparam rx BatchCorrManifold BruteForce true")
    sh = _drive(flowapi, files, n, extra, epochs, xfile)
    st = sh.stats("rx")
    assert st["run_count"] == epochs
    rows = np.loadtxt(xfile, delimiter=",")
    assert rows.shape == (epochs, 8)

    # oracle closed loop: cuChanMgr restatement + BCS + BCM + pass-through EKF on the same files
    nav = chm.read_rinex_nav(files["rinex"])
    h = chm.read_handoff(files["handoff"])
    x = h["X_ECEF"].copy()
    x[:4] += (6.0, -4.0, 3.0, 7.0)
    _, tg = synth.uniform_grid(n, (5.0, 5.0, 5.0, 6.0))
    vgrid, _ = synth.uniform_grid(5, 1.0)                  # VelGridDimSize 5, GridDimSpacing 1.0 (the default)
    ch = chm.chanmgr_start(nav, h, sc.cfg.T, x)
    for e in range(epochs):
        sat, R = chm.grid_prep(ch, x, tg)
        iq = sc.block(e)
        bcs = orc.batch_corr_scores(iq, ch.prn, ch.rc_start, ch.ri_start, ch.fc, ch.fi, ch.cp_start, ch.cp_ref,
                                    sc.cfg.fs, want_carrier=True)
        r = orc.pos_meas_ml(bcs["code_scores"], grid, x, R, sat, len(tg), ch.fc, ch.rc_end, ch.cp_ref_tow, ch.cp_end,
                            ch.cp_ref, ch.rx_time, sc.cfg.fs, sc.S)
        v = orc.vel_meas_ml(bcs["carr_scores"], vgrid, x, R, sat, len(tg), ch.fi, 1, sc.cfg.fs, bcs["n_fft"])
        x = np.concatenate([r["z"], v["z"]])
        assert np.max(np.abs(rows[e, 4:] - x[4:])) < 1e-5, "velocity fix, epoch %d" % e
        # logged fix (6 decimals) vs oracle: position 0.1 m, clock 1 ns = 0.2998 m
        assert np.max(np.abs(rows[e, :3] - x[:3])) < 0.1, "epoch %d" % e
        assert abs(rows[e, 3] - x[3]) < 0.2998
        chm.chanmgr_update(ch, nav, x)
    # host channel manager == oracle channel manager after the run
    assert np.max(np.abs(sh.read_port("rx", "cuChanMgr", "CodePhaseEnd") - ch.rc_end)) < 1e-6
    assert np.max(np.abs(sh.read_port("rx", "cuChanMgr", "CarrierFrequency") - ch.fi)) < 1e-6
    assert np.array_equal(sh.read_port("rx", "cuChanMgr", "cpElapsedEnd").astype(int), ch.cp_end)
    sh.close()


@pytest.mark.gpu
def test_dpe_flow_fed_over_tcp_equals_the_file_source(flowapi, tmp_path):
    """SampleBlock's socket source (`InputSourceType` 1, `Hostname`, `PortNo`; sampleblock.cu:134-156):
    the same capture streamed by a TCP server in ragged chunks gives the same fixes as the file."""
    import socket
    import threading
    n, epochs = 5, 5
    sc, grid, files = _write_scenario(tmp_path, n, epochs, first_block=0)
    x_file, x_sock = str(tmp_path / "XFile_file.csv"), str(tmp_path / "XFile_sock.csv")
    _drive(flowapi, files, n, (), epochs, x_file).close()

    srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
    srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
    srv.bind(("127.0.0.1", 0))
    srv.listen(1)
    port = srv.getsockname()[1]

    def serve():
        conn, _ = srv.accept()
        data = open(files["dat"], "rb").read()
        try:
            for lo in range(0, len(data), 70001):                  # not a multiple of the 200000-byte block
                conn.sendall(data[lo:lo + 70001])
        except OSError:
            pass                                                   # the flow stopped after `epochs` blocks
        finally:
            conn.close()

    th = threading.Thread(target=serve, daemon=True)
    th.start()
    extra = ["setparam rx SampleBlock InputSourceType \\x01", 'setparam rx SampleBlock Hostname "127.0.0.1"',
             "setparam rx SampleBlock PortNo %d" % port]
    _drive(flowapi, files, n, extra, epochs, x_sock).close()
    th.join(timeout=10)
    srv.close()
    a, b = np.loadtxt(x_file, delimiter=","), np.loadtxt(x_sock, delimiter=",")
    assert a.shape == (epochs, 8) and np.array_equal(a, b)


@pytest.mark.gpu
def test_dpe_flow_reproduces_the_reference_epochs(flowapi, tmp_path):
    """Same files, same offset as oracle/make_golden_ref.py fed to the UNMODIFIED reference: the
    fixes and the channel parameters handed to BCS / BCM must agree epoch by epoch."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_epochs_n9.npz"))
    n, epochs, first = int(g["n"]), int(g["epochs"]), int(g["first_block"])
    sc, grid, files = _write_scenario(tmp_path, n, epochs + first, first_block=first)
    off = g["offset"]
    extra = ["setparam rx DPInit InitDeltaX %r" % float(off[0]), "setparam rx DPInit InitDeltaY %r" % float(off[1]),
             "setparam rx DPInit InitDeltaZ %r" % float(off[2]), "setparam rx DPInit InitDeltaT %r" % float(off[3])]
    xfile = str(tmp_path / "XFile.csv")
    sh = _drive(flowapi, files, n, extra, epochs, xfile)
    rows = np.loadtxt(xfile, delimiter=",")
    for e in range(epochs):
        ref = g["e%d_x_k1k1" % e]
        assert np.max(np.abs(rows[e, :3] - ref[:3])) < 0.1 and abs(rows[e, 3] - ref[3]) < 0.2998
        assert np.max(np.abs(rows[e, 4:] - ref[4:])) < 1e-5        # velocity manifold arg-max (5^4 grid, 1 m/s)
    sh.close()


def test_host_file_readers_on_the_reference_demo_files(flowapi, tmp_path):
    """DPInit's handoff-CSV grammar (dpinit.cpp:247-400) on the reference's own shipped file
    (tests/golden/handoff_params_usrp6.csv = demofiles/handoff_params_usrp6.csv) against the oracle's
    reader, and the `x,y,z,delta_t` grid CSV (batchcorrmanifold.cu:2433-2444) round trip."""
    path = os.path.join(ROOT, "tests", "golden", "handoff_params_usrp6.csv")
    got, ref = flowapi.read_handoff(path), chm.read_handoff(path)
    assert list(got["prn"].astype(int)) == [2, 3, 6, 12, 17, 19, 24, 28]         # prn_list row of the shipped file
    assert got["rxTime"] == ref["rxTime"] and got["bytes_read"] == ref["bytes_read"]
    assert np.array_equal(got["X_ECEF"][:len(ref["X_ECEF"])], ref["X_ECEF"])
    for k_host, k_ref in (("rc", "rc"), ("ri", "ri"), ("fc", "fc"), ("fi", "fi"), ("cp", "cp"),
                          ("cp_timestamp", "cp_timestamp"), ("TOW", "TOW")):
        assert np.array_equal(got[k_host], np.asarray(ref[k_ref], dtype=float)), k_host
    assert flowapi.read_handoff.__doc__ and pytest.raises(RuntimeError, flowapi.read_handoff, str(tmp_path / "missing.csv"))
    grid, _ = synth.uniform_grid(5, (5.0, 5.0, 5.0, 6.0))
    p = str(tmp_path / "rngrid.csv")
    np.savetxt(p, grid, delimiter=",", fmt="%.17g")
    assert np.array_equal(flowapi.read_grid(p), grid)


@pytest.mark.parametrize("grid_type", [0, 2])
def test_host_grid_axes_match_the_reference_generator(flowapi, grid_type):
    """BCM_InitPosGrid (batchcorrmanifold.cu:148-255), Uniform and ArthurBasis: the axis values the UNMODIFIED
    reference generated for a 9^4 grid with GridDimSpacing 2.5 (its TimeGrid port, tests/golden/ref_grid_t*_n9.npz,
    oracle/make_golden_ref.py --gen-grid) against the host generator, all four axes, t fastest."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_grid_t%d_n9.npz" % grid_type))
    n, sp = int(g["n"]), float(g["spacing"])
    ref_axis = g["e0_time_grid"]
    grid = flowapi.make_grid([n] * 4, [sp] * 4, grid_type).reshape(n, n, n, n, 4)
    assert np.array_equal(grid[0, 0, 0, :, 3], ref_axis)             # clock axis: the fastest index
    assert np.array_equal(grid[0, 0, :, 0, 2], ref_axis)
    assert np.array_equal(grid[0, :, 0, 0, 1], ref_axis)
    assert np.array_equal(grid[:, 0, 0, 0, 0], ref_axis)             # x: the slowest index
    if grid_type == 2:
        assert ref_axis[0] == -6 * sp and ref_axis[1] == -3 * sp      # stretched outer quarter, uniform core


@pytest.mark.gpu
@pytest.mark.parametrize("grid_type", [0, 2])
def test_dpe_flow_with_a_generated_grid(flowapi, tmp_path, grid_type):
    """LoadPosGrid false: the flow generates its own position grid (GridType, GridDimSpacing).  Its TimeGrid port
    equals the one the reference produced with the same parameters (tests/golden/ref_grid_t*_n9.npz), and the
    fixes equal those of the same flow fed the same grid through a CSV file.  (The reference's own fixes of
    that run are not compared: with identical inputs its centre-candidate score at epoch 0 was 4.03e7 where its
    other run -- ref_epochs_n9.npz -- and the oracle have 6.14e7: the flip / no-flip race of BCS_ChooseCodeCorr,
    batchcorrscores.cu:508-541, hit channels 0 and 7.)"""
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_grid_t%d_n9.npz" % grid_type))
    n, epochs, first, sp = int(g["n"]), int(g["epochs"]), int(g["first_block"]), float(g["spacing"])
    sc, _, files = _write_scenario(tmp_path, n, epochs + first, first_block=first)
    off = g["offset"]
    init = ["setparam rx DPInit InitDeltaX %r" % float(off[0]), "setparam rx DPInit InitDeltaY %r" % float(off[1]),
            "setparam rx DPInit InitDeltaZ %r" % float(off[2]), "setparam rx DPInit InitDeltaT %r" % float(off[3]),
            "setparam rx BatchCorrManifold GridType %d" % grid_type,
            "setparam rx BatchCorrManifold GridDimSpacing %r" % sp]
    x_gen, x_csv = str(tmp_path / "XFile_gen.csv"), str(tmp_path / "XFile_csv.csv")
    sh = _drive(flowapi, files, n, init + ["setparam rx BatchCorrManifold LoadPosGrid false"], epochs, x_gen)
    assert np.array_equal(sh.read_port("rx", "BatchCorrManifold", "TimeGrid"), g["e0_time_grid"])
    sh.close()
    grid = flowapi.make_grid([n] * 4, [sp] * 4, grid_type)
    np.savetxt(files["grid"], grid, delimiter=",", fmt="%.17g")
    _drive(flowapi, files, n, init, epochs, x_csv).close()
    a, b = np.loadtxt(x_gen, delimiter=","), np.loadtxt(x_csv, delimiter=",")
    assert a.shape == (epochs, 8) and np.array_equal(a, b)


@pytest.mark.gpu
def test_dpe_flow_with_the_kalman_filter_enabled_matches_the_reference(flowapi, tmp_path):
    """`setparam <flow> cuEKF EnableEKF true` (the reference's 8-state KF: H = I, random-walk F,
    speed-adaptive Q, cuekf.cu:42-81,625-742) against the UNMODIFIED reference run with the same switch
    (tests/golden/ref_ekf_n9.npz, oracle/make_golden_ref.py --ekf)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_ekf_n9.npz"))
    n, epochs, first = int(g["n"]), int(g["epochs"]), int(g["first_block"])
    sc, grid, files = _write_scenario(tmp_path, n, epochs + first, first_block=first)
    off = g["offset"]
    extra = ["setparam rx DPInit InitDeltaX %r" % float(off[0]), "setparam rx DPInit InitDeltaY %r" % float(off[1]),
             "setparam rx DPInit InitDeltaZ %r" % float(off[2]), "setparam rx DPInit InitDeltaT %r" % float(off[3]),
             "setparam rx cuEKF EnableEKF true"]
    xfile = str(tmp_path / "XFile.csv")
    sh = _drive(flowapi, files, n, extra, epochs, xfile)
    rows = np.loadtxt(xfile, delimiter=",")
    assert rows.shape == (epochs, 8)
    for e in range(epochs):
        ref = g["e%d_x_k1k1" % e]
        assert np.max(np.abs(rows[e, :3] - ref[:3])) < 1e-3 and abs(rows[e, 3] - ref[3]) < 1e-3, "epoch %d" % e
        assert np.max(np.abs(rows[e, 4:] - ref[4:])) < 1e-5
    # the prediction handed to the next epoch's grid (x_{k|k-1}) moved with the clock drift
    xkk1 = sh.read_port("rx", "cuEKF", "xCurrkk1")
    assert abs((xkk1[3] - g["e0_x_kk1"][3]) - epochs * 0.02 * g["e0_x_kk1"][7]) < 1e-3
    sh.close()


REF_RINEX = "/root/reference/demofiles/nist1860.18n"


@pytest.mark.skipif(not os.path.exists(REF_RINEX), reason="the reference checkout is not on this box")
def test_host_rinex_reader_and_nearest_toe_on_the_reference_demo_file(flowapi):
    """The reference's shipped RINEX 2.10 nav file holds 26 TOE sets (172752 ... 432000 s) with 1 to 15
    satellites each: ReadRinexNav's grouping (rinexparse.cpp:199-217) and the brute-force nearest-TOE
    selection (cuchanmgr.cu:276-292) of the host C++ against the oracle, for every PRN across the day,
    including times exactly between two sets and the 16-second-apart pairs (…584 / …600)."""
    nav = chm.read_rinex_nav(REF_RINEX)
    assert len(nav) == 26 and {s.toes for s in nav} >= {417600.0, 395968.0, 395984.0, 396000.0, 172752.0}
    toes = sorted(s.toes for s in nav)
    times = [345000.0 + 1777.0 * k for k in range(52)]
    times += [0.5 * (a + b) for a, b in zip(toes[:-1], toes[1:])]           # ties / midpoints
    times += [t + d for t in (395968.0, 395984.0, 396000.0, 410384.0) for d in (-8.0, -0.001, 0.0, 0.001, 8.0)]
    n = 0
    for prn in range(1, 33):
        for t in times:
            e = chm.select_eph(nav, prn, t)
            assert e is not None
            ref = chm.get_sat_pos(e, t)
            got = flowapi.sat_position(REF_RINEX, prn, t)
            assert np.max(np.abs(got[:3] - ref[:3])) < 1e-5, (prn, t, e.toes)      # a wrong set is off by metres to km
            assert np.max(np.abs(got[4:7] - ref[4:7])) < 1e-8
            assert abs(got[3] - ref[3]) < 1e-14 and abs(got[7] - ref[7]) < 1e-17
            n += 1
    assert n == 32 * len(times)


@pytest.mark.gpu
@pytest.mark.parametrize("brute", [False, True])
def test_dpe_flow_sharded_over_gpus_equals_the_single_gpu_flow(flowapi, tmp_path, brute):
    """`setparam <flow> BatchCorrManifold NumGPUs N`: the console flow shards the grid over N GPUs of this
    process (one host thread + one NCCL rank per GPU, one dpe_epoch_run_dist per rank and epoch) and logs
    the same fixes as the single-GPU flow."""
    import torch
    ngpu = min(torch.cuda.device_count(), 4)
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    n, epochs = 7, 5
    sc, grid, files = _write_scenario(tmp_path, n, epochs, first_block=0)
    extra = ["setparam rx DPInit InitDeltaX 6.0", "setparam rx DPInit InitDeltaY -4.0",
             "setparam rx DPInit InitDeltaZ 3.0", "setparam rx DPInit InitDeltaT 7.0"]
    if brute:
        extra.append("setparam rx BatchCorrManifold BruteForce true")
    x1, xn = str(tmp_path / "X1.csv"), str(tmp_path / "XN.csv")
    _drive(flowapi, files, n, extra, epochs, x1).close()
    sh = _drive(flowapi, files, n, extra + ["setparam rx BatchCorrManifold NumGPUs %d" % ngpu], epochs, xn)
    assert sh.stats("rx")["run_count"] == epochs
    sh.close()
    a, b = np.loadtxt(x1, delimiter=","), np.loadtxt(xn, delimiter=",")
    assert a.shape == b.shape == (epochs, 8)
    assert np.array_equal(a, b)
    # a device ordinal that does not exist is refused at start
    sh = flowapi.Shell()
    for c in ["newflow dpe rx", "loadflow rx", 'setparam rx SampleBlock Filename "%s"' % files["dat"],
              'setparam rx DPInit HandoffFilename "%s"' % files["handoff"],
              'setparam rx DPInit RINEXFilename "%s"' % files["rinex"],
              "setparam rx BatchCorrManifold PosGridDimSize 5", "setparam rx BatchCorrManifold Device 64"]:
        assert sh.exec(c) == 0
    assert sh.run_blocking("rx", 1) != 0
    sh.close()


def test_flow_start_failure_is_clean_and_repeatable(flowapi, tmp_path):
    """A flow whose modules cannot all start (here: the grid file does not exist; on a box without a GPU the sample
    ring fails first) returns an error from startflow, stops every module it had started -- the failing one included:
    reader thread, buffers, stream, context -- and can be started again or destroyed without hanging."""
    sc, grid, files = _write_scenario(tmp_path, 3, 1, first_block=0)
    sh = flowapi.Shell()
    for c in ["newflow dpe rx", "loadflow rx", 'setparam rx SampleBlock Filename "%s"' % files["dat"],
              'setparam rx DPInit HandoffFilename "%s"' % files["handoff"],
              'setparam rx DPInit RINEXFilename "%s"' % files["rinex"],
              'setparam rx BatchCorrManifold LoadPosGridFilename "%s"' % str(tmp_path / "no_such_grid.csv"),
              "setparam rx BatchCorrManifold LoadPosGrid true", "setparam rx BatchCorrManifold PosGridDimSize 3",
              'setparam rx XECEFLogger Filename "%s"' % str(tmp_path / "X.csv")]:
        assert sh.exec(c) == 0, c
    assert sh.run_blocking("rx", 1) != 0
    assert sh.run_blocking("rx", 1) != 0                 # nothing was left half started
    assert sh.stats("rx")["run_count"] == 0
    sh.close()


def test_host_readers_survive_malformed_files(flowapi, tmp_path):
    """Truncated, bit-flipped, shuffled and random handoff / RINEX / grid files: the readers return an error or a
    result, never crash or hang (run in a child process so that a crash would fail this test, not the session)."""
    import random
    rnd = random.Random(20180704)
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    bases = {"handoff": open(os.path.join(here, "golden", "handoff_params_usrp6.csv"), "rb").read(),
             "rinex": open(os.path.join(root, "navlab-dpe-sdr_b200", "data", "brdc_toe417600.18n"), "rb").read(),
             "grid": b"\n".join(b"%f,%f,%f,%f" % (i, i + 1, i + 2, i + 3) for i in range(81))}
    jobs = []
    for kind, base in bases.items():
        for t in range(25):
            b = bytearray(base)
            mode = t % 5
            if mode == 0:
                b = b[:rnd.randrange(0, len(b))]
            elif mode == 1:
                for _ in range(20):
                    b[rnd.randrange(len(b))] = rnd.randrange(256)
            elif mode == 2:
                b = bytes(rnd.randrange(256) for _ in range(rnd.randrange(0, 2000)))
            elif mode == 3:
                i = rnd.randrange(len(b))
                b[i:i] = b"9" * rnd.randrange(1, 400)
            else:
                lines = bytes(b).split(b"\n")
                rnd.shuffle(lines)
                b = b"\n".join(lines[:rnd.randrange(1, len(lines) + 1)])
            p = tmp_path / ("%s_%d" % (kind, t))
            p.write_bytes(bytes(b))
            jobs.append("%s %s" % (kind, p))
    (tmp_path / "jobs.txt").write_text("\n".join(jobs))
    child = ("import sys\nsys.path.insert(0, %r)\nimport dpe_pkg\nf = dpe_pkg.submodule('flowapi')\nn = 0\n"
             "for line in open(sys.argv[1]):\n    kind, path = line.split()\n    try:\n"
             "        {'handoff': f.read_handoff, 'grid': f.read_grid, 'rinex': lambda p: f.sat_position(p, 5, 417600.0)}[kind](path)\n"
             "    except Exception:\n        pass\n    n += 1\nprint('done', n)\n" % root)
    r = subprocess.run([sys.executable, "-c", child, str(tmp_path / "jobs.txt")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and ("done %d" % len(jobs)) in r.stdout, r.stderr[-1000:]
