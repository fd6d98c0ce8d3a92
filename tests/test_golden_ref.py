"""Golden vectors from the UNMODIFIED reference (CUDARecv rebuilt for sm_100a, run on a
B200 through oracle/ref_driver.cu; tests/golden/ref_epochs_n9.npz, made by
oracle/make_golden_ref.py).  CPU part: pins the NumPy oracle.  GPU part: pins the CUDA path.

Reference quirk that shapes the CodeScores check: BCS_ChooseCodeCorr has an inter-block race
(batchcorrscores.cu:508-541) -- only the threads of block 0 are guaranteed
to see the epoch's flip / no-flip decision, the other 7 blocks may copy the row using a stale
flag.  A reference row is therefore element-wise either the no-flip or the flipped correlogram;
rows where the race did not strike equal the deterministic rule to ~1e-15.
"""
import os

import numpy as np
import pytest

import helpers as H
from helpers import orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_epochs_n9.npz")


class Gold:
    def __init__(self):
        g = np.load(GOLD)
        self.g = g
        self.C, self.T, self.S, self.W = int(g["C"]), int(g["T"]), int(g["S"]), int(g["W"])
        self.fs, self.epochs, self.grid = float(g["fs"]), int(g["epochs"]), g["grid"]
        self.NL = 2 * self.W + 2

    def k(self, e, name):
        return self.g["e%d_%s" % (e, name)]

    def ref_window(self, e):
        w = self.k(e, "code_scores_win").reshape(self.C, self.NL, 2)
        return w[..., 0] + 1j * w[..., 1]

    def epoch_dict(self, e):
        k = lambda n: self.k(e, n)
        return dict(prn=k("prn"), rc_start=k("rc_start"), ri_start=k("ri_start"), fc=k("fc"), fi=k("fi"),
                    cp_start=k("cp_start"), cp_ref=k("cp_ref"), rc_end=k("rc_end"), cp_end=k("cp_end"),
                    cp_ref_tow=k("cp_ref_tow"), rx_time=float(k("rx_time")[0]), center=k("x_kk1"),
                    enu2ecef=k("enu2ecef"), sat_states=k("sat_states").reshape(-1, 8), doppler_sign=1,
                    S=self.S, fs=self.fs, time_dim=self.T)

    def both_correlograms(self, e):
        """Oracle no-flip and flipped windows [C][NL] + decision, from the reference's own inputs."""
        ep = self.epoch_dict(e)
        S, W, NL = self.S, self.W, self.NL
        x = orc.iq_to_complex(self.k(e, "iq"))
        t = orc.time_idcs(S, self.fs)
        nxt = orc.nav_bit_boundary(ep["cp_start"], ep["cp_ref"], ep["rc_start"], ep["fc"], self.fs)
        nf = np.empty((self.C, NL), complex)
        fl = np.empty((self.C, NL), complex)
        for c in range(self.C):
            w = orc.doppler_wipeoff(ep["fi"][c], ep["ri_start"][c], t)
            _, r_nf, r_fl = orc.code_replica(int(ep["prn"][c]), t, ep["fc"][c], ep["rc_start"][c], int(nxt[c]), S)
            xf = np.fft.fft(x * w)
            sl = slice(S // 2 - W, S // 2 - W + NL)
            nf[c] = np.fft.fftshift(np.fft.ifft(xf * np.conj(np.fft.fft(r_nf))))[sl]
            fl[c] = np.fft.fftshift(np.fft.ifft(xf * np.conj(np.fft.fft(r_fl))))[sl]
        edge = (nxt > 0) & (nxt < S)
        keep_nf = ~edge | (np.abs(nf[:, W]) > np.abs(fl[:, W]))
        return nf, fl, keep_nf


@pytest.fixture(scope="module")
def gold():
    return Gold()


def test_golden_file_shape(gold):
    assert (gold.C, gold.S, gold.T, gold.epochs) == (8, 50000, 9, 3)
    assert gold.grid.shape == (9 ** 4, 4)


def test_oracle_correlogram_matches_reference_rows(gold):
    exact_rows = 0
    for e in range(gold.epochs):
        ref = gold.ref_window(e)
        nf, fl, keep_nf = gold.both_correlograms(e)
        scale = np.max(np.abs(ref), axis=1)[:, None]
        d_nf, d_fl = np.abs(ref - nf) / scale, np.abs(ref - fl) / scale
        # every element of a reference row is one of the two correlograms (race => mixture)
        assert np.max(np.minimum(d_nf, d_fl)) < 1e-12
        chosen = np.where(keep_nf[:, None], nf, fl)
        exact_rows += int(np.sum(np.max(np.abs(ref - chosen) / scale, axis=1) < 1e-12))
    assert exact_rows >= gold.epochs * gold.C // 2          # the race strikes a minority of rows


def test_oracle_manifold_matches_reference_scores_and_fix(gold):
    for e in range(gold.epochs):
        ep = gold.epoch_dict(e)
        full = np.zeros((gold.C, gold.S), complex)
        full[:, gold.S // 2 - gold.W: gold.S // 2 - gold.W + gold.NL] = gold.ref_window(e)
        r = orc.pos_meas_ml(full, gold.grid, ep["center"], ep["enu2ecef"], ep["sat_states"], gold.T, ep["fc"],
                            ep["rc_end"], ep["cp_ref_tow"], ep["cp_end"], ep["cp_ref"], ep["rx_time"], gold.fs,
                            gold.S)
        ps = gold.k(e, "pos_scores")
        assert r["valid"].all()
        assert np.max(np.abs(r["scores"] - ps) / ps) < 1e-9          # FP64 both sides (FMA contraction only)
        assert r["argmax"] == int(np.argmax(ps))
        assert np.max(np.abs(r["z"] - gold.k(e, "zval")[:4])) < 1e-6
        # EKF is a pass-through in the DPE flow (cuekf.cu:147-159): x_k1k1[0:4] == zVal[0:4]
        assert np.array_equal(gold.k(e, "x_k1k1")[:4], gold.k(e, "zval")[:4])


def _bcm_gold(L):
    g = np.load(os.path.join(os.path.dirname(GOLD), "ref_bcm_L%d_n9.npz" % L))
    C, T, S, W, fs = int(g["C"]), int(g["T"]), int(g["S"]), int(g["W"]), float(g["fs"])
    k = lambda n: g["e0_" + n]
    w = k("code_scores_win").reshape(C, 2 * W + 2, 2)
    ep = dict(prn=k("prn"), rc_start=k("rc_start"), ri_start=k("ri_start"), fc=k("fc"), fi=k("fi"),
              cp_start=k("cp_start"), cp_ref=k("cp_ref"), rc_end=k("rc_end"), cp_end=k("cp_end"),
              cp_ref_tow=k("cp_ref_tow"), rx_time=float(k("rx_time")[0]), center=k("x_kk1"), enu2ecef=k("enu2ecef"),
              sat_states=k("sat_states").reshape(-1, 8), doppler_sign=1, S=S, fs=fs, time_dim=T)
    return dict(C=C, T=T, S=S, W=W, fs=fs, grid=g["grid"], win=w[..., 0] + 1j * w[..., 1], ep=ep,
                pos_scores=k("pos_scores"), zval=k("zval"), lpower=int(g["lpower"]))


@pytest.mark.parametrize("L", [2, 3])
def test_oracle_manifold_matches_reference_for_other_lpower(L):
    """`setparam <flow> BatchCorrManifold LPower 2|3` in the UNMODIFIED reference (score = sum_prn |v|^L,
    batchcorrmanifold.cu:1816; tests/golden/ref_bcm_L*_n9.npz, oracle/make_golden_ref.py --lpower)."""
    b = _bcm_gold(L)
    assert b["lpower"] == L
    ep, S, W = b["ep"], b["S"], b["W"]
    full = np.zeros((b["C"], S), complex)
    full[:, S // 2 - W: S // 2 - W + 2 * W + 2] = b["win"]
    r = orc.pos_meas_ml(full, b["grid"], ep["center"], ep["enu2ecef"], ep["sat_states"], b["T"], ep["fc"], ep["rc_end"],
                        ep["cp_ref_tow"], ep["cp_end"], ep["cp_ref"], ep["rx_time"], b["fs"], S, lpower=L)
    assert np.max(np.abs(r["scores"] - b["pos_scores"]) / b["pos_scores"]) < 1e-9
    assert r["argmax"] == int(np.argmax(b["pos_scores"]))
    assert np.max(np.abs(r["z"] - b["zval"][:4])) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("L", [2, 3])
def test_cuda_manifold_matches_reference_for_other_lpower(capi, L):
    b = _bcm_gold(L)
    G = b["grid"].shape[0]
    ctx = capi.Context(fs=b["fs"], S=b["S"], max_chan=b["C"], G=G, time_dim=b["T"], lag_halfwidth=b["W"], lpower=L)
    ctx.grid_set(b["grid"])
    ctx.epoch_set(b["ep"])
    ctx.code_scores_set(b["win"])
    ctx.score_pos(capi.SCORE_LOOKUP, capi.SAT_MIDDLE)
    ctx.estimate(capi.EST_ARGMAX)
    res = ctx.result_fetch()
    scores = ctx.copy_out(capi.PTR_POS_SCORES, np.float64, G)
    assert np.max(np.abs(scores - b["pos_scores"]) / b["pos_scores"]) < 1e-9
    assert res.argmax == int(np.argmax(b["pos_scores"]))
    assert np.max(np.abs(np.array(res.z[:4]) - b["zval"][:4])) < 1e-6
    ctx.close()


def test_oracle_carrier_spectrum_matches_reference(gold):
    """Velocity branch: DC removal, chosen replica, zero-padded 524288-point spectrum
    (batchcorrscores.cu:1158-1180) -- the reference's CarrScores window around 0 Hz."""
    g = gold.g
    Wd, n_fft = int(g["Wd"]), int(g["n_fft"])
    NBd = 2 * Wd + 2
    assert n_fft == 8 * (1 << int(np.ceil(np.log2(gold.S))))
    for e in range(gold.epochs):
        ep = gold.epoch_dict(e)
        bcs = orc.batch_corr_scores(gold.k(e, "iq"), ep["prn"], ep["rc_start"], ep["ri_start"], ep["fc"], ep["fi"],
                                    ep["cp_start"], ep["cp_ref"], gold.fs, want_carrier=True)
        ref = gold.k(e, "carr_scores_win").reshape(gold.C, NBd, 2)
        ref = ref[..., 0] + 1j * ref[..., 1]
        mine = bcs["carr_scores"][:, n_fft // 2 - Wd: n_fft // 2 - Wd + NBd]
        assert np.max(np.abs(mine - ref) / np.max(np.abs(ref), axis=1)[:, None]) < 1e-12
        # velocity arg-max of the reference (5^4 grid, 1 m/s) from the oracle's scores
        vgrid, _ = H.synth.uniform_grid(int(g["vel_dim"]), 1.0)
        v = orc.vel_meas_ml(bcs["carr_scores"], vgrid, ep["center"], ep["enu2ecef"], ep["sat_states"], gold.T,
                            ep["fi"], 1, gold.fs, n_fft)
        assert np.max(np.abs(v["z"] - gold.k(e, "zval")[4:8])) < 1e-9


def test_channel_manager_oracle_reproduces_reference_epoch_inputs(gold):
    """oracle/chanmgr_oracle.py (cuChanMgr restatement) started from the same handoff must
    reproduce what the reference's cuChanMgr handed to BCS/BCM at epoch 0."""
    from oracle import chanmgr_oracle as chm
    sc = H.scenario()
    nav = chm.read_rinex_nav(sc.cfg.rinex)
    h = sc.handoff(int(gold.g["first_block"]))
    x0 = h["X_ECEF"].copy()
    x0[:4] += gold.g["offset"]
    ch = chm.chanmgr_start(nav, h, sc.cfg.T, x0)
    k = lambda n: gold.k(0, n)
    assert abs(ch.rx_time - float(k("rx_time")[0])) < 1e-9
    assert np.array_equal(ch.cp_start, k("cp_start")) and np.array_equal(ch.cp_end, k("cp_end"))
    assert np.max(np.abs(ch.rc_start - k("rc_start"))) < 1e-9
    assert np.max(np.abs(ch.rc_end - k("rc_end"))) < 1e-6
    assert np.max(np.abs(ch.ri_start - k("ri_start"))) < 1e-9
    assert np.max(np.abs(ch.fc - k("fc"))) < 1e-6 and np.max(np.abs(ch.fi - k("fi"))) < 1e-6
    sat, R = chm.grid_prep(ch, k("x_kk1"), k("time_grid"))
    assert np.max(np.abs(R - k("enu2ecef"))) < 1e-12
    ref_sat = k("sat_states").reshape(-1, 8)
    assert np.max(np.abs(sat[:, :3] - ref_sat[:, :3])) < 1e-3          # metres
    assert np.max(np.abs(sat[:, 3] - ref_sat[:, 3])) < 1e-12           # seconds


# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_cuda_path_matches_reference_golden(gold, capi):
    n_race_free = 0
    for e in range(gold.epochs):
        ep = gold.epoch_dict(e)
        ctx = capi.Context(fs=gold.fs, S=gold.S, max_chan=gold.C, G=gold.grid.shape[0], time_dim=gold.T,
                           lag_halfwidth=gold.W, flags=capi.FLAG_BRUTE_TILES)
        ctx.grid_set(gold.grid)
        ctx.block_stage(gold.k(e, "iq"))
        ctx.epoch_set(ep)
        ctx.replica_prepare()
        ctx.correlogram()
        cs = ctx.copy_out(capi.PTR_CODE_SCORES, np.float64, gold.C * gold.NL * 2).reshape(gold.C, gold.NL, 2)
        got = cs[..., 0] + 1j * cs[..., 1]
        ref = gold.ref_window(e)
        nf, fl, keep_nf = gold.both_correlograms(e)
        _, no_flip = ctx.channel_flags(gold.C)
        assert np.array_equal(no_flip.astype(bool), keep_nf)
        scale = np.max(np.abs(ref), axis=1)[:, None]
        chosen = np.where(keep_nf[:, None], nf, fl)
        race_free = np.max(np.abs(ref - chosen) / scale, axis=1) < 1e-12
        n_race_free += int(race_free.sum())                           # the race strikes different rows every run
        # CUDA correlogram (FP32 products, FP64 sums) against the reference's own rows
        if race_free.any():
            assert np.max(np.abs(got[race_free] - ref[race_free]) / scale[race_free]) < 1e-6
        # raced rows: every CUDA element still equals the deterministic choice, and the reference
        # element is one of the two candidates
        assert np.max(np.abs(got - chosen) / scale) < 1e-6

        # BatchCorrManifold stage on the reference's own CodeScores: FP64 end to end
        ctx.code_scores_set(ref)
        ctx.score_pos(capi.SCORE_LOOKUP, capi.SAT_MIDDLE)
        ctx.estimate(capi.EST_ARGMAX)
        res = ctx.result_fetch()
        ps = gold.k(e, "pos_scores")
        scores = ctx.copy_out(capi.PTR_POS_SCORES, np.float64, gold.grid.shape[0])
        assert np.max(np.abs(scores - ps) / ps) < 1e-9
        assert res.argmax == int(np.argmax(ps))
        assert np.max(np.abs(np.array(res.z[:4]) - gold.k(e, "zval")[:4])) < 1e-6     # bar: 0.1 m / 0.2998 m
        f, _ = ctx.debug_bins(0, gold.grid.shape[0], gold.C)
        f_ref = orc.pos_bins(gold.grid, ep["center"], ep["enu2ecef"], ep["sat_states"], gold.T, ep["fc"],
                             ep["rc_end"], ep["cp_ref_tow"], ep["cp_end"], ep["cp_ref"], ep["rx_time"], gold.fs,
                             gold.S)[2]
        assert np.array_equal(f, f_ref)                                # code-phase bins bit-exact
        with pytest.raises(capi.DpeError):
            ctx.score_pos(capi.SCORE_BRUTE)                            # foreign correlogram: lookup only
        ctx.close()

        # velocity manifold against the reference's CarrScores window and velocity fix
        Wd, n_fft, vd = int(gold.g["Wd"]), int(gold.g["n_fft"]), int(gold.g["vel_dim"])
        vgrid, _ = H.synth.uniform_grid(vd, 1.0)
        cv = capi.Context(fs=gold.fs, S=gold.S, max_chan=gold.C, G=gold.grid.shape[0], time_dim=gold.T,
                          lag_halfwidth=gold.W, Gv=vgrid.shape[0], dopp_halfwidth=Wd)
        cv.grid_set(gold.grid)
        cv.vel_grid_set(vgrid)
        rv = cv.epoch_run(gold.k(e, "iq"), ep, with_vel=1)
        NBd = 2 * Wd + 2
        carr = cv.copy_out(capi.PTR_CARR_SCORES, np.float64, gold.C * NBd * 2).reshape(gold.C, NBd, 2)
        refc = gold.k(e, "carr_scores_win").reshape(gold.C, NBd, 2)
        d = (carr[..., 0] - refc[..., 0]) + 1j * (carr[..., 1] - refc[..., 1])
        assert np.max(np.abs(d)) / np.max(np.hypot(refc[..., 0], refc[..., 1])) < 5e-6
        assert np.max(np.abs(np.array(rv.z[4:8]) - gold.k(e, "zval")[4:8])) < 1e-9
        cv.close()
    assert n_race_free >= gold.epochs * gold.C // 3


# ---- a13: the reference's DORMANT score-weighted estimator, executed -----------------
# tests/golden/ref_weighted_n9.npz: oracle/_ref/ref_dpe_weighted launches BCM_PosMeasReduction <<<8,64>>> +
# BCM_ReduceAndPosMeas <<<1,8>>> (batchcorrmanifold.cu:816-1056, 1365-1510; launches commented out at
# :2547-2567) on the module's own buffers after every epoch (oracle/make_golden_ref.py --weighted).
def _weighted_gold():
    g = np.load(os.path.join(os.path.dirname(GOLD), "ref_weighted_n9.npz"))
    C, T, S, W, fs = int(g["C"]), int(g["T"]), int(g["S"]), int(g["W"]), float(g["fs"])
    out = []
    for e in range(int(g["epochs"])):
        k = lambda n: g["e%d_%s" % (e, n)]
        w = k("code_scores_win").reshape(C, 2 * W + 2, 2)
        ep = dict(prn=k("prn"), rc_start=k("rc_start"), ri_start=k("ri_start"), fc=k("fc"), fi=k("fi"),
                  cp_start=k("cp_start"), cp_ref=k("cp_ref"), rc_end=k("rc_end"), cp_end=k("cp_end"),
                  cp_ref_tow=k("cp_ref_tow"), rx_time=float(k("rx_time")[0]), center=k("x_kk1"), enu2ecef=k("enu2ecef"),
                  sat_states=k("sat_states").reshape(-1, 8), doppler_sign=1, S=S, fs=fs, time_dim=T)
        rec = dict(ep=ep, win=w[..., 0] + 1j * w[..., 1], tx_time=k("tx_time"), z=k("zval_weighted"),
                   parts=k("weighted_parts"), z_ml=k("zval"))
        if "e%d_zval_weighted_vel" % e in g.files:          # the velocity twins (golden regenerated in round 2)
            Wd = int(g["Wd"])
            cw = k("carr_scores_win").reshape(C, 2 * Wd + 2, 2)
            rec.update(carr_win=cw[..., 0] + 1j * cw[..., 1], z_vel=k("zval_weighted_vel"), vel_parts=k("weighted_vel_parts"))
        out.append(rec)
    return dict(C=C, T=T, S=S, W=W, fs=fs, grid=g["grid"], epochs=out, Wd=int(g["Wd"]), n_fft=int(g["n_fft"]),
                vel_dim=int(g["vel_dim"]))


def test_oracle_weighted_estimator_matches_the_reference_kernels():
    wg = _weighted_gold()
    S, W, C = wg["S"], wg["W"], wg["C"]
    assert len(wg["epochs"]) == 3
    for e in wg["epochs"]:
        ep = e["ep"]
        full = np.zeros((C, S), complex)
        full[:, S // 2 - W: S // 2 - W + 2 * W + 2] = e["win"]
        # the kernel as written: txTime form of the code index, per-time satellite state, <<<8,64>>> partition
        r = orc.pos_meas_reduction(full, wg["grid"], ep["center"], ep["enu2ecef"], ep["sat_states"], wg["T"], ep["fc"],
                                   e["tx_time"], ep["rx_time"], wg["fs"], S)
        assert np.max(np.abs(r["z"] - e["z"])) < 1e-7                           # metres
        assert np.max(np.abs(r["parts"] - e["parts"]) / np.abs(e["parts"])) < 1e-11
        # the form the product uses (the ML kernel's code-index expression, :1779-1791): same estimate to 2e-5 m
        # (txTime ~ 4e5 s carries 6e-11 s = 6e-5 chip of rounding into every bin of the reduction kernel)
        r2 = orc.pos_meas_weighted(full, wg["grid"], ep["center"], ep["enu2ecef"], ep["sat_states"], wg["T"], ep["fc"],
                                   ep["rc_end"], ep["cp_ref_tow"], ep["cp_end"], ep["cp_ref"], ep["rx_time"], wg["fs"], S,
                                   per_time_sat=True)
        assert np.max(np.abs(r2["z"] - e["z"])) < 1e-4
        assert abs(r2["sum_score"] / e["parts"][:, 4].sum() - 1.0) < 1e-5
        # and it is a different estimate from the arg-max the reference actually publishes
        assert np.max(np.abs(e["z"] - e["z_ml"][:4])) > 1e-3


def test_oracle_weighted_velocity_estimator_matches_the_reference_kernels():
    """BCM_VelMeasReduction <<<8,64>>> + BCM_ReduceAndVelMeas <<<1,8>>> (batchcorrmanifold.cu:1090-1347, 1525-1667; dormant,
    launches commented out at :2555-2566), executed by oracle/_ref/ref_dpe_weighted on the module's own CarrScores."""
    wg = _weighted_gold()
    C, Wd, n_fft = wg["C"], wg["Wd"], wg["n_fft"]
    vgrid, _ = H.synth.uniform_grid(wg["vel_dim"], 1.0)
    assert all("z_vel" in e for e in wg["epochs"])
    for e in wg["epochs"]:
        ep = e["ep"]
        full = np.zeros((C, n_fft), complex)
        full[:, n_fft // 2 - Wd: n_fft // 2 - Wd + 2 * Wd + 2] = e["carr_win"]
        r = orc.vel_meas_reduction(full, vgrid, ep["center"], ep["enu2ecef"], ep["sat_states"], wg["T"], ep["fi"],
                                   ep["doppler_sign"], wg["fs"], n_fft)
        assert np.max(np.abs(r["z"] - e["z_vel"])) < 1e-9                        # m/s
        assert np.max(np.abs(r["parts"] - e["vel_parts"]) / np.abs(e["vel_parts"][:, 4:5])) < 1e-11
        # and it is a different estimate from the arg-max the reference actually publishes
        assert np.max(np.abs(e["z_vel"] - e["z_ml"][4:8])) > 1e-3


@pytest.mark.gpu
def test_cuda_weighted_estimator_matches_the_reference_kernels(capi):
    wg = _weighted_gold()
    G = wg["grid"].shape[0]
    for e in wg["epochs"]:
        ctx = capi.Context(fs=wg["fs"], S=wg["S"], max_chan=wg["C"], G=G, time_dim=wg["T"], lag_halfwidth=wg["W"])
        ctx.grid_set(wg["grid"])
        ctx.epoch_set(e["ep"])
        ctx.code_scores_set(e["win"])                       # the reference's own CodeScores rows
        ctx.score_pos(capi.SCORE_LOOKUP, capi.SAT_PER_TIME)
        ctx.estimate(capi.EST_WEIGHTED)
        res = ctx.result_fetch()
        assert np.max(np.abs(np.array(res.z[:4]) - e["z"])) < 1e-4          # bar: 0.1 m / 0.2998 m
        assert abs(res.sum_score / e["parts"][:, 4].sum() - 1.0) < 1e-5
        ctx.close()
