"""GPU parity: libdpe_b200 (through the C ABI) against the CPU oracle on the same
seeded inputs.  Bars (BASELINE.json): chip indices and code-phase bins bit-exact,
per-candidate scores <= 1e-5 relative, position <= 0.1 m, clock <= 1 ns (0.2998 m).
"""
import numpy as np
import pytest

import helpers as H
from helpers import orc, synth

pytestmark = pytest.mark.gpu

SCORE_RTOL = 1e-5          # BASELINE.json: "per-candidate correlation scores within 1e-5 relative in fp32"
W = 32


def _ctx(capi, S, C, G, T, fs, flags=0, lpower=1, W_=W, **kw):
    return capi.Context(fs=fs, S=S, max_chan=C, G=G, time_dim=T, lpower=lpower, lag_halfwidth=W_,
                        flags=flags, **kw)


def _run_prepare(capi, iq, grid, ep, flags=0, lpower=1, W_=W):
    C = len(ep["prn"])
    ctx = _ctx(capi, ep["S"], C, grid.shape[0], ep["time_dim"], ep["fs"], flags, lpower, W_)
    ctx.grid_set(grid)
    ctx.block_stage(iq)
    ctx.epoch_set(ep)
    ctx.replica_prepare()
    ctx.correlogram()
    return ctx


def test_ca_table_matches_oracle(capi):
    ctx = _ctx(capi, 5000, 1, 16, 1, 2.5e6)
    tab = ctx.copy_out(capi.PTR_CA_TABLE, np.int8, 37 * 1024).reshape(37, 1024)
    ref = orc.ca_table()
    assert np.array_equal(tab[:, :1023], ref)


@pytest.mark.parametrize("fs,prns", [(2.5e6, synth.PRNS_8), (10.0e6, synth.PRNS_12)])
def test_prepare_chip_index_and_edges_bit_exact(capi, fs, prns):
    sc, iq, grid, ep = H.epoch_case(fs=fs, prns=prns, n=3)
    C, S = len(prns), ep["S"]
    ctx = _run_prepare(capi, iq, grid, ep, flags=capi.FLAG_KEEP_CHIP_IDX)
    chip = ctx.copy_out(capi.PTR_CHIP_IDX, np.int16, C * S).reshape(C, S)
    rs = ctx.copy_out(capi.PTR_REPLICA_SIGN, np.int8, C * S).reshape(C, S)
    idx_next, no_flip = ctx.channel_flags(C)
    t = orc.time_idcs(S, fs)
    ref_next = orc.nav_bit_boundary(ep["cp_start"], ep["cp_ref"], ep["rc_start"], ep["fc"], fs)
    assert np.array_equal(idx_next, ref_next)
    for c in range(C):
        ci, r_nf, _ = orc.code_replica(int(ep["prn"][c]), t, ep["fc"][c], ep["rc_start"][c], int(ref_next[c]), S)
        assert np.array_equal(chip[c].astype(np.int32), ci), "chip index mismatch on channel %d" % c
        assert np.array_equal(rs[c].astype(np.float64), r_nf)


def test_wiped_samples_match_oracle(capi):
    sc, iq, grid, ep = H.epoch_case(n=3)
    C, S = 8, ep["S"]
    ctx = _run_prepare(capi, iq, grid, ep)
    xw = ctx.copy_out(capi.PTR_XW, np.float32, C * S * 2).reshape(C, S, 2)
    x = orc.iq_to_complex(iq)
    t = orc.time_idcs(S, ep["fs"])
    for c in range(C):
        ref = x * orc.doppler_wipeoff(ep["fi"][c], ep["ri_start"][c], t)
        got = xw[c, :, 0].astype(np.float64) + 1j * xw[c, :, 1]
        # FP32 NCO on a phase range-reduced in FP64 (<= 6e-8 cycle = 3.8e-7 rad) + FP32 products
        assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)) < 1.0e-6


@pytest.mark.parametrize("fs,prns,W_", [(2.5e6, synth.PRNS_8, 32), (2.5e6, synth.PRNS_8, 5),
                                         (10.0e6, synth.PRNS_12, 32)])
def test_correlogram_window_matches_fft_oracle(capi, fs, prns, W_):
    sc, iq, grid, ep = H.epoch_case(fs=fs, prns=prns, n=3)
    C, S = len(prns), ep["S"]
    bcs = H.oracle_bcs(fs=fs, prns=prns)
    ctx = _run_prepare(capi, iq, grid, ep, W_=W_)
    NL = 2 * W_ + 2
    cs = ctx.copy_out(capi.PTR_CODE_SCORES, np.float64, C * NL * 2).reshape(C, NL, 2)
    got = cs[..., 0] + 1j * cs[..., 1]
    ref = bcs["code_scores"][:, S // 2 - W_: S // 2 - W_ + NL]
    _, no_flip = ctx.channel_flags(C)
    assert np.array_equal(no_flip.astype(bool), bcs["no_flip"])
    # error relative to the noise floor of the correlogram (off-peak magnitude), far below 1e-5 of the peak
    floor = np.median(np.abs(bcs["code_scores"]), axis=1)[:, None]
    assert np.max(np.abs(got - ref) / floor) < 2e-5
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 1e-6


@pytest.mark.parametrize("sat_mode", [0, 1])
def test_code_phase_bins_bit_exact(capi, sat_mode):
    sc, iq, grid, ep = H.epoch_case(n=9, center_offset=(3.0, -2.0, 4.0, 7.0))
    C, S, G = 8, ep["S"], grid.shape[0]
    ctx = _run_prepare(capi, iq, grid, ep)
    f, a = ctx.debug_bins(0, G, C, sat_mode=sat_mode)
    _, idxo, f_ref, _, valid = orc.pos_bins(grid, ep["center"], ep["enu2ecef"], ep["sat_states"], ep["time_dim"],
                                            ep["fc"], ep["rc_end"], ep["cp_ref_tow"], ep["cp_end"], ep["cp_ref"],
                                            ep["rx_time"], ep["fs"], S, per_time_sat=bool(sat_mode))
    assert valid.all()
    assert np.array_equal(f, f_ref)
    assert np.max(np.abs(a - (idxo - f_ref))) < 1e-9


@pytest.mark.parametrize("case", ["demo25", "c4shard", "wide"])
@pytest.mark.parametrize("sat_mode", [0, 1])
def test_fast_geometry_equals_the_reference_chain(capi, case, sat_mode):
    """The scoring kernels evaluate the code-phase index without norm(), sqrt and the division (centre-relative
    range, dpe_geom.cuh::code_index_fast) but with the SAME rounded tx as the reference's FP64 chain
    (batchcorrmanifold.cu:1779-1791), falling back to that chain next to a rounding boundary: floor index AND
    lerp weight must be bit-identical for every (candidate, PRN) pair -- checked on the whole demo grid, a c4
    shard (10 MHz, 12 PRNs, 2 m steps: many indices next to integers) and a grid 1 km wide (the Taylor bound)."""
    if case == "demo25":
        sc = H.scenario()
        grid, tg = synth.spread_grid(), 6.0 * synth.spread_axis()
        W = 16
    elif case == "c4shard":
        sc = H.scenario(10.0e6, synth.PRNS_12)
        grid, tg = synth.uniform_grid(51, (2.0, 2.0, 2.0, 2.0))
        grid = np.ascontiguousarray(grid[3 * 845651: 3 * 845651 + 400000])
        W = 16
    else:
        sc = H.scenario()
        grid, tg = synth.uniform_grid(15, (150.0, 150.0, 150.0, 150.0))
        W = 40
    n_diff = 0
    for b, off in ((0, (4.0, -3.0, 2.0, 5.0)), (3, (-41.7, 12.3, 77.7, -95.1))):
        center = sc.rx_state(sc.cfg.rx_time0 + (b + 1) * sc.cfg.T).copy()
        center[:4] += off
        ep = sc.epoch_inputs(b, center=center, time_grid=tg)
        G = grid.shape[0]
        ctx = capi.Context(fs=ep["fs"], S=ep["S"], max_chan=sc.C, G=G, time_dim=len(tg), lag_halfwidth=W)
        ctx.grid_set(grid)
        ctx.epoch_set(ep)
        f0, a0 = ctx.debug_bins(0, G, sc.C, sat_mode=sat_mode, exact=True)
        f1, a1 = ctx.debug_bins(0, G, sc.C, sat_mode=sat_mode)
        assert np.array_equal(f0, f1)
        assert np.array_equal(a0, a1)                                   # bit for bit, not "close"
        n_diff += int(np.unique(f0).size)
        ctx.close()
    assert n_diff > 2


@pytest.mark.parametrize("lpower", [1, 2, 3])
def test_lookup_scores_argmax_and_fix(capi, lpower):
    sc, iq, grid, ep = H.epoch_case(n=9, center_offset=(10.0, -5.0, 5.0, 12.0))
    G = grid.shape[0]
    bcs = H.oracle_bcs()
    ref = H.oracle_pos(bcs, grid, ep, lpower=lpower)
    ctx = _run_prepare(capi, iq, grid, ep, lpower=lpower)
    ctx.score_pos(capi.SCORE_LOOKUP, capi.SAT_MIDDLE)
    ctx.estimate(capi.EST_ARGMAX)
    res = ctx.result_fetch()
    scores = ctx.copy_out(capi.PTR_POS_SCORES, np.float64, G)
    assert np.max(np.abs(scores - ref["scores"]) / ref["scores"]) < SCORE_RTOL
    assert res.argmax == ref["argmax"]
    assert res.out_of_window == 0
    z = np.array(res.z[:4])
    assert np.max(np.abs(z[:3] - ref["z"][:3])) < 1e-6       # bar: 0.1 m
    assert abs(z[3] - ref["z"][3]) < 1e-6                    # bar: 0.2998 m (1 ns)
    zval = ctx.copy_out(capi.PTR_ZVAL, np.float64, 4)
    assert np.array_equal(zval, z)
    rval = ctx.copy_out(capi.PTR_RVAL, np.float64, 64).reshape(8, 8)
    assert np.array_equal(rval[:4], np.eye(8)[:4])
    # the fix recovers the truth: the grid is centred 10/-5/5/12 m off, 5/6 m spacing
    truth = sc.rx_state(ep["rx_time"])
    assert np.linalg.norm(z[:3] - truth[:3]) < 9.0


def test_weighted_estimate_matches_oracle(capi):
    sc, iq, grid, ep = H.epoch_case(n=9, center_offset=(3.0, 1.0, -2.0, 4.0))
    bcs = H.oracle_bcs()
    ref = H.oracle_pos(bcs, grid, ep, weighted=True, per_time=True)
    ctx = _run_prepare(capi, iq, grid, ep)
    ctx.score_pos(capi.SCORE_LOOKUP, capi.SAT_PER_TIME)
    ctx.estimate(capi.EST_WEIGHTED)
    res = ctx.result_fetch()
    z = np.array(res.z[:4])
    assert np.max(np.abs(z - ref["z"])) < 1e-4
    assert abs(res.sum_score - ref["sum_score"]) / ref["sum_score"] < 1e-7


@pytest.mark.parametrize("est,sat,score", [("EST_ARGMAX", "SAT_MIDDLE", "SCORE_LOOKUP"), ("EST_WEIGHTED", "SAT_PER_TIME", "SCORE_LOOKUP"),
                                           ("EST_ARGMAX", "SAT_MIDDLE", "SCORE_BRUTE")])
def test_fold_estimate_in_the_stage_api(capi, est, sat, score):
    """dpe_fold_estimate: the scoring kernel's last CTA writes the single-GPU estimate itself; the result block equals
    the one of the separate k_finalize launch bit for bit, dpe_estimate then launches nothing, and a different mode asked
    of dpe_estimate afterwards still gets its own launch."""
    est, sat, score = getattr(capi, est), getattr(capi, sat), getattr(capi, score)
    sc, iq, grid, ep = H.epoch_case(n=9, center_offset=(3.0, 1.0, -2.0, 4.0))
    out = []
    for fold in (False, True):
        ctx = _run_prepare(capi, iq, grid, ep, flags=capi.FLAG_BRUTE_TILES)
        if fold:
            ctx.fold_estimate(est)
        ctx.score_pos(score, sat)
        n0 = ctx.launch_count()
        ctx.estimate(est)
        out.append((ctx.result_fetch(), ctx.launch_count() - n0, ctx.copy_out(capi.PTR_ZVAL, np.float64, 4),
                    ctx.copy_out(capi.PTR_RVAL, np.float64, 64)))
        if fold:
            other = capi.EST_ARGMAX if est == capi.EST_WEIGHTED else capi.EST_WEIGHTED
            ctx.estimate(other)
            assert ctx.launch_count() - n0 == 1
            ctx.fold_estimate(-1)
        ctx.close()
    (r0, l0, z0, rv0), (r1, l1, z1, rv1) = out
    assert (l0, l1) == (1, 0)
    assert list(r0.z) == list(r1.z) and r0.argmax == r1.argmax and r0.max_score == r1.max_score
    assert r0.sum_score == r1.sum_score and r0.out_of_window == r1.out_of_window
    assert np.array_equal(z0, z1) and np.array_equal(rv0, rv1)


@pytest.mark.parametrize("fs,prns,n", [(2.5e6, synth.PRNS_8, 9), (2.5e6, synth.PRNS_12, 7),
                                        (10.0e6, synth.PRNS_12, 7)])
def test_brute_force_scores_match_oracle(capi, fs, prns, n):
    sp = (5.0, 5.0, 5.0, 6.0) if fs < 5e6 else (2.0, 2.0, 2.0, 2.0)
    sc, iq, grid, ep = H.epoch_case(fs=fs, prns=prns, n=n, spacing=sp, center_offset=(4.0, -3.0, 2.0, 5.0))
    G = grid.shape[0]
    bcs = H.oracle_bcs(fs=fs, prns=prns)
    ref = H.oracle_pos(bcs, grid, ep)
    ctx = _run_prepare(capi, iq, grid, ep, flags=capi.FLAG_BRUTE_TILES)
    ctx.score_pos(capi.SCORE_BRUTE, capi.SAT_MIDDLE)
    ctx.estimate(capi.EST_ARGMAX)
    res = ctx.result_fetch()
    scores = ctx.copy_out(capi.PTR_POS_SCORES, np.float64, G)
    rel = np.abs(scores - ref["scores"]) / ref["scores"]
    assert rel.max() < SCORE_RTOL, "brute-force score error %.3g" % rel.max()
    assert res.argmax == ref["argmax"]
    assert np.max(np.abs(np.array(res.z[:4]) - ref["z"])) < 1e-6
    # and the two device paths agree with each other
    ctx.score_pos(capi.SCORE_LOOKUP, capi.SAT_MIDDLE)
    lookup = ctx.copy_out(capi.PTR_POS_SCORES, np.float64, G)
    assert np.max(np.abs(scores - lookup) / lookup) < SCORE_RTOL


def test_brute_force_ragged_groups_and_flipped_channels(capi):
    """Grid sizes that leave partial groups (G not a multiple of 16) and a block whose
    nav-bit edge falls inside (flipped replica on some channel)."""
    found = None
    for block in range(6):
        bcs = H.oracle_bcs(block=block)
        if (~bcs["no_flip"]).any():
            found = block
            break
    assert found is not None, "no flipped channel in the first blocks of the scenario"
    sc, iq, grid, ep = H.epoch_case(block=found, n=5)
    grid = grid[:601]                                  # ragged: 601 = 37*16 + 9
    ref = H.oracle_pos(H.oracle_bcs(block=found), grid, ep)
    ctx = _run_prepare(capi, iq, grid, ep, flags=capi.FLAG_BRUTE_TILES)
    ctx.score_pos(capi.SCORE_BRUTE, capi.SAT_MIDDLE)
    scores = ctx.copy_out(capi.PTR_POS_SCORES, np.float64, grid.shape[0])
    assert np.max(np.abs(scores - ref["scores"]) / ref["scores"]) < SCORE_RTOL


def test_brute_presort_on_a_second_stream_changes_nothing(capi):
    """dpe_brute_presort (the pair sort run ahead of time, beside the sample pre-pass): same bits as the
    in-line sort; a new epoch_set invalidates it; a presort for the other sat_mode is not used."""
    import torch
    sc, iq, grid, ep = H.epoch_case(n=7, center_offset=(3.0, 2.0, -1.0, 4.0))
    G = grid.shape[0]
    ctx = _ctx(capi, ep["S"], sc.C, G, ep["time_dim"], ep["fs"], flags=capi.FLAG_BRUTE_TILES)
    ctx.grid_set(grid)
    aux = torch.cuda.Stream()

    def run(presort_mode):
        e = capi.make_epoch(ep)
        ctx.block_stage(iq)
        ctx.epoch_set(e, np.ascontiguousarray(ep["sat_states"]))
        if presort_mode is not None:
            ctx.brute_presort(presort_mode, aux.cuda_stream)
        ctx.replica_prepare()
        ctx.correlogram()
        ctx.score_pos(capi.SCORE_BRUTE, capi.SAT_MIDDLE)
        ctx.estimate(capi.EST_ARGMAX)
        r = ctx.result_fetch()
        return r, ctx.copy_out(capi.PTR_POS_SCORES, np.float64, G)

    r0, s0 = run(None)
    r1, s1 = run(capi.SAT_MIDDLE)
    r2, s2 = run(capi.SAT_PER_TIME)            # wrong mode presorted: score_pos sorts again for SAT_MIDDLE
    assert np.array_equal(s0, s1) and np.array_equal(s0, s2)
    assert r0.argmax == r1.argmax == r2.argmax
    with pytest.raises(capi.DpeError):
        lookup_only = _ctx(capi, ep["S"], sc.C, G, ep["time_dim"], ep["fs"])
        try:
            lookup_only.brute_presort(capi.SAT_MIDDLE)
        finally:
            lookup_only.close()
    ctx.close()


def test_out_of_window_candidates_are_counted_not_scored(capi):
    sc, iq, grid, ep = H.epoch_case(n=3, spacing=(400.0, 400.0, 400.0, 400.0))
    G, C = grid.shape[0], 8
    ctx = _run_prepare(capi, iq, grid, ep, W_=2)        # +-2 samples = +-240 m of range
    ctx.score_pos(capi.SCORE_LOOKUP, capi.SAT_MIDDLE)
    ctx.estimate(capi.EST_ARGMAX)
    res = ctx.result_fetch()
    f, _ = ctx.debug_bins(0, G, C)
    lag = f - (np.arange(C) * ep["S"])[None, :] - ep["S"] // 2
    n_out = int(((lag < -2) | (lag > 2)).sum())
    assert n_out > 0 and res.out_of_window == n_out


def test_call_order_and_argument_errors(capi):
    ctx = _ctx(capi, 5000, 2, 16, 1, 2.5e6)
    with pytest.raises(capi.DpeError) as e:
        ctx.replica_prepare()
    assert e.value.code == capi.DPE_ESTATE
    with pytest.raises(capi.DpeError) as e:
        ctx.grid_set(np.zeros((15, 4)))
    assert e.value.code == capi.DPE_EINVAL
    with pytest.raises(capi.DpeError) as e:
        ctx.score_pos(capi.SCORE_BRUTE)
    assert e.value.code == capi.DPE_ESTATE
    with pytest.raises(capi.DpeError):
        capi.Context(fs=2.5e6, S=5001, max_chan=2, G=16)      # odd block length


def test_epoch_run_host_buffers_and_determinism(capi):
    sc, iq, grid, ep = H.epoch_case(n=7)
    ctx = _ctx(capi, ep["S"], 8, grid.shape[0], ep["time_dim"], ep["fs"], flags=capi.FLAG_BRUTE_TILES)
    ctx.grid_set(grid)
    r1 = ctx.epoch_run(iq, ep, score_mode=capi.SCORE_LOOKUP)
    r2 = ctx.epoch_run(iq, ep, score_mode=capi.SCORE_LOOKUP)
    r3 = ctx.epoch_run(iq, ep, score_mode=capi.SCORE_BRUTE)
    assert r1.argmax == r2.argmax == r3.argmax
    assert r1.max_score == r2.max_score and r1.sum_score == r2.sum_score     # bit-reproducible
    assert abs(r3.max_score - r1.max_score) / r1.max_score < SCORE_RTOL
    assert ctx.launch_count() > 0


def test_chunk_totals_are_bit_reproducible_and_two_contexts_agree(capi):
    """The chunk partials of the correlogram and of the carrier spectrum meet in fixed-point integer accumulators
    (order-independent atomics, cleared by the CTA that reads them): CodeScores, CarrScores and both fixes are
    bit-identical from epoch to epoch, between the kernel-by-kernel and the CUDA-graph call, and between two contexts."""
    fs, prns = 2.5e6, synth.PRNS_8
    sc = H.scenario(fs, prns)
    grid, tg = synth.uniform_grid(5, (5.0, 5.0, 5.0, 6.0))
    vgrid, _ = synth.uniform_grid(7, (0.5, 0.5, 0.5, 0.25))
    eps = [sc.epoch_inputs(b, time_grid=tg) for b in (0, 1)]
    blks = [sc.block(b) for b in (0, 1)]
    C, S, W_, Wd = len(prns), eps[0]["S"], 16, 64
    seen = []
    for _ in range(2):
        ctx = capi.Context(fs=fs, S=S, max_chan=C, G=grid.shape[0], time_dim=len(tg), lag_halfwidth=W_,
                           Gv=vgrid.shape[0], dopp_halfwidth=Wd)
        ctx.grid_set(grid)
        ctx.vel_grid_set(vgrid)
        for k in range(6):
            blk, ep = blks[k % 2], eps[k % 2]                      # alternate two blocks: a stale accumulator would show
            r = ctx.epoch_run(blk, ep, with_vel=1) if k % 3 else ctx.epoch_run_dist(blk, ep, with_vel=1)
            cs = ctx.copy_out(capi.PTR_CODE_SCORES, np.float64, C * (2 * W_ + 2) * 2)
            carr = ctx.copy_out(capi.PTR_CARR_SCORES, np.float64, C * (2 * Wd + 2) * 2)
            seen.append((k % 2, cs.copy(), carr.copy(), tuple(r.z), r.max_score, r.vel_max_score))
        ctx.close()
    for b in (0, 1):
        same = [x for x in seen if x[0] == b]
        assert len(same) == 6
        for x in same[1:]:
            assert np.array_equal(x[1], same[0][1]) and np.array_equal(x[2], same[0][2])
            assert x[3:] == same[0][3:]
    assert not np.array_equal(seen[0][1], seen[1][1])              # the two blocks do differ


# ---- velocity / clock-drift manifold ------------------------------------------
@pytest.mark.parametrize("fs,prns", [(2.5e6, synth.PRNS_8), (10.0e6, synth.PRNS_12)])
def test_velocity_manifold_matches_oracle(capi, fs, prns):
    sc = H.scenario(fs, prns)
    grid, tg = synth.uniform_grid(5, (5.0, 5.0, 5.0, 6.0))
    vgrid, _ = synth.uniform_grid(9, (0.5, 0.5, 0.5, 0.25))
    center = sc.rx_state(sc.cfg.rx_time0 + sc.cfg.T).copy()
    center[4:] += (1.0, -0.5, 0.5, 0.3)                       # velocity / drift error of the prediction
    ep = sc.epoch_inputs(0, center=center, time_grid=tg)
    iq = sc.block(0)
    C, S, Wd = len(prns), ep["S"], 64
    bcs = orc.batch_corr_scores(iq, ep["prn"], ep["rc_start"], ep["ri_start"], ep["fc"], ep["fi"], ep["cp_start"],
                                ep["cp_ref"], ep["fs"], want_carrier=True)
    n_fft = bcs["n_fft"]
    assert n_fft == 8 * (1 << int(np.ceil(np.log2(S))))
    ref = orc.vel_meas_ml(bcs["carr_scores"], vgrid, ep["center"], ep["enu2ecef"], ep["sat_states"], ep["time_dim"],
                          ep["fi"], ep["doppler_sign"], ep["fs"], n_fft)
    assert ref["valid"].all()
    ctx = capi.Context(fs=fs, S=S, max_chan=C, G=grid.shape[0], time_dim=len(tg), lag_halfwidth=16,
                       Gv=vgrid.shape[0], dopp_halfwidth=Wd)
    ctx.grid_set(grid)
    ctx.vel_grid_set(vgrid)
    res = ctx.epoch_run(iq, ep, with_vel=1)
    NBd = 2 * Wd + 2
    carr = ctx.copy_out(capi.PTR_CARR_SCORES, np.float64, C * NBd * 2).reshape(C, NBd, 2)
    got = carr[..., 0] + 1j * carr[..., 1]
    want = bcs["carr_scores"][:, n_fft // 2 - Wd: n_fft // 2 - Wd + NBd]
    assert np.max(np.abs(got - want)) / np.max(np.abs(want)) < 5e-6
    vs = ctx.copy_out(capi.PTR_VEL_SCORES, np.float64, vgrid.shape[0])
    assert np.max(np.abs(vs - ref["scores"]) / ref["scores"]) < SCORE_RTOL
    assert res.vel_argmax == ref["argmax"] and res.vel_out_of_window == 0
    assert np.max(np.abs(np.array(res.z[4:8]) - ref["z"])) < 1e-9
    rval = ctx.copy_out(capi.PTR_RVAL, np.float64, 64).reshape(8, 8)
    assert np.array_equal(rval, np.eye(8))
    # the manifold pulls the velocity prediction back towards the (static) truth
    truth = sc.rx_state(ep["rx_time"])
    assert np.linalg.norm(np.array(res.z[4:7]) - truth[4:7]) < np.linalg.norm(center[4:7] - truth[4:7])
    ctx.close()


# ---- brute-force velocity manifold (SURVEY.md 8 a', last sentence: the blended-carrier identity) ----
@pytest.mark.parametrize("fs,prns", [(2.5e6, synth.PRNS_8), (10.0e6, synth.PRNS_12)])
def test_brute_force_velocity_manifold_matches_oracle_and_lookup(capi, fs, prns):
    """Every (velocity candidate, PRN) pair correlates the whole block against its own blended carrier
    (batchcorrmanifold.cu:1950-1958 by linearity): scores <= 1e-5 of the oracle's and of the lookup kernel's,
    identical arg-max and fix; the one-call and the CUDA-graph (submit / collect) paths agree bit for bit."""
    sc = H.scenario(fs, prns)
    grid, tg = synth.uniform_grid(5, (5.0, 5.0, 5.0, 6.0))
    vgrid, _ = synth.uniform_grid(9, (0.5, 0.5, 0.5, 0.25))
    center = sc.rx_state(sc.cfg.rx_time0 + sc.cfg.T).copy()
    center[4:] += (1.0, -0.5, 0.5, 0.3)
    ep = sc.epoch_inputs(0, center=center, time_grid=tg)
    iq = sc.block(0)
    C, S, Wd = len(prns), ep["S"], 64
    bcs = orc.batch_corr_scores(iq, ep["prn"], ep["rc_start"], ep["ri_start"], ep["fc"], ep["fi"], ep["cp_start"],
                                ep["cp_ref"], ep["fs"], want_carrier=True)
    ref = orc.vel_meas_ml(bcs["carr_scores"], vgrid, ep["center"], ep["enu2ecef"], ep["sat_states"], ep["time_dim"],
                          ep["fi"], ep["doppler_sign"], ep["fs"], bcs["n_fft"])
    ctx = capi.Context(fs=fs, S=S, max_chan=C, G=grid.shape[0], time_dim=len(tg), lag_halfwidth=16,
                       Gv=vgrid.shape[0], dopp_halfwidth=Wd, flags=capi.FLAG_BRUTE_VEL)
    ctx.grid_set(grid)
    ctx.vel_grid_set(vgrid)
    r_look = ctx.epoch_run(iq, ep, with_vel=1)
    vs_look = ctx.copy_out(capi.PTR_VEL_SCORES, np.float64, vgrid.shape[0])
    r_brute = ctx.epoch_run(iq, ep, with_vel=2)
    vs = ctx.copy_out(capi.PTR_VEL_SCORES, np.float64, vgrid.shape[0])
    assert np.max(np.abs(vs - ref["scores"]) / ref["scores"]) < SCORE_RTOL
    assert np.max(np.abs(vs - vs_look) / vs_look) < SCORE_RTOL
    assert r_brute.vel_argmax == ref["argmax"] == r_look.vel_argmax and r_brute.vel_out_of_window == 0
    assert np.max(np.abs(np.array(r_brute.z[4:8]) - ref["z"])) < 1e-9
    assert np.array_equal(np.array(r_brute.z[:4]), np.array(r_look.z[:4]))      # the position fix is untouched
    rval = ctx.copy_out(capi.PTR_RVAL, np.float64, 64).reshape(8, 8)
    assert np.array_equal(rval, np.eye(8))
    r_graph = ctx.epoch_run_dist(iq, ep, with_vel=2)
    vs_graph = ctx.copy_out(capi.PTR_VEL_SCORES, np.float64, vgrid.shape[0])
    assert np.array_equal(vs_graph, vs) and r_graph.vel_argmax == r_brute.vel_argmax
    assert r_graph.vel_max_score == r_brute.vel_max_score
    ctx.close()


def test_brute_force_velocity_wide_grid_ragged_buckets_and_window_edge(capi):
    """13^4 velocity candidates spread over many Doppler bins (ragged (PRN, bin) buckets, several slots per
    bucket) with a Doppler window so narrow that part of the grid falls outside: those pairs are counted, not
    scored, exactly as the lookup kernel counts them."""
    fs, prns = 2.5e6, synth.PRNS_8
    sc = H.scenario(fs, prns)
    grid, tg = synth.uniform_grid(3, (5.0, 5.0, 5.0, 6.0))
    vgrid, _ = synth.uniform_grid(13, (1.1, 0.9, 1.3, 0.7))
    center = sc.rx_state(sc.cfg.rx_time0 + sc.cfg.T).copy()
    center[4:] += (2.0, -1.5, 0.7, -0.9)
    ep = sc.epoch_inputs(0, center=center, time_grid=tg)
    iq = sc.block(0)
    C, S, Wd = len(prns), ep["S"], 6
    ctx = capi.Context(fs=fs, S=S, max_chan=C, G=grid.shape[0], time_dim=len(tg), lag_halfwidth=16,
                       Gv=vgrid.shape[0], dopp_halfwidth=Wd, flags=capi.FLAG_BRUTE_VEL)
    ctx.grid_set(grid)
    ctx.vel_grid_set(vgrid)
    r_look = ctx.epoch_run(iq, ep, with_vel=1)
    vs_look = ctx.copy_out(capi.PTR_VEL_SCORES, np.float64, vgrid.shape[0])
    r_brute = ctx.epoch_run(iq, ep, with_vel=2)
    vs = ctx.copy_out(capi.PTR_VEL_SCORES, np.float64, vgrid.shape[0])
    assert 0 < r_look.vel_out_of_window < C * vgrid.shape[0]
    assert r_brute.vel_out_of_window == r_look.vel_out_of_window
    ok = vs_look > 0
    assert np.array_equal(ok, vs > 0)
    assert np.max(np.abs(vs[ok] - vs_look[ok]) / vs_look[ok]) < SCORE_RTOL
    assert r_brute.vel_argmax == r_look.vel_argmax
    assert np.array_equal(np.array(r_brute.z[4:8]), np.array(r_look.z[4:8]))
    bcs = orc.batch_corr_scores(iq, ep["prn"], ep["rc_start"], ep["ri_start"], ep["fc"], ep["fi"], ep["cp_start"],
                                ep["cp_ref"], ep["fs"], want_carrier=True)
    ref = orc.vel_meas_ml(bcs["carr_scores"], vgrid, ep["center"], ep["enu2ecef"], ep["sat_states"], ep["time_dim"],
                          ep["fi"], ep["doppler_sign"], ep["fs"], bcs["n_fft"])
    n_fft = bcs["n_fft"]
    l = ref["f_idx"] - n_fft * np.arange(C)[None, :] - n_fft // 2 + Wd
    pair_in = ref["valid"] & (l >= 0) & (l + 1 < 2 * Wd + 2)
    assert r_look.vel_out_of_window == int((~pair_in).sum())
    inside = pair_in.all(axis=1)                                   # candidates with every PRN inside the narrow window
    assert inside.sum() > vgrid.shape[0] // 20
    assert np.max(np.abs(vs[inside] - ref["scores"][inside]) / ref["scores"][inside]) < SCORE_RTOL
    ctx.close()


def test_brute_force_velocity_needs_its_flag(capi):
    ctx = capi.Context(fs=2.5e6, S=5000, max_chan=2, G=16, time_dim=1, lag_halfwidth=16, Gv=16, dopp_halfwidth=8)
    ctx.grid_set(np.zeros((16, 4)))
    ctx.vel_grid_set(np.zeros((16, 4)))
    with pytest.raises(capi.DpeError) as e:
        ctx.score_vel_brute()
    assert e.value.code == capi.DPE_ESTATE
    ctx.close()


@pytest.mark.parametrize("fs,prns", [(2.5e6, synth.PRNS_8), (10.0e6, synth.PRNS_12)])
def test_weighted_velocity_estimate_matches_oracle(capi, fs, prns):
    """DPE_EST_WEIGHTED on the velocity manifold (the reference's dormant BCM_VelMeasReduction + BCM_ReduceAndVelMeas;
    the oracle's restatement is pinned to those kernels in test_golden_ref.py): zVal[4:8] = sum s v / sum s."""
    sc = H.scenario(fs, prns)
    grid, tg = synth.uniform_grid(3, (5.0, 5.0, 5.0, 6.0))
    vgrid, _ = synth.uniform_grid(9, (0.5, 0.5, 0.5, 0.25))
    center = sc.rx_state(sc.cfg.rx_time0 + sc.cfg.T).copy()
    center[4:] += (1.0, -0.5, 0.5, 0.3)
    ep = sc.epoch_inputs(0, center=center, time_grid=tg)
    iq = sc.block(0)
    bcs = orc.batch_corr_scores(iq, ep["prn"], ep["rc_start"], ep["ri_start"], ep["fc"], ep["fi"], ep["cp_start"],
                                ep["cp_ref"], ep["fs"], want_carrier=True)
    ref = orc.vel_meas_reduction(bcs["carr_scores"], vgrid, ep["center"], ep["enu2ecef"], ep["sat_states"],
                                 ep["time_dim"], ep["fi"], ep["doppler_sign"], ep["fs"], bcs["n_fft"])
    ctx = capi.Context(fs=fs, S=ep["S"], max_chan=len(prns), G=grid.shape[0], time_dim=len(tg), lag_halfwidth=16,
                       Gv=vgrid.shape[0], dopp_halfwidth=64)
    ctx.grid_set(grid)
    ctx.vel_grid_set(vgrid)
    r_ml = ctx.epoch_run(iq, ep, with_vel=1)
    r_w = ctx.epoch_run(iq, ep, with_vel=3)
    vs = ctx.copy_out(capi.PTR_VEL_SCORES, np.float64, vgrid.shape[0])
    assert np.max(np.abs(vs - ref["scores"]) / ref["scores"]) < SCORE_RTOL
    assert np.max(np.abs(np.array(r_w.z[4:8]) - ref["z"])) < 1e-6           # m/s: FP32 carrier spectrum under an FP64 mean
    assert r_w.vel_argmax == r_ml.vel_argmax and r_w.vel_max_score == r_ml.vel_max_score
    assert np.array_equal(np.array(r_w.z[:4]), np.array(r_ml.z[:4]))
    assert np.max(np.abs(np.array(r_w.z[4:8]) - np.array(r_ml.z[4:8]))) > 1e-3     # a different estimate from the arg-max
    # stage by stage: the same through dpe_score_vel_est
    ctx.score_vel_est(capi.EST_WEIGHTED)
    assert list(ctx.result_fetch().z[4:8]) == list(r_w.z[4:8])
    ctx.score_vel_est(capi.EST_ARGMAX)
    assert list(ctx.result_fetch().z[4:8]) == list(r_ml.z[4:8])
    ctx.close()


def test_velocity_needs_its_grid(capi):
    ctx = _ctx(capi, 5000, 2, 16, 1, 2.5e6)
    with pytest.raises(capi.DpeError) as e:
        ctx.vel_grid_set(np.zeros((16, 4)))
    assert e.value.code == capi.DPE_ESTATE
