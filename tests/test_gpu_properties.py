"""Size-independent properties of the CUDA path at BASELINE.json's full sizes (where the oracle is
too slow to be the checker) and at the edges: the two scoring formulations agree, sharding is
exact, the path is linear in the samples, channel order does not matter, minimum sizes work."""
import numpy as np
import pytest

import helpers as H
from helpers import orc, synth

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _demo_case():
    sc = H.scenario()
    grid = synth.spread_grid()                                   # 25^4 = 390625 candidates (demo config)
    tg = 6.0 * synth.spread_axis()
    center = sc.rx_state(sc.cfg.rx_time0 + sc.cfg.T).copy()
    center[:4] += (4.0, -3.0, 2.0, 5.0)
    return sc, grid, tg, sc.epoch_inputs(0, center=center, time_grid=tg), sc.block(0)


def _scores(capi, ctx, iq, ep, mode, G, est=0):
    res = ctx.epoch_run(iq, ep, score_mode=mode, est_mode=est)
    return res, ctx.copy_out(capi.PTR_POS_SCORES, np.float64, G)


def test_full_demo_grid_brute_force_equals_lookup(capi):
    sc, grid, tg, ep, iq = _demo_case()
    G = grid.shape[0]
    ctx = capi.Context(fs=ep["fs"], S=ep["S"], max_chan=sc.C, G=G, time_dim=len(tg), lag_halfwidth=16,
                       flags=capi.FLAG_BRUTE_TILES)
    ctx.grid_set(grid)
    r_l, s_l = _scores(capi, ctx, iq, ep, capi.SCORE_LOOKUP, G)
    r_b, s_b = _scores(capi, ctx, iq, ep, capi.SCORE_BRUTE, G)
    assert ctx.brute_pairs() == G * sc.C and r_b.out_of_window == 0
    assert np.max(np.abs(s_b - s_l) / s_l) < RTOL
    assert r_b.argmax == r_l.argmax == int(np.argmax(s_l))       # first maximum
    assert np.array_equal(np.array(r_b.z[:4]), np.array(r_l.z[:4]))
    # spot-check 2000 random candidates of the full grid against the oracle
    idx = np.random.default_rng(7).choice(G, 2000, replace=False)
    ref = H.oracle_pos(H.oracle_bcs(), grid[idx], ep)
    assert np.max(np.abs(s_b[idx] - ref["scores"]) / ref["scores"]) < RTOL
    # weighted estimate, per-time satellite state, both formulations
    w_l, _ = _scores(capi, ctx, iq, ep, capi.SCORE_LOOKUP, G, est=capi.EST_WEIGHTED)
    w_b, _ = _scores(capi, ctx, iq, ep, capi.SCORE_BRUTE, G, est=capi.EST_WEIGHTED)
    assert np.max(np.abs(np.array(w_b.z[:4]) - np.array(w_l.z[:4]))) < 1e-3
    assert abs(w_b.sum_score - w_l.sum_score) / w_l.sum_score < 1e-6
    ctx.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_sharded_contexts_reproduce_the_single_context_result(capi, mode):
    """3 contexts holding contiguous shards (what 3 ranks would hold) + dpe_estimate on the gathered
    partials == one context holding the whole grid: bit for bit on the lookup path; on the
    brute-force path to FP32 summation order (a slot that a CTA boundary of k_brute's tile-granular work
    split cuts is summed in parts, and which pairs it holds depends on the shard), with the same arg-max."""
    import dpe_pkg
    import torch
    sharding = dpe_pkg.submodule("sharding")
    sc, iq, grid, ep = H.epoch_case(n=9, center_offset=(6.0, -4.0, 3.0, 7.0))
    G, C, T = grid.shape[0], sc.C, ep["time_dim"]
    full = capi.Context(fs=ep["fs"], S=ep["S"], max_chan=C, G=G, time_dim=T, lag_halfwidth=16, flags=capi.FLAG_BRUTE_TILES)
    full.grid_set(grid)
    for est in (capi.EST_ARGMAX, capi.EST_WEIGHTED):
        rf = full.epoch_run(iq, ep, score_mode=mode, est_mode=est)
        sf = full.copy_out(capi.PTR_POS_SCORES, np.float64, G)
        parts, scores = [], []
        world = 3
        ctxs = []
        for r in range(world):
            lo, hi = sharding.shard_range(G, world, r)
            c = capi.Context(fs=ep["fs"], S=ep["S"], max_chan=C, G=hi - lo, time_dim=T, lag_halfwidth=16,
                             flags=capi.FLAG_BRUTE_TILES, grid_offset=lo, G_total=G)
            c.grid_set(grid[lo:hi])
            c.epoch_run(iq, ep, score_mode=mode, est_mode=est)
            parts.append(c.copy_out(capi.PTR_PARTIAL, np.float64, capi.DPE_PARTIAL_LEN))
            scores.append(c.copy_out(capi.PTR_POS_SCORES, np.float64, hi - lo))
            ctxs.append(c)
        if mode == 0:
            assert np.array_equal(np.concatenate(scores), sf)       # same kernels, same inputs: identical bits
        else:
            assert np.max(np.abs(np.concatenate(scores) - sf) / sf) < 2e-6
        gathered = torch.from_numpy(np.concatenate(parts)).cuda()
        ctxs[0].estimate(est, gathered, world)
        rs = ctxs[0].result_fetch()
        assert rs.argmax == rf.argmax and abs(rs.max_score - rf.max_score) <= (0 if mode == 0 else 2e-6 * rf.max_score)
        assert np.max(np.abs(np.array(rs.z[:4]) - np.array(rf.z[:4]))) < 1e-6
        host = sharding.combine_partials(np.stack(parts), est)       # host mirror of k_finalize
        assert host["argmax"] == rf.argmax and np.max(np.abs(host["z"] - np.array(rf.z[:4]))) < 1e-6
        for c in ctxs:
            c.close()
    full.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_linearity_in_the_samples(capi, mode):
    """Doubling the int16 samples (no clipping) doubles every score exactly: all products and sums
    scale by a power of two."""
    sc, iq, grid, ep = H.epoch_case(n=7)
    assert np.abs(iq).max() < 16000
    G = grid.shape[0]
    ctx = capi.Context(fs=ep["fs"], S=ep["S"], max_chan=sc.C, G=G, time_dim=ep["time_dim"], lag_halfwidth=16,
                       flags=capi.FLAG_BRUTE_TILES)
    ctx.grid_set(grid)
    _, s1 = _scores(capi, ctx, iq, ep, mode, G)
    _, s2 = _scores(capi, ctx, (2 * iq).astype(np.int16), ep, mode, G)
    assert np.array_equal(s2, 2.0 * s1)
    ctx.close()


def test_channel_order_does_not_matter(capi):
    sc, iq, grid, ep = H.epoch_case(n=7, center_offset=(3.0, 2.0, -1.0, 4.0))
    G, C, T = grid.shape[0], sc.C, ep["time_dim"]
    perm = np.array([5, 2, 7, 0, 3, 6, 1, 4])
    ep2 = dict(ep)
    for k in ("prn", "rc_start", "ri_start", "fc", "fi", "cp_start", "cp_ref", "rc_end", "cp_end", "cp_ref_tow"):
        ep2[k] = np.asarray(ep[k])[perm]
    ep2["sat_states"] = ep["sat_states"].reshape(C, T, 8)[perm].reshape(C * T, 8)
    ctx = capi.Context(fs=ep["fs"], S=ep["S"], max_chan=C, G=G, time_dim=T, lag_halfwidth=16, flags=capi.FLAG_BRUTE_TILES)
    ctx.grid_set(grid)
    for mode in (capi.SCORE_LOOKUP, capi.SCORE_BRUTE):
        r1, s1 = _scores(capi, ctx, iq, ep, mode, G)
        r2, s2 = _scores(capi, ctx, iq, ep2, mode, G)
        # not bit-identical: the reference adds the row offset S*chan BEFORE the floor (batchcorrmanifold.cu:1797),
        # so the lerp fraction depends on the channel slot at the 1e-10 level
        # (and, in the brute-force path, its FP32 rounding and summation order)
        assert np.max(np.abs(s1 - s2) / s1) < (1e-7 if mode == capi.SCORE_LOOKUP else 2e-6) and r1.argmax == r2.argmax
    ctx.close()


@pytest.mark.parametrize("n_cand", [40, 700, 4700, 5100, 6000, 6561])
def test_brute_force_split_tail_slots(capi, n_cand):
    """k_brute cuts the tile sequence of all slots into equal shares per CTA, so slots straddling a CTA
    boundary are summed in parts (more parts the fewer slots there are): same scores as the lookup path
    and as the oracle, identical bits from run to run (the parts are added in a fixed order)."""
    sc, iq, grid, ep = H.epoch_case(n=9, center_offset=(6.0, -4.0, 3.0, 7.0))
    g = np.ascontiguousarray(grid[:n_cand])
    ctx = capi.Context(fs=ep["fs"], S=ep["S"], max_chan=sc.C, G=n_cand, time_dim=ep["time_dim"], lag_halfwidth=16,
                       flags=capi.FLAG_BRUTE_TILES)
    ctx.grid_set(g)
    r_l, s_l = _scores(capi, ctx, iq, ep, capi.SCORE_LOOKUP, n_cand)
    r_b, s_b = _scores(capi, ctx, iq, ep, capi.SCORE_BRUTE, n_cand)
    r_b2, s_b2 = _scores(capi, ctx, iq, ep, capi.SCORE_BRUTE, n_cand)
    assert np.array_equal(s_b, s_b2) and r_b.argmax == r_b2.argmax
    assert np.max(np.abs(s_b - s_l) / s_l) < RTOL and r_b.argmax == r_l.argmax
    idx = np.random.default_rng(11).choice(n_cand, min(500, n_cand), replace=False)
    ref = H.oracle_pos(H.oracle_bcs(), g[idx], ep)
    assert np.max(np.abs(s_b[idx] - ref["scores"]) / ref["scores"]) < RTOL
    ctx.close()


def test_minimum_sizes_one_candidate_one_channel(capi):
    sc, iq, grid, ep = H.epoch_case(n=3)
    one = dict(ep)
    for k in ("prn", "rc_start", "ri_start", "fc", "fi", "cp_start", "cp_ref", "rc_end", "cp_end", "cp_ref_tow"):
        one[k] = np.asarray(ep[k])[:1]
    T = ep["time_dim"]
    one["sat_states"] = ep["sat_states"][:T]
    g1 = grid[40:41]                                               # the centre candidate
    ctx = capi.Context(fs=ep["fs"], S=ep["S"], max_chan=1, G=1, time_dim=T, lag_halfwidth=1, flags=capi.FLAG_BRUTE_TILES)
    ctx.grid_set(g1)
    bcs = orc.batch_corr_scores(iq, one["prn"], one["rc_start"], one["ri_start"], one["fc"], one["fi"], one["cp_start"],
                                one["cp_ref"], one["fs"])
    ref = H.oracle_pos(bcs, g1, one)
    for mode in (capi.SCORE_LOOKUP, capi.SCORE_BRUTE):
        r, s = _scores(capi, ctx, iq, one, mode, 1)
        assert r.argmax == 0 and abs(s[0] - ref["scores"][0]) / ref["scores"][0] < RTOL
    ctx.close()


def test_maximum_channel_count_37(capi):
    """DPE_MAX_CHAN = CONST_PRN_MAX = 37 channels (the 12 satellites of the ephemeris set repeated: the channel arrays,
    the per-channel shared tables and the (PRN, lag) / (PRN, Doppler bin) bucket lists at their maximum): lookup, brute
    force and both velocity formulations against the oracle."""
    prns = (synth.PRNS_12 * 4)[:37]
    sc = H.scenario(2.5e6, prns)
    grid, tg = synth.uniform_grid(5, (5.0, 5.0, 5.0, 6.0))
    vgrid, _ = synth.uniform_grid(5, (0.5, 0.5, 0.5, 0.25))
    center = sc.rx_state(sc.cfg.rx_time0 + sc.cfg.T).copy()
    center[:4] += (3.0, -2.0, 1.0, 4.0)
    center[4:] += (0.6, -0.4, 0.2, 0.1)
    ep = sc.epoch_inputs(0, center=center, time_grid=tg)
    iq = sc.block(0)
    C = len(prns)
    assert C == 37 == sc.C
    bcs = orc.batch_corr_scores(iq, ep["prn"], ep["rc_start"], ep["ri_start"], ep["fc"], ep["fi"], ep["cp_start"],
                                ep["cp_ref"], ep["fs"], want_carrier=True)
    ref = H.oracle_pos(bcs, grid, ep)
    vref = orc.vel_meas_ml(bcs["carr_scores"], vgrid, ep["center"], ep["enu2ecef"], ep["sat_states"], ep["time_dim"],
                           ep["fi"], ep["doppler_sign"], ep["fs"], bcs["n_fft"])
    ctx = capi.Context(fs=ep["fs"], S=ep["S"], max_chan=C, G=grid.shape[0], time_dim=len(tg), lag_halfwidth=16,
                       Gv=vgrid.shape[0], dopp_halfwidth=32, flags=capi.FLAG_BRUTE_TILES | capi.FLAG_BRUTE_VEL)
    ctx.grid_set(grid)
    ctx.vel_grid_set(vgrid)
    for mode in (capi.SCORE_LOOKUP, capi.SCORE_BRUTE):
        for vel in (1, 2):
            r = ctx.epoch_run(iq, ep, score_mode=mode, with_vel=vel)
            s = ctx.copy_out(capi.PTR_POS_SCORES, np.float64, grid.shape[0])
            vs = ctx.copy_out(capi.PTR_VEL_SCORES, np.float64, vgrid.shape[0])
            assert r.argmax == ref["argmax"] and r.out_of_window == 0
            assert np.max(np.abs(s - ref["scores"]) / ref["scores"]) < RTOL
            assert r.vel_argmax == vref["argmax"] and r.vel_out_of_window == 0
            assert np.max(np.abs(vs - vref["scores"]) / vref["scores"]) < RTOL
    ctx.close()


@pytest.mark.parametrize("case", range(8))
def test_randomised_epochs_against_the_oracle(capi, case):
    """Eight seeded draws of everything an epoch is made of -- receiver seed and C/N0, the PRN subset (3..12 channels, in a
    shuffled order), the block, an IRREGULAR grid (not a lattice: every candidate drawn on its own, so buckets are ragged
    and lags / fractions arbitrary), the offset of the grid centre, the lag window, the estimator, LPower -- through
    lookup and brute force, weighted or arg-max, against the oracle: scores <= 1e-5, identical arg-max, the fix."""
    rng = np.random.RandomState(1000 + case)
    prns = list(synth.PRNS_12)
    rng.shuffle(prns)
    prns = tuple(prns[:int(rng.randint(3, 13))])
    sc = H.scenario(2.5e6, prns, seed=20180704 + 17 * case, cn0=float(rng.uniform(38.0, 50.0)))
    block = int(rng.randint(0, 4))
    G = int(rng.randint(200, 3000))
    spread = rng.uniform(5.0, 120.0)
    grid = np.ascontiguousarray(rng.uniform(-spread, spread, size=(G, 4)))
    grid[int(rng.randint(0, G))] = 0.0                             # the centre itself is a candidate
    T = int(rng.choice([1, 5, 9]))
    tg = np.sort(rng.uniform(-spread, spread, size=T))
    rx_time = sc.cfg.rx_time0 + (block + 1) * sc.cfg.T
    center = sc.rx_state(rx_time).copy()
    center[:4] += rng.uniform(-30.0, 30.0, size=4)
    ep = sc.epoch_inputs(block, center=center, time_grid=tg)
    iq = sc.block(block)
    lpower = int(rng.choice([1, 1, 2, 3]))
    weighted = bool(rng.randint(0, 2))
    W_ = int(rng.choice([8, 16, 32]))
    bcs = orc.batch_corr_scores(iq, ep["prn"], ep["rc_start"], ep["ri_start"], ep["fc"], ep["fi"], ep["cp_start"],
                                ep["cp_ref"], ep["fs"])
    ref = H.oracle_pos(bcs, grid, ep, lpower=lpower, weighted=weighted, per_time=weighted)
    ctx = capi.Context(fs=ep["fs"], S=ep["S"], max_chan=len(prns), G=G, time_dim=T, lpower=lpower, lag_halfwidth=W_,
                       flags=capi.FLAG_BRUTE_TILES)
    ctx.grid_set(grid)
    est = capi.EST_WEIGHTED if weighted else capi.EST_ARGMAX
    for mode in (capi.SCORE_LOOKUP, capi.SCORE_BRUTE):
        r, s_ = _scores(capi, ctx, iq, ep, mode, G, est)
        assert r.out_of_window == 0
        assert np.max(np.abs(s_ - ref["scores"]) / ref["scores"]) < RTOL, (case, mode)
        assert r.argmax == int(np.argmax(ref["scores"]))
        dz = np.abs(np.array(r.z[:4]) - ref["z"])
        assert dz[:3].max() < (1e-3 if weighted else 1e-9) and dz[3] < (1e-3 if weighted else 1e-9), (case, mode, dz)
    ctx.close()


def test_ten_megahertz_block_length_beyond_16_bits(capi):
    """S = 200000 overflows the reference's unsigned short block length (sampleblock.h:81)."""
    sc, iq, grid, ep = H.epoch_case(fs=10.0e6, prns=synth.PRNS_12, n=5, spacing=(2.0, 2.0, 2.0, 2.0))
    assert ep["S"] == 200000
    G = grid.shape[0]
    ctx = capi.Context(fs=ep["fs"], S=ep["S"], max_chan=12, G=G, time_dim=ep["time_dim"], lag_halfwidth=16,
                       flags=capi.FLAG_BRUTE_TILES)
    ctx.grid_set(grid)
    r_l, s_l = _scores(capi, ctx, iq, ep, capi.SCORE_LOOKUP, G)
    r_b, s_b = _scores(capi, ctx, iq, ep, capi.SCORE_BRUTE, G)
    assert np.max(np.abs(s_b - s_l) / s_l) < RTOL and r_b.argmax == r_l.argmax
    ctx.close()


# ---- BASELINE.json configs at size, each spot-checked against the oracle -----------
def _spot_check(capi, ctx, sc, grid_shard, ep, iq, bcs, lo, seed, n_spot=2000):
    """brute == lookup on every candidate of the shard, and n_spot random candidates == the oracle."""
    G = grid_shard.shape[0]
    r_l, s_l = _scores(capi, ctx, iq, ep, capi.SCORE_LOOKUP, G)
    r_b, s_b = _scores(capi, ctx, iq, ep, capi.SCORE_BRUTE, G)
    assert r_b.out_of_window == 0 and r_l.out_of_window == 0
    assert np.max(np.abs(s_b - s_l) / s_l) < RTOL
    assert r_b.argmax == r_l.argmax == lo + int(np.argmax(s_l))
    idx = np.random.default_rng(seed).choice(G, n_spot, replace=False)
    ref = H.oracle_pos(bcs, grid_shard[idx], ep)
    assert np.max(np.abs(s_l[idx] - ref["scores"]) / ref["scores"]) < 1e-6      # lookup: FP32 products in the correlogram
    assert np.max(np.abs(s_b[idx] - ref["scores"]) / ref["scores"]) < RTOL
    return r_b


def test_c3_full_grid_21_pow_4_times_12_prns_against_the_oracle(capi):
    """BASELINE.json config 3 at size: 2.5 MHz, 12 PRNs, uniform 21^4 grid (194 481 candidates)."""
    sc = H.scenario(2.5e6, synth.PRNS_12)
    grid, tg = synth.uniform_grid(21, (5.0, 5.0, 5.0, 6.0))
    center = sc.rx_state(sc.cfg.rx_time0 + sc.cfg.T).copy()
    center[:4] += (4.0, -3.0, 2.0, 5.0)
    ep = sc.epoch_inputs(0, center=center, time_grid=tg)
    iq = sc.block(0)
    ctx = capi.Context(fs=ep["fs"], S=ep["S"], max_chan=sc.C, G=grid.shape[0], time_dim=len(tg), lag_halfwidth=16,
                       flags=capi.FLAG_BRUTE_TILES)
    ctx.grid_set(grid)
    r = _spot_check(capi, ctx, sc, grid, ep, iq, H.oracle_bcs(2.5e6, synth.PRNS_12), 0, seed=3)
    assert ctx.brute_pairs() == grid.shape[0] * 12
    truth = sc.rx_state(ep["rx_time"])
    assert np.max(np.abs(np.array(r.z[:3]) - truth[:3])) <= 10.0 and abs(r.z[3] - truth[3]) <= 12.0   # within two grid steps
    ctx.close()


def test_c4_shard_rank_3_of_8_of_51_pow_4_at_10_megahertz_against_the_oracle(capi):
    """BASELINE.json config 4: what rank 3 of 8 holds of the 51^4 grid (845 651 candidates, 10 MHz, 12 PRNs,
    S = 200 000): global indices, per-shard arg-max, scores against the oracle."""
    sc = H.scenario(10.0e6, synth.PRNS_12)
    grid, tg = synth.uniform_grid(51, (2.0, 2.0, 2.0, 2.0))
    G = grid.shape[0]
    per = (G + 7) // 8
    lo, hi = 3 * per, 4 * per
    center = sc.rx_state(sc.cfg.rx_time0 + sc.cfg.T).copy()
    center[:4] += (4.0, -3.0, 2.0, 5.0)
    ep = sc.epoch_inputs(0, center=center, time_grid=tg)
    iq = sc.block(0)
    shard = np.ascontiguousarray(grid[lo:hi])
    ctx = capi.Context(fs=ep["fs"], S=ep["S"], max_chan=sc.C, G=hi - lo, time_dim=len(tg), lag_halfwidth=16,
                       flags=capi.FLAG_BRUTE_TILES, grid_offset=lo, G_total=G)
    ctx.grid_set(shard)
    _spot_check(capi, ctx, sc, shard, ep, iq, H.oracle_bcs(10.0e6, synth.PRNS_12), lo, seed=4)
    assert ctx.brute_pairs() == (hi - lo) * 12
    ctx.close()


def test_c5_independent_streams_own_seed_each_through_two_contexts(capi):
    """BASELINE.json config 5: independent receiver streams, seeds 20180704 + k, 21^4 grid each.  Streams
    k = 0, 1, 100, 255 go through two contexts in flight (what bench.py does for all 256); every stream's
    scores are spot-checked against the oracle and the fix lands within a grid step of its own truth."""
    grid, tg = synth.uniform_grid(21, (5.0, 5.0, 5.0, 6.0))
    G = grid.shape[0]
    ks = (0, 1, 100, 255)
    cases = []
    for k in ks:
        sc = H.scenario(2.5e6, synth.PRNS_8, 20180704 + k)
        center = sc.rx_state(sc.cfg.rx_time0 + sc.cfg.T).copy()
        center[:4] += (4.0, -3.0, 2.0, 5.0)
        cases.append((sc, sc.block(0), sc.epoch_inputs(0, center=center, time_grid=tg)))
    assert len({c[1].tobytes() for c in cases}) == len(ks)       # the streams really differ
    ctxs = []
    for _ in range(2):
        c = capi.Context(fs=2.5e6, S=cases[0][0].S, max_chan=8, G=G, time_dim=len(tg), lag_halfwidth=16,
                         flags=capi.FLAG_BRUTE_TILES)
        c.grid_set(grid)
        ctxs.append(c)
    out = {}
    for i, (sc, iq, ep) in enumerate(cases):
        c = ctxs[i % 2]
        if c.lib.dpe_epoch_pending(c.h):
            out[i - 2] = (c.epoch_collect(), c.copy_out(capi.PTR_POS_SCORES, np.float64, G))
        c.epoch_submit(iq, ep, score_mode=capi.SCORE_BRUTE)
    for i in (len(cases) - 2, len(cases) - 1):
        c = ctxs[i % 2]
        out[i] = (c.epoch_collect(), c.copy_out(capi.PTR_POS_SCORES, np.float64, G))
    for i, (sc, iq, ep) in enumerate(cases):
        r, s = out[i]
        idx = np.random.default_rng(50 + i).choice(G, 2000, replace=False)
        bcs = orc.batch_corr_scores(iq, ep["prn"], ep["rc_start"], ep["ri_start"], ep["fc"], ep["fi"], ep["cp_start"],
                                    ep["cp_ref"], ep["fs"])
        ref = H.oracle_pos(bcs, grid[idx], ep)
        assert np.max(np.abs(s[idx] - ref["scores"]) / ref["scores"]) < RTOL
        assert r.argmax == int(np.argmax(s)) and r.out_of_window == 0
        truth = sc.rx_state(ep["rx_time"])
        assert np.max(np.abs(np.array(r.z[:3]) - truth[:3])) <= 10.0 and abs(r.z[3] - truth[3]) <= 12.0
    for c in ctxs:
        c.close()
