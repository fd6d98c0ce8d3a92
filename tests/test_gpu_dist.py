"""Asynchronous epochs (dpe_epoch_submit / dpe_epoch_collect), two contexts in flight, and the
in-library multi-GPU epoch (NCCL communicator per context, include/dpe_b200.h "multi-GPU").

The multi-GPU tests drive one rank per GPU from one THREAD per GPU of this process (what
dpe_console's NumGPUs does) and need >= 2 GPUs: `gpurun --gpus 2 -- pytest tests -m gpu`."""
import threading

import numpy as np
import pytest

import helpers as H
from helpers import orc, synth

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _mk(capi, sc, grid, ep, device=0, lo=0, hi=None, flags=None, W=16):
    hi = grid.shape[0] if hi is None else hi
    ctx = capi.Context(fs=ep["fs"], S=ep["S"], max_chan=sc.C, G=hi - lo, time_dim=ep["time_dim"], lag_halfwidth=W,
                       flags=capi.FLAG_BRUTE_TILES if flags is None else flags, device=device, grid_offset=lo,
                       G_total=grid.shape[0])
    ctx.grid_set(np.ascontiguousarray(grid[lo:hi]))
    return ctx


def _res_tuple(r):
    return (r.argmax, r.max_score, r.sum_score, r.out_of_window, tuple(r.z[i] for i in range(4)))


@pytest.mark.parametrize("mode", [0, 1])
def test_submit_collect_equals_epoch_run(capi, mode):
    sc, iq, grid, ep = H.epoch_case(n=9, center_offset=(4.0, -3.0, 2.0, 5.0))
    ctx = _mk(capi, sc, grid, ep)
    r0 = ctx.epoch_run(iq, ep, score_mode=mode)
    s0 = ctx.copy_out(capi.PTR_POS_SCORES, np.float64, grid.shape[0])
    ctx.epoch_submit(iq, ep, score_mode=mode)
    with pytest.raises(capi.DpeError) as e:                     # one epoch in flight per context
        ctx.epoch_submit(iq, ep, score_mode=mode)
    assert e.value.code == capi.DPE_ESTATE
    r1 = ctx.epoch_collect()
    s1 = ctx.copy_out(capi.PTR_POS_SCORES, np.float64, grid.shape[0])
    assert _res_tuple(r0) == _res_tuple(r1)
    assert np.array_equal(s0, s1)
    with pytest.raises(capi.DpeError):
        ctx.epoch_collect()                                     # nothing submitted
    # a device-resident block (zero copy) gives the same epoch
    import torch
    d_iq = torch.from_numpy(iq.copy()).cuda()
    r2 = ctx.epoch_run_dist(d_iq, ep, score_mode=mode)
    assert _res_tuple(r2) == _res_tuple(r0)
    ctx.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_two_contexts_in_flight_match_one_at_a_time(capi, mode):
    """Epochs alternate between two contexts with the next submitted before the previous is collected:
    same results as one context, one epoch at a time (bit for bit: nothing is shared between contexts)."""
    sc = H.scenario()
    grid, tg = synth.uniform_grid(11, (5.0, 5.0, 5.0, 6.0))
    eps, iqs = [], []
    for b in range(6):
        center = sc.rx_state(sc.cfg.rx_time0 + (b + 1) * sc.cfg.T).copy()
        center[:4] += (3.0 - b, -2.0, 1.0 + b, 4.0)
        eps.append(sc.epoch_inputs(b, center=center, time_grid=tg))
        iqs.append(sc.block(b))
    one = _mk(capi, sc, grid, eps[0])
    serial = [_res_tuple(one.epoch_run(iqs[b], eps[b], score_mode=mode)) for b in range(6)]
    one.close()
    ctxs = [_mk(capi, sc, grid, eps[0]) for _ in range(2)]
    got = [None] * 6
    for b in range(6):
        c = ctxs[b % 2]
        if c.lib.dpe_epoch_pending(c.h):
            got[b - 2] = _res_tuple(c.epoch_collect())
        c.epoch_submit(iqs[b], eps[b], score_mode=mode)
    for b in (4, 5):
        got[b] = _res_tuple(ctxs[b % 2].epoch_collect())
    assert got == serial
    for c in ctxs:
        c.close()


def test_side_kernels_fit_beside_k_brute(capi):
    """The overlap of two epochs relies on every kernel of a brute-force epoch except k_brute itself fitting on an SM
    next to a k_brute CTA (k_score_lookup only runs in lookup epochs, where there is no k_brute to share with)."""
    rb, _, tb = capi.kernel_attr("k_brute")
    free = 65536 - ((rb + 7) // 8 * 8) * 256
    for k, thr in (("k_prep_corr", 128), ("k_sample_planes", 256),
                   ("k_replica_rd", 256), ("k_pair_bins", 128), ("k_block_scan", 256), ("k_scatter", 128),
                   ("k_score_pairs", 128), ("k_finalize", 32)):
        r, _, _ = capi.kernel_attr(k)
        assert ((r + 7) // 8 * 8) * thr <= free, (k, r, thr, free)


def _run_ranks(capi, n, sc, grid, eps, iqs, mode, est):
    """One thread per rank / GPU: contexts with contiguous shards, NCCL communicator, dpe_epoch_run_dist."""
    uid = capi.comm_unique_id()
    G = grid.shape[0]
    per = (G + n - 1) // n
    out = [None] * n
    scores = [None] * n
    errs = []

    def rank_main(r):
        try:
            lo, hi = min(r * per, G), min((r + 1) * per, G)
            ctx = _mk(capi, sc, grid, eps[0], device=r, lo=lo, hi=hi)
            ctx.comm_init(n, r, uid)
            assert ctx.comm_info()[:2] == (n, r)
            res = []
            for b in range(len(eps)):
                res.append(_res_tuple(ctx.epoch_run_dist(iqs[b] if r == 0 else None, eps[b], score_mode=mode,
                                                         est_mode=est)))
            out[r] = res
            scores[r] = ctx.copy_out(capi.PTR_POS_SCORES, np.float64, hi - lo)
            ctx.close()
        except Exception as exc:       # surface in the main thread
            errs.append((r, repr(exc)))

    th = [threading.Thread(target=rank_main, args=(r,)) for r in range(n)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert not errs, errs
    assert all(not t.is_alive() for t in th), "a rank hung"
    return out, np.concatenate(scores)


@pytest.mark.parametrize("mode,est", [(0, 0), (0, 1), (1, 0)])
def test_nccl_ranks_equal_the_single_context(capi, mode, est):
    """N ranks (broadcast of rank 0's packet, sharded scoring, all-gather + finalize inside the library)
    == one context holding the whole grid: lookup scores / arg-max / fix bit for bit; brute force to
    FP32 summation order with the same arg-max."""
    n = min(_n_gpus(), 4)
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    sc = H.scenario()
    grid, tg = synth.uniform_grid(11, (5.0, 5.0, 5.0, 6.0))
    eps, iqs = [], []
    for b in range(3):
        center = sc.rx_state(sc.cfg.rx_time0 + (b + 1) * sc.cfg.T).copy()
        center[:4] += (4.0, -3.0 + b, 2.0, 5.0)
        eps.append(sc.epoch_inputs(b, center=center, time_grid=tg))
        iqs.append(sc.block(b))
    one = _mk(capi, sc, grid, eps[0])
    want = [_res_tuple(one.epoch_run(iqs[b], eps[b], score_mode=mode, est_mode=est)) for b in range(3)]
    want_scores = one.copy_out(capi.PTR_POS_SCORES, np.float64, grid.shape[0])
    one.close()
    got, got_scores = _run_ranks(capi, n, sc, grid, eps, iqs, mode, est)
    for r in range(n):
        for b in range(3):
            g, w = got[r][b], want[b]
            if est == 0:
                assert g[0] == w[0] and g[3] == w[3]                      # arg-max, out-of-window
                if mode == 0:
                    assert g[1] == w[1] and g[4] == w[4]                  # max score and fix bit for bit
                else:
                    assert abs(g[1] - w[1]) / w[1] < 2e-6 and g[4] == w[4]
            else:                                                          # weighted: sums in another order
                assert np.max(np.abs(np.array(g[4]) - np.array(w[4]))) < 1e-6
            assert abs(g[2] - w[2]) / w[2] < (1e-9 if mode == 0 else 2e-6)   # sum of scores
    assert got[0] == got[-1]                                               # every rank ends on the same estimate
    if mode == 0:
        assert np.array_equal(got_scores, want_scores)
    else:
        assert np.max(np.abs(got_scores - want_scores) / want_scores) < 2e-6


def test_context_on_another_device_than_the_callers(capi):
    """every entry binds the context's device and restores the caller's."""
    if _n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch
    torch.cuda.set_device(0)
    sc, iq, grid, ep = H.epoch_case(n=7)
    c0 = _mk(capi, sc, grid, ep, device=0)
    c1 = _mk(capi, sc, grid, ep, device=1)
    assert torch.cuda.current_device() == 0
    r0 = c0.epoch_run(iq, ep, score_mode=capi.SCORE_BRUTE)
    r1 = c1.epoch_run(iq, ep, score_mode=capi.SCORE_BRUTE)
    assert torch.cuda.current_device() == 0
    assert _res_tuple(r0) == _res_tuple(r1)
    s1 = c1.copy_out(capi.PTR_POS_SCORES, np.float64, grid.shape[0])
    s0 = c0.copy_out(capi.PTR_POS_SCORES, np.float64, grid.shape[0])
    assert np.array_equal(s0, s1)
    c0.close(); c1.close()


def test_context_create_failure_paths_do_not_leak(capi):
    """parameter errors are caught before anything is allocated; repeated failing creates leave
    the free memory where it was."""
    import torch
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(20):
        with pytest.raises(capi.DpeError):
            capi.Context(fs=2.5e6, S=50000, max_chan=8, G=10 ** 6, Gv=10 ** 6, n_fft=12345)   # n_fft not a power of two
        with pytest.raises(capi.DpeError):
            capi.Context(fs=2.5e6, S=50000, max_chan=8, G=10 ** 6, Gv=100, n_fft=1 << 25)      # twiddle angle not exact
        with pytest.raises(capi.DpeError):
            capi.Context(fs=2.5e6, S=64, max_chan=8, G=16, lag_halfwidth=40)                   # window wider than block
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < (64 << 20)
