#!/usr/bin/env python
"""bench.py -- DPE batch-correlation-manifold throughput on B200.

A "step" is one 20 ms epoch of the hot path over one synthetic block:
    int16 I/Q unpack + carrier wipe-off + C/A replica  ->  windowed correlogram (flip choice)
    ->  every (candidate, PRN) pair correlated against the whole block (brute-force BCM)
    ->  sum over PRNs, arg-max / score-weighted fix.
metric = candidate-PRN correlations per second (BASELINE.json), whole job over all ranks.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload demo|c3|c4|c5|tiny]
                    [--configs c3,c4,c5|none] [--impl reference]

Every epoch is ONE C-ABI call pair, dpe_epoch_submit / dpe_epoch_collect (include/dpe_b200.h), at
every N.  N > 1 (torchrun, one rank per GPU): the candidate grid is sharded in contiguous index
ranges; inside the library rank 0 uploads one packet {block, parameters, satellite states}, NCCL
broadcasts it, every rank scores its shard, the 16-double partials are all-gathered and reduced on
every rank (lowest global index wins arg-max ties) -- no Python between the stages.  Two contexts
per rank alternate, so the pre-pass / pair sort / reductions of one epoch run under the k_brute of
the other (the side kernels are sized to share an SM with k_brute).

`value`   K pipelined epochs, inputs resident in HBM, one CUDA-event bracket around all K.
`e2e`     the same with HOST buffers (page-locked block + parameters H2D, result D2H per epoch).
`latency` one epoch at a time (no overlap), L2 flushed between epochs, per-stage event brackets:
          the k_brute duration used for `roofline` comes from here.
`configs` sub-records for the other BASELINE.json workloads (c3, c4; c5 at N = 8), same code path.

--impl reference: the reference's CPU DPE path (NumPy restatement of PyGNSS / CUDARecv in
oracle/, the one place outside tests where oracle/ may be executed) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: fs, PRN set, grid, spacing, description (BASELINE.json configs)
    "demo": dict(fs=2.5e6, prns="8", grid="spread25", desc="demofile stand-in: synthetic static_opensky 2.5 MHz, 8 PRNs, "
                 "20 ms, 25^4 rngrid3-style spread grid (390625 candidates)"),
    "c3": dict(fs=2.5e6, prns="12", grid=("uniform", 21, (5.0, 5.0, 5.0, 6.0)),
               desc="synthetic L1 C/A 2.5 MHz, 12 PRNs, uniform 21^4 grid (194481 candidates)"),
    "c4": dict(fs=10.0e6, prns="12", grid=("uniform", 51, (2.0, 2.0, 2.0, 2.0)),
               desc="synthetic L1 C/A 10 MHz, 12 PRNs, uniform 51^4 grid (6765201 candidates)"),
    "c5": dict(fs=2.5e6, prns="8", grid=("uniform", 21, (5.0, 5.0, 5.0, 6.0)), streams=256,
               desc="256 independent synthetic receiver streams (seeds 20180704+k), 2.5 MHz, 8 PRNs, uniform 21^4 "
                    "grid each (streams sharded over ranks, no collective: replicas)"),
    "tiny": dict(fs=2.5e6, prns="8", grid=("uniform", 9, (5.0, 5.0, 5.0, 6.0)),
                 desc="synthetic 2.5 MHz, 8 PRNs, 9^4 grid (CI-sized)"),
}
FLOP_PER_SAMPLE_PAIR = 6.0     # 1 blend FMA + 2 accumulate FMAs per (candidate, PRN, sample); DESIGN.md section 4
SEED0 = 20180704


def build_workload(name, seed=SEED0):
    import dpe_pkg
    synth = dpe_pkg.submodule("synth")
    w = WORKLOADS[name]
    prns = synth.PRNS_8 if w["prns"] == "8" else synth.PRNS_12
    sc = synth.Scenario(synth.ScenarioConfig(fs=w["fs"], prns=prns, seed=seed))
    if w["grid"] == "spread25":
        grid = synth.spread_grid()
        tg = 6.0 * synth.spread_axis()
    else:
        _, n, sp = w["grid"]
        grid, tg = synth.uniform_grid(n, sp)
    return sc, grid, tg


def epoch_for_block(sc, b, tg, offset=(4.0, -3.0, 2.0, 5.0)):
    center = sc.rx_state(sc.cfg.rx_time0 + (b + 1) * sc.cfg.T).copy()
    center[:4] += offset
    return sc.epoch_inputs(b, center=center, time_grid=tg)


def config_for(args, workload=None):
    """The `config` object: identical in both arms (ours / --impl reference) for the same command line."""
    name = workload or args.workload
    w = WORKLOADS[name]
    fs = w["fs"]
    S = int(round(fs * 0.02))
    C = 8 if w["prns"] == "8" else 12
    n = 25 if w["grid"] == "spread25" else w["grid"][1]
    return dict(workload="%s: %s" % (name, w["desc"]), S=S, prns=C, candidates=n ** 4, streams=w.get("streams", 1),
                path=args.path, estimate=args.estimate,
                unit_of_work="one candidate-PRN pair scored per 20 ms epoch")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        self.path = "/tmp/dpe_clocks_%d_%d.csv" % (os.getpid(), int(time.time() * 1e3) % 100000)
        try:
            self.f = open(self.path, "w")
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            busy = [s for s in sm if s > 0.5 * max(sm)] or sm
            out.update(sm_mhz=float(np.median(busy)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        try:
            os.unlink(self.path)
        except OSError:
            pass
        return out


class Rig:
    """torch / torch.distributed plumbing shared by every measurement of one bench run (timing barrier,
    max over ranks, exchange of the NCCL unique ids).  The data path itself never touches torch."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        import dpe_pkg
        self.torch, self.dist = torch, dist
        self.capi = dpe_pkg.submodule("capi")
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)      # > 126 MB L2
        self.flush_stream = torch.cuda.Stream(device=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def unique_id(self):
        """ncclGetUniqueId on rank 0 (through the library), shipped to the other ranks."""
        t = self.torch.zeros(self.capi.DPE_COMM_ID_BYTES, dtype=self.torch.uint8, device=self.dev)
        if self.rank == 0:
            t.copy_(self.torch.frombuffer(bytearray(self.capi.comm_unique_id()), dtype=self.torch.uint8))
        self.dist.broadcast(t, 0)
        return bytes(t.cpu().numpy().tobytes())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def measure(rig, args, workload, steps, warmup, with_latency=True, with_e2e=True, with_other=False):
    """All numbers of one workload on this rank set.  Returns a dict on every rank."""
    torch, capi = rig.torch, rig.capi
    world, rank, dev = rig.world, rig.rank, rig.dev
    streams = WORKLOADS[workload].get("streams", 0)          # c5: independent receivers, sharded by stream
    sc, grid, tg = build_workload(workload)
    G_total, C, S, T = grid.shape[0], sc.C, sc.S, len(tg)
    if streams:
        lo, hi = 0, G_total
        my_streams = list(range(rank, streams, world))
    else:
        per = (G_total + world - 1) // world
        lo, hi = min(rank * per, G_total), min((rank + 1) * per, G_total)
        my_streams = [0]
        fake = int(os.environ.get("DPE_BENCH_SHARD_OF", "0"))       # tuning aid: one GPU holding what rank 3 of `fake` would
        if fake > 1 and world == 1:
            per = (G_total + fake - 1) // fake
            lo, hi = min(3 * per, G_total), min(4 * per, G_total)
    shard = np.ascontiguousarray(grid[lo:hi])
    score_mode = capi.SCORE_LOOKUP if args.path == "lookup" else capi.SCORE_BRUTE
    est_mode = capi.EST_WEIGHTED if args.estimate == "weighted" else capi.EST_ARGMAX
    use_comm = world > 1 and not streams
    W = args.lag_halfwidth

    depth = max(1, args.depth)
    ctxs = []
    for d in range(depth):
        ctx = capi.Context(fs=sc.cfg.fs, S=S, max_chan=C, G=hi - lo, time_dim=T, lag_halfwidth=W,
                           flags=capi.FLAG_BRUTE_TILES, device=rig.local, grid_offset=lo, G_total=G_total)
        ctx.grid_set(shard)
        if use_comm:
            ctx.comm_init(world, rank, rig.unique_id())
        ctxs.append(ctx)
    torch.cuda.synchronize()

    # inputs: `n_in` distinct (block, epoch parameters) sets; c5: one per stream this rank owns (own seed each)
    if streams:
        scen = [build_workload(workload, SEED0 + k)[0] for k in my_streams]
        inputs = [(s_.block(0), epoch_for_block(s_, 0, tg)) for s_ in scen]
    else:
        n_in = 8 if S <= 60000 else 3
        inputs = [(sc.block(b), epoch_for_block(sc, b, tg)) for b in range(n_in)]
    blocks_host = [torch.from_numpy(b.copy()).pin_memory() for b, _ in inputs]
    need_dev = rank == 0 or not use_comm
    blocks_dev = [b.to(dev) if need_dev else None for b in blocks_host]
    eps = [capi.make_epoch(e) for _, e in inputs]
    sats = [np.ascontiguousarray(e["sat_states"]) for _, e in inputs]
    n_in = len(inputs)
    epochs_per_step = len(my_streams)

    def run_steps(K, host, score=None, first=0):
        """K pipelined steps (c5: one epoch of every stream of this rank per step); returns the last result."""
        sm = score_mode if score is None else score
        res = None
        n = 0
        for i in range(K):
            with torch.cuda.stream(rig.flush_stream):
                rig.flush.zero_()                                   # L2 flush every step, on a side stream
            for s_ in range(epochs_per_step):
                ctx = ctxs[n % depth]
                if ctx.lib.dpe_epoch_pending(ctx.h):
                    res = ctx.epoch_collect()
                b = (first + i + s_) % n_in if not streams else s_
                src = (blocks_host if host else blocks_dev)[b] if need_dev else None
                ctx.epoch_submit(src, eps[b], sats[b], sm, est_mode, 0)
                n += 1
        for d in range(depth):
            ctx = ctxs[(n + d) % depth]
            if ctx.lib.dpe_epoch_pending(ctx.h):
                res = ctx.epoch_collect()
        return res

    def timed(K, Wm, host, score=None):
        run_steps(max(Wm, depth), host, score)                      # every context (and its communicator) has run once
        rig.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        res = run_steps(K, host, score, first=max(Wm, depth))
        e1.record()
        rig.barrier()
        wall = time.perf_counter() - t0
        return rig.max_over_ranks(e0.elapsed_time(e1)), wall, res

    pairs_per_step = G_total * C * (streams if streams else 1)
    out = dict(workload=workload)

    # --- pipelined throughput, resident inputs (`value`) -----------------------------------------
    sampler = ClockSampler(rig.local) if rank == 0 else None
    launches0 = sum(c.launch_count() for c in ctxs)
    ms_total, wall, res = timed(steps, warmup, host=False)
    launches = sum(c.launch_count() for c in ctxs) - launches0
    clocks = sampler.stop() if sampler else None
    ms_per_step = ms_total / steps
    out.update(ms_per_step=ms_per_step, value=pairs_per_step / (ms_per_step * 1e-3), wall_s=wall, clocks=clocks,
               launches_per_epoch=launches // ((steps + warmup) * epochs_per_step),
               gpu_launches=(launches // (steps + warmup)) * steps,
               fix=dict(z=[res.z[i] for i in range(4)], argmax=res.argmax, out_of_window=res.out_of_window))

    # --- end to end: host buffers through the same call ------------------------------------------
    if with_e2e:
        e2e_ms, _, _ = timed(steps, warmup, host=True)
        e2e_ms /= steps
        out["e2e"] = dict(value=pairs_per_step / (e2e_ms * 1e-3), unit="corr/s",
                          h2d_bytes_per_step=(4 * S + 8 * 8 * C * T + 2048) * epochs_per_step,    # rank 0's uploads
                          d2h_bytes_per_step=16 * 8 * epochs_per_step, ms_per_step=e2e_ms,
                          epochs_per_s=1e3 * (streams if streams else 1) / e2e_ms,
                          note="page-locked host block + parameters -> one H2D packet on rank 0 (-> ncclBroadcast), "
                               "result D2H per epoch, all inside dpe_epoch_submit / dpe_epoch_collect")

    # --- latency mode: one epoch at a time, L2 flushed between epochs, per-stage brackets ---------
    if with_latency:
        ctx = ctxs[0]
        K = max(3, min(steps, 10))
        for i in range(2):
            ctx.epoch_run_dist(blocks_dev[i % n_in] if need_dev else None, eps[i % n_in], sats[i % n_in], score_mode,
                               est_mode, 0)
        rig.barrier()
        ctx.profile_enable(True)
        own = torch.cuda.ExternalStream(ctx.stream(), device=dev)
        lat = 0.0
        for i in range(K):
            rig.flush.zero_()
            rig.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(own)
            ctx.epoch_run_dist(blocks_dev[i % n_in] if need_dev else None, eps[i % n_in], sats[i % n_in], score_mode,
                               est_mode, 0)
            e1.record(own)
            torch.cuda.synchronize()
            lat += e0.elapsed_time(e1)
        stage_ms, stage_cnt = ctx.profile_read()
        ctx.profile_enable(False)
        names = ("prepare", "correlogram", "lookup", "brute_bins", "brute_corr", "brute_score", "estimate", "velocity")
        out["latency"] = dict(ms_per_epoch=rig.max_over_ranks(lat / K), epochs=K,
                              stage_ms={n_: round(float(stage_ms[i] / K), 5) for i, n_ in enumerate(names)},
                              note="one epoch at a time on one context, 256 MiB L2 flush between epochs (untimed); "
                                   "the pair sort (brute_bins) overlaps prepare + correlogram on a second stream")
        if score_mode == capi.SCORE_BRUTE:
            k_ms = float(stage_ms[capi.STAGE_BRUTE_CORR] / max(int(stage_cnt[capi.STAGE_BRUTE_CORR]), 1))
            valid_pairs = ctx.brute_pairs()
            out["k_brute"] = dict(ms=k_ms, valid_pairs=valid_pairs, flop=FLOP_PER_SAMPLE_PAIR * S * valid_pairs)
        else:
            k_ms = float(stage_ms[capi.STAGE_LOOKUP] / max(int(stage_cnt[capi.STAGE_LOOKUP]), 1))
            out["k_lookup"] = dict(ms=k_ms, bytes=(32 + 8) * (hi - lo) + 16 * C * (2 * W + 2))

    # --- the other scoring path (lookup when the headline is brute and vice versa) -----------------
    if with_other:
        o_mode = capi.SCORE_LOOKUP if score_mode == capi.SCORE_BRUTE else capi.SCORE_BRUTE
        o_steps = max(steps, 50) if o_mode == capi.SCORE_LOOKUP else steps
        o_ms, _, _ = timed(o_steps, warmup, host=False, score=o_mode)
        o_e2e, _, _ = timed(o_steps, warmup, host=True, score=o_mode)
        # latency of the other path, one epoch at a time (resident inputs, no flush: launch-latency bound)
        ctx = ctxs[0]
        rig.barrier()
        t0 = time.perf_counter()
        for i in range(o_steps):
            ctx.epoch_run_dist(blocks_dev[i % n_in] if need_dev else None, eps[i % n_in], sats[i % n_in], o_mode, est_mode, 0)
        rig.barrier()
        o_lat = (time.perf_counter() - t0) / o_steps * 1e3
        launches1 = ctx.launch_count()
        ctx.epoch_run_dist(blocks_dev[0] if need_dev else None, eps[0], sats[0], o_mode, est_mode, 0)
        byts = (32 + 8) * (hi - lo) + 16 * C * S
        out["other_path"] = dict(path="lookup" if o_mode == capi.SCORE_LOOKUP else "brute", steps=o_steps,
                                 ms_per_step=o_ms / o_steps, epochs_per_s=1e3 * o_steps / o_ms,
                                 value=pairs_per_step * o_steps / (o_ms * 1e-3),
                                 e2e_value=pairs_per_step * o_steps / (o_e2e * 1e-3), e2e_ms_per_step=o_e2e / o_steps,
                                 e2e_epochs_per_s=1e3 * o_steps / o_e2e, latency_ms_per_epoch=o_lat,
                                 launches_per_epoch=ctx.launch_count() - launches1,
                                 hbm=dict(algorithmic_bytes_per_epoch=byts,
                                          achieved_gbs=byts / (o_ms / o_steps * 1e-3) / 1e9,
                                          note="SURVEY 8(d): 32 B grid + 8 B score per candidate + 16 B x C x S of "
                                               "correlogram the reference's formulation touches, / pipelined epoch time"))
    out["side_kernels"] = side_kernel_report(capi)
    for c in ctxs:
        c.close()
    return out


def velocity_brute_leg(rig, args, fp32_peak):
    """SURVEY 8 a', last sentence: the velocity / clock-drift manifold with every (velocity candidate, PRN) pair
    correlating the whole block against its own blended carrier (k_brute_vel), beside the lookup formulation
    (k_score_vel) on the same 25^4 velocity grid -- the grid the console flow uses."""
    torch, capi = rig.torch, rig.capi
    import dpe_pkg
    synth = dpe_pkg.submodule("synth")
    sc, grid, tg = build_workload("demo")
    vgrid, _ = synth.uniform_grid(25, (0.5, 0.5, 0.5, 0.25))
    C, S = sc.C, sc.S
    small = np.ascontiguousarray(grid[:4096])
    ctx = capi.Context(fs=sc.cfg.fs, S=S, max_chan=C, G=small.shape[0], time_dim=len(tg), lag_halfwidth=args.lag_halfwidth,
                       Gv=vgrid.shape[0], dopp_halfwidth=64, flags=capi.FLAG_BRUTE_VEL, device=rig.local)
    ctx.grid_set(small)
    ctx.vel_grid_set(vgrid)
    out = {}
    for mode, name in ((1, "lookup"), (2, "brute")):
        iqs = [sc.block(b) for b in range(3)]
        eps = [epoch_for_block(sc, b, tg) for b in range(3)]
        res = ctx.epoch_run_dist(iqs[0], eps[0], None, capi.SCORE_LOOKUP, capi.EST_ARGMAX, mode)      # warm-up (graph capture)
        ctx.profile_enable(True)
        K = 3
        for i in range(K):
            rig.flush.zero_()
            torch.cuda.synchronize()
            res = ctx.epoch_run_dist(iqs[i % 3], eps[i % 3], None, capi.SCORE_LOOKUP, capi.EST_ARGMAX, mode)
        stage_ms, stage_cnt = ctx.profile_read()
        ctx.profile_enable(False)
        ms = float(stage_ms[capi.STAGE_VELOCITY] / max(int(stage_cnt[capi.STAGE_VELOCITY]), 1))
        pairs = vgrid.shape[0] * C - res.vel_out_of_window
        rec = dict(stage_ms=ms, valid_pairs=int(pairs), vel_fix=[res.z[i] for i in range(4, 8)], vel_argmax=res.vel_argmax)
        if mode == 2:
            flop = 12.0 * S * pairs
            rec.update(flop=flop, achieved_tflops=flop / (ms * 1e-3) / 1e12, frac=flop / (ms * 1e-3) / 1e12 / fp32_peak,
                       algorithmic="12 FLOP x S x valid (velocity candidate, PRN) pairs: 1 FFMA2 blends the complex carrier, "
                                   "2 FFMA2 do the complex MAC; the stage time includes the plane, the pair sort and the "
                                   "reduction (5 side launches)")
        out[name] = rec
    ctx.close()
    out["same_fix"] = out["lookup"]["vel_argmax"] == out["brute"]["vel_argmax"]
    out["config"] = dict(velocity_candidates=int(vgrid.shape[0]), prns=C, S=S, doppler_halfwidth_bins=64)
    return out


def side_kernel_report(capi):
    """Registers x threads of every kernel: do the side kernels fit on an SM beside a k_brute CTA?"""
    rep = {}
    try:
        rb, _, _ = capi.kernel_attr("k_brute")
        alloc = lambda r: (r + 7) // 8 * 8
        free = 65536 - alloc(rb) * 256
        rep["k_brute_regs"] = rb
        rep["free_regs_per_sm_beside_k_brute"] = free
        side = {"k_prep_corr": 128, "k_sample_planes": 256,
                "k_replica_rd": 256, "k_pair_bins": 128, "k_block_scan": 256, "k_scatter": 128,
                "k_score_pairs": 128, "k_finalize": 32}
        worst = 0
        for k, thr in side.items():
            r, _, _ = capi.kernel_attr(k)
            rep[k] = [r, thr]
            worst = max(worst, alloc(r) * thr)
        rep["all_fit"] = worst <= free
    except Exception as exc:
        rep["error"] = repr(exc)
    return rep


def ncu_traffic(workload, timeout_s=150):
    """DRAM bytes of ONE k_brute launch, measured now: a child process (scripts/brute_probe.py on the same workload) under
    `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`, outside every timed region.  None when ncu is not
    available or refuses (no permission for the counters): the caller then falls back to the committed capture."""
    import shutil
    import subprocess
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "--print-units", "base",
           "-k", "k_brute", "-s", "1", "-c", "1", "--csv", sys.executable, os.path.join(ROOT, "scripts", "brute_probe.py"), workload]
    try:
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout_s, cwd=ROOT)
    except Exception:
        return None
    rd = wr = None
    for line in r.stdout.splitlines():
        if "dram__bytes_read.sum" in line or "dram__bytes_write.sum" in line:
            try:
                val = float(line.rstrip().rstrip('"').split('"')[-1].replace(",", ""))
            except ValueError:
                continue
            if "dram__bytes_read.sum" in line:
                rd = val
            else:
                wr = val
    if rd is None or wr is None:
        return None
    return dict(dram_read=rd, dram_write=wr, dram_bytes_per_launch=rd + wr,
                source="ncu child process of this bench run (scripts/brute_probe.py %s, second k_brute launch): "
                       "dram__bytes_read.sum + dram__bytes_write.sum" % workload)


def roofline_of(m, S, clocks, peaks, fp32_peak):
    if "k_brute" in m:
        kb = m["k_brute"]
        achieved = kb["flop"] / (kb["ms"] * 1e-3) / 1e12
        mhz = (clocks or {}).get("sm_mhz") or 1965.0
        nominal = 148 * 128 * 2 * mhz * 1e6 / 1e12
        step_frac = kb["flop"] / (m["ms_per_step"] * 1e-3) / 1e12 / fp32_peak if WORKLOADS[m["workload"]].get("streams") is None else None
        traffic, traffic_note = None, "no ncu capture committed for this workload / GPU count"
        try:                                                       # DRAM bytes of one launch, from the committed ncu capture
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_k_brute_traffic.json")))
            if tr.get("workload") == m["workload"] and int(os.environ.get("WORLD_SIZE", "1")) == tr.get("n_gpus"):
                traffic, traffic_note = tr["dram_bytes_per_launch"], tr["source"]
        except Exception:
            pass
        return dict(bound="fp32", kernel="k_brute", achieved=achieved, peak=fp32_peak, unit="TFLOP/s",
                    frac=achieved / fp32_peak, traffic=traffic, traffic_note=traffic_note,
                    peak_source="FFMA2 issue-limit micro-benchmark in this run (dpe_microbench_fp32: scalar multiplier, "
                                "shared pair); nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4",
                    algorithmic="6 FLOP x S x valid (candidate,PRN) pairs per launch (this rank's shard)",
                    kernel_ms=kb["ms"], kernel_share=kb["ms"] / m["latency"]["ms_per_epoch"],
                    kernel_share_pipelined=kb["ms"] / m["ms_per_step"] if step_frac is not None else None,
                    whole_step_frac=step_frac,
                    peak_nominal=nominal, frac_nominal=achieved / nominal,
                    nominal_source="148 SM x 128 FP32 lanes x 2 FLOP x %.0f MHz (median SM clock under load)" % mhz)
    kl = m["k_lookup"]
    hbm = peaks.get("hbm_gbs", 6650.0)
    achieved = kl["bytes"] / (kl["ms"] * 1e-3) / 1e9
    return dict(bound="hbm", kernel="k_score_lookup", achieved=achieved, peak=hbm, unit="GB/s", frac=achieved / hbm,
                traffic=None, peak_source="MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                algorithmic="32 B grid + 8 B score per candidate + correlogram window", kernel_ms=kl["ms"],
                kernel_share=kl["ms"] / m["latency"]["ms_per_epoch"])


def run_ours(args):
    rig = Rig()
    capi = rig.capi
    main = measure(rig, args, args.workload, args.steps, args.warmup, with_other=args.both)
    # the other BASELINE.json workloads, same code path, fewer steps (sub-records)
    subs = {}
    names = [n for n in args.configs.split(",") if n and n != "none"]
    if args.configs == "auto":
        names = ["c3", "c4"] + (["c5"] if rig.world >= 8 else [])
    for n in names:
        if n == args.workload or n not in WORKLOADS:
            continue
        k = dict(c3=10, c4=2 if rig.world == 1 else 4, c5=2, demo=10, tiny=10)[n]
        try:
            subs[n] = measure(rig, args, n, steps=k, warmup=1 if n in ("c4", "c5") else 3, with_e2e=(n != "c5"))
        except Exception as exc:                                   # reported, never silently dropped
            subs[n] = dict(error=repr(exc))
    if rig.rank != 0:
        rig.close()
        return
    vel_brute = None

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    fp32_peak = capi.microbench_fp32(rig.local, True)              # FFMA2 stream, measured in this run
    if main.get("other_path") and main["other_path"]["path"] == "lookup":
        # the lookup formulation is bound by the FP64 pipe's latency, not by HBM: report it against the DFMA rate measured here
        o = main["other_path"]
        fp64_peak = capi.microbench_fp64(rig.local)
        slots = 37.0 * (config_for(args)["candidates"] // rig.world) * config_for(args)["prns"]   # FP64-pipe instructions of k_score_lookup's loop, per pair (SASS)
        o["fp64_pipe"] = dict(peak_tflops=fp64_peak, fp64_instructions_per_pair=37,
                              achieved_tflops_equiv=2.0 * slots / (o["ms_per_step"] * 1e-3) / 1e12,
                              frac_of_epoch=2.0 * slots / (o["ms_per_step"] * 1e-3) / 1e12 / fp64_peak,
                              note="k_score_lookup executes 37 FP64-pipe instructions per (candidate, PRN) pair (static SASS count of its "
                                   "loop); counted as FMA slots (2 FLOP) over the pipelined epoch time against the DFMA stream measured "
                                   "in this run -- the kernel is latency bound (ncu: FP64 pipe 28 % active while it runs)")
    cfg = config_for(args)
    S = cfg["S"]
    roofline = roofline_of(main, S, main["clocks"], peaks, fp32_peak)
    if rig.world == 1 and args.path == "brute" and args.ncu_traffic:
        tr = ncu_traffic(args.workload)                            # measured in this run, outside the timed regions
        if tr:
            roofline["traffic"], roofline["traffic_note"] = tr["dram_bytes_per_launch"], tr["source"]
            roofline["traffic_read_write"] = [tr["dram_read"], tr["dram_write"]]
    if rig.world == 1 and args.workload == "demo" and args.vel_brute:
        try:
            vel_brute = velocity_brute_leg(rig, args, fp32_peak)
        except Exception as exc:
            vel_brute = dict(error=repr(exc))
    configs = {}
    for n, m in subs.items():
        if "error" in m:
            configs[n] = m
            continue
        c2 = config_for(args, n)
        configs[n] = dict(config=c2, n_gpus=rig.world, value=m["value"], unit="corr/s", ms_per_step=m["ms_per_step"],
                          epochs_per_s=1e3 * c2["streams"] / m["ms_per_step"],
                          e2e=m.get("e2e"), latency=m.get("latency"), clocks=m["clocks"],
                          roofline=roofline_of(m, c2["S"], m["clocks"], peaks, fp32_peak), fix=m["fix"],
                          launches_per_epoch=m["launches_per_epoch"])

    cpu = None if args.no_cpu_baseline else cpu_baseline(args.workload, budget_s=args.cpu_budget)
    flow = None
    ref_gpu = None
    if rig.world == 1 and args.flow_epochs > 0 and args.workload == "demo":
        try:
            flow = [flow_realtime(args, "brute"), flow_realtime(args, "lookup")]
        except Exception as exc:
            flow = dict(error=repr(exc))
        try:
            ref_gpu = reference_gpu_leg(args)
        except Exception as exc:
            ref_gpu = dict(error=repr(exc))
    like = None
    if cpu and main.get("other_path") and main["other_path"]["path"] == "lookup":
        o = main["other_path"]
        like = dict(ours_lookup_e2e=o["e2e_value"], ours_lookup_resident=o["value"], cpu_lookup=cpu["value"],
                    cpu_cores=cpu["cores"], ratio_e2e=o["e2e_value"] / cpu["value"],
                    note="like for like: both sides score a pair by interpolating a correlogram (the reference's own "
                         "formulation); the headline metric instead runs a full S-sample correlation per pair")
        if ref_gpu and "epochs_per_s" in ref_gpu:
            like["reference_gpu_epochs_per_s"] = ref_gpu["epochs_per_s"]
            like["ratio_vs_reference_gpu"] = o["e2e_epochs_per_s"] / ref_gpu["epochs_per_s"]

    streams = cfg["streams"]
    line = dict(metric="DPE candidate-PRN correlations/s (20 ms epochs, %s BCM)" % args.path, value=main["value"],
                unit="corr/s", n_gpus=rig.world, steps=args.steps, warmup=args.warmup, ms_per_step=main["ms_per_step"],
                higher_is_better=True, scaling="strong" if streams == 1 else "weak", vs_baseline=None,
                dtype="f32 (f64 geometry/bins)", data="synthetic", config=cfg,
                run=dict(lag_halfwidth=args.lag_halfwidth, contexts_in_flight=args.depth,
                         sharding=("independent streams, no collective" if streams > 1 else
                                   "grid candidates, contiguous index ranges; ncclBroadcast of the epoch packet + "
                                   "ncclAllGather of the partials inside libdpe_b200"),
                         l2="256 MiB flush memset every step on a side stream (inside the timed region); inputs rotate over "
                            "8 blocks and 2 contexts",
                         brute="a full S-sample correlation per pair (6*S FLOP); the reference arm and other_path=lookup "
                               "score a pair by interpolating a precomputed correlogram (same result)"),
                epochs_per_s=1e3 * streams / main["ms_per_step"],
                realtime_factor=(1e3 / main["ms_per_step"]) / 50.0,
                e2e=main["e2e"], gpu_launches=int(main["gpu_launches"]),
                gpu_launches_per_epoch=int(main["launches_per_epoch"]),
                latency=main.get("latency"), roofline=roofline, cpu_baseline=cpu, like_for_like=like,
                reference_gpu=ref_gpu, clocks=main["clocks"], other_path=main.get("other_path"), flow=flow,
                configs=configs, velocity_brute=vel_brute, side_kernels=main.get("side_kernels"), wall_s=main["wall_s"], fix=main["fix"])
    print(json.dumps(line))
    rig.close()


_FLOW_FILES = {}


import contextlib


@contextlib.contextmanager
def _c_stdout_to_stderr():
    """The console mirror answers commands on the process's stdout like the reference's console does ("Flow 0 (DPE)
    created."); bench.py's stdout carries exactly one JSON line, so file descriptor 1 points at stderr while a flow runs."""
    sys.stdout.flush()
    saved = os.dup(1)
    try:
        os.dup2(2, 1)
        yield
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def _flow_files(n_epochs):
    sc, grid, tg = build_workload("demo")
    d = "/tmp/dpe_bench_flow"
    need = 4 * sc.S * (n_epochs + 36)                                    # + the reader's 32-block read-ahead
    files = _FLOW_FILES.get("files")
    if not files or os.path.getsize(files["dat"]) < need:
        files = sc.write_files(d, n_epochs + 36, grid=grid, handoff_block=1)
        _FLOW_FILES["files"] = files
    return sc, d, files


def flow_realtime(args, path):
    """BASELINE.json config 2: the same synthetic capture pushed through the rebuilt `newflow dpe /
    loadflow / startflow` flow (file reader -> BCS -> BCM -> pass-through EKF -> host channel manager ->
    CSV logger), 25^4 position + 25^4 velocity grids.  Returns the flow's own per-epoch statistics
    (the reference's `[Flow] Average ... block duration`, flow.cu:172-191) and the real-time factor."""
    import dpe_pkg
    flowapi = dpe_pkg.submodule("flowapi")
    n_epochs = args.flow_epochs
    sc, d, files = _flow_files(n_epochs)
    with _c_stdout_to_stderr():
        return _flow_realtime(flowapi, sc, d, files, n_epochs, path)


def _flow_realtime(flowapi, sc, d, files, n_epochs, path):
    sh = flowapi.Shell()
    cmds = ["newflow dpe rx", "loadflow rx",
            'setparam rx SampleBlock Filename "%s"' % files["dat"],
            'setparam rx DPInit HandoffFilename "%s"' % files["handoff"],
            'setparam rx DPInit RINEXFilename "%s"' % files["rinex"],
            'setparam rx BatchCorrManifold LoadPosGridFilename "%s"' % files["grid"],
            "setparam rx BatchCorrManifold LoadPosGrid true",
            "setparam rx BatchCorrManifold PosGridDimSize 25", "setparam rx BatchCorrManifold VelGridDimSize 25",
            "setparam rx BatchCorrManifold GridDimSpacing 0.5",
            "setparam rx BatchCorrManifold BruteForce %s" % ("true" if path == "brute" else "false"),
            'setparam rx XECEFLogger Filename "%s/XFile.csv"' % d]
    for c in cmds:
        if sh.exec(c) != 0:
            raise RuntimeError("console command failed: " + c)
    if sh.run_blocking("rx", n_epochs) != 0:
        raise RuntimeError("flow failed")
    st = sh.stats("rx")
    rows = np.loadtxt(os.path.join(d, "XFile.csv"), delimiter=",")
    truth = sc.rx_state(sc.cfg.rx_time0 + (n_epochs + 1) * sc.cfg.T)
    err = float(np.linalg.norm(rows[-1, :3] - truth[:3]))
    sh.close()
    return dict(path=path, epochs=st["run_count"], avg_epoch_us=st["avg_us"], min_epoch_us=st["min_us"],
                max_epoch_us=st["max_us"], epochs_per_s=1e6 / st["avg_us"], realtime_factor=20000.0 / st["avg_us"],
                final_position_error_m=err,
                note="dpe_console flow on a synthetic 2.5 MHz capture: 25^4 spread position grid + 25^4 velocity "
                     "grid (0.5 m/s), 8 PRNs, host buffers, per-epoch D2H of the fix, CSV logging")


def reference_gpu_leg(args):
    """The reference's own kernels (CUDARecv rebuilt unmodified for sm_100a, oracle/_ref/ref_dpe -- a
    checker binary, executed here as a reported baseline only) on the same capture, same box."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_dpe")
    if not os.path.exists(exe):
        return None
    n_epochs = min(args.flow_epochs, 100)
    sc, d, files = _flow_files(args.flow_epochs)
    cmd = [exe, files["dat"], files["handoff"], files["rinex"], files["grid"], "25", "25", str(n_epochs),
           os.path.join(d, "ref_out"), "32", repr(sc.cfg.fs), "0"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=180)
    us = None
    for l in r.stdout.splitlines():
        if l.startswith("REF_MEAN_EPOCH_US"):
            us = float(l.split()[1])
    if r.returncode != 0 or not us:
        return dict(error="ref_dpe exit %d: %s" % (r.returncode, r.stdout[-300:]))
    return dict(avg_epoch_us=us, epochs_per_s=1e6 / us, epochs=n_epochs,
                value=390625 * 8 * 1e6 / us, unit="corr/s (lookups)",
                note="reference CUDARecv kernels rebuilt for sm_100a (FFT correlogram + BCM_PosMeasML lookup + 25^4 "
                     "velocity grid), timer as flow.cu:132-135")


# ---------------------------------------------------------------------------------------------
# CPU reference arm: NumPy restatement of the reference's DPE epoch (oracle/), host cores.
# ---------------------------------------------------------------------------------------------
_CPU_CACHE = {}


def _cpu_prepare(name, max_cand):
    """Per-process cache of the epoch inputs (building the synthetic scenario is not part of the path)."""
    key = (name, max_cand)
    if key not in _CPU_CACHE:
        sc, grid, tg = build_workload(name)
        if max_cand and grid.shape[0] > max_cand:
            grid = grid[:max_cand]
        _CPU_CACHE[key] = (sc, grid, [(sc.block(b), epoch_for_block(sc, b, tg)) for b in range(2)])
    return _CPU_CACHE[key]


def _cpu_epoch(a):
    """One epoch of the reference's CPU DPE path: FFT correlogram (BCS) + vectorised grid lookup and
    arg-max (BCM), oracle/dpe_oracle.py.  Returns (seconds, candidate-PRN pairs, arg-max)."""
    name, b, max_cand = a
    from oracle import dpe_oracle as orc
    sc, grid, eps = _cpu_prepare(name, max_cand)
    iq, ep = eps[b % len(eps)]
    t0 = time.perf_counter()
    bcs = orc.batch_corr_scores(iq, ep["prn"], ep["rc_start"], ep["ri_start"], ep["fc"], ep["fi"], ep["cp_start"],
                                ep["cp_ref"], ep["fs"])
    r = orc.pos_meas_ml(bcs["code_scores"], grid, ep["center"], ep["enu2ecef"], ep["sat_states"], ep["time_dim"],
                        ep["fc"], ep["rc_end"], ep["cp_ref_tow"], ep["cp_end"], ep["cp_ref"], ep["rx_time"],
                        ep["fs"], ep["S"])
    return time.perf_counter() - t0, grid.shape[0] * sc.C, int(r["argmax"])


def cpu_baseline(name, budget_s=20.0, max_cand=400000):
    """Single-process reference CPU path on a bounded sample (PyGNSS runs one thread per receiver)."""
    _cpu_prepare(name, max_cand)
    t_first, pairs, _ = _cpu_epoch((name, 0, max_cand))
    n = max(1, min(8, int(budget_s / max(t_first, 1e-3))))
    t0 = time.perf_counter()
    out = [_cpu_epoch((name, b, max_cand)) for b in range(n)]
    wall = time.perf_counter() - t0
    return dict(value=sum(o[1] for o in out) / wall, unit="corr/s", cores=1, kind="port",
                sample="%d epoch(s) of workload %s, %d candidate-PRN pairs per epoch (NumPy FFT correlogram + "
                       "grid lookup, oracle/dpe_oracle.py), one process" % (n, name, pairs),
                epochs_per_s=n / wall, s_per_epoch=wall / n)


def run_reference(args):
    """--impl reference: the reference's CPU DPE path on the host cores, one process per core over
    independent epochs (the way PyGNSS scales: one receiver per process).  A step = one epoch per
    process; inputs are built once per process outside the timed region."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs = max(1, min(os.cpu_count() or 1, args.cpu_procs))
    jobs = lambda i: [(args.workload, (i + k) % 2, 400000) for k in range(procs)]
    with mp.get_context("spawn").Pool(procs) as pool:
        pool.map(_cpu_epoch, jobs(0))                      # builds the per-process caches
        for i in range(args.warmup):
            pool.map(_cpu_epoch, jobs(i))
        t0 = time.perf_counter()
        pairs = 0
        for i in range(args.steps):
            pairs += sum(o[1] for o in pool.map(_cpu_epoch, jobs(i)))
        wall = time.perf_counter() - t0
    v = pairs / wall
    cfg = config_for(args)
    cpu = dict(value=v, unit="corr/s", cores=procs, kind="port",
               sample="%d step(s) x %d process(es), one epoch of workload %s per process and step "
                      "(NumPy FFT correlogram + grid lookup, oracle/dpe_oracle.py)" % (args.steps, procs, args.workload),
               epochs_per_s=args.steps * procs / wall)
    line = dict(impl="reference", metric="DPE candidate-PRN correlations/s (20 ms epochs, %s BCM)" % args.path,
                value=v, unit="corr/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * wall / args.steps, higher_is_better=True,
                scaling="strong" if cfg["streams"] == 1 else "weak", vs_baseline=None,
                dtype="f64", data="synthetic", config=cfg,
                run=dict(note="reference CPU DPE path (py3/NumPy restatement of PyGNSS / CUDARecv): a pair is scored by "
                              "interpolating the FFT correlogram; %d process(es), each step = one epoch per process" % procs),
                cpu_baseline=cpu, e2e=dict(value=v, unit="corr/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("DPE_BENCH_WORKLOAD", "demo"), choices=sorted(WORKLOADS))
    ap.add_argument("--configs", default=os.environ.get("DPE_BENCH_CONFIGS", "auto"),
                    help="comma list of further workloads timed as sub-records (auto: c3,c4 and c5 at >= 8 GPUs; none)")
    ap.add_argument("--path", default="brute", choices=["brute", "lookup"])
    ap.add_argument("--estimate", default="argmax", choices=["argmax", "weighted"])
    ap.add_argument("--lag-halfwidth", type=int, default=16)
    ap.add_argument("--depth", type=int, default=2, help="contexts in flight per rank (1 = no cross-epoch overlap)")
    ap.add_argument("--both", action="store_true", default=True, help="also time the other scoring path")
    ap.add_argument("--no-both", dest="both", action="store_false")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ncu-traffic", dest="ncu_traffic", action="store_false", default=True,
                    help="do not measure roofline.traffic with an ncu child process (1 GPU); use the committed capture")
    ap.add_argument("--no-vel-brute", dest="vel_brute", action="store_false", default=True,
                    help="skip the brute-force velocity manifold leg (1 GPU, demo)")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--cpu-procs", type=int, default=64)
    ap.add_argument("--flow-epochs", type=int, default=100,
                    help="epochs of the dpe_console flow run for the real-time factor (0 = skip; N=1, demo only)")
    args = ap.parse_args()
    if args.steps < 1 or args.warmup < 0:
        raise SystemExit("steps >= 1, warmup >= 0")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
