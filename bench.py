#!/usr/bin/env python
"""bench.py -- DPE batch-correlation-manifold throughput on B200.

A "step" is one 20 ms epoch of the hot path over one synthetic block:
    int16 I/Q unpack + carrier wipe-off + C/A replica  ->  windowed correlogram (flip choice)
    ->  every (candidate, PRN) pair correlated against the whole block (brute-force BCM)
    ->  sum over PRNs, arg-max / score-weighted fix.
metric = candidate-PRN correlations per second (BASELINE.json), whole job over all ranks.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload demo|c3|c4|tiny] [--impl reference]

N > 1 (torchrun, one rank per GPU): the candidate grid is sharded in contiguous index ranges,
rank 0 broadcasts the 20 ms block over NCCL, every rank scores its shard, the per-rank partial
estimates are all-gathered and reduced (lowest global index wins arg-max ties).

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
               desc="256 independent synthetic receiver streams, 2.5 MHz, 8 PRNs, uniform 21^4 grid each "
                    "(streams sharded over ranks, no collective: replicas)"),
    "tiny": dict(fs=2.5e6, prns="8", grid=("uniform", 9, (5.0, 5.0, 5.0, 6.0)),
                 desc="synthetic 2.5 MHz, 8 PRNs, 9^4 grid (CI-sized)"),
}
FLOP_PER_SAMPLE_PAIR = 6.0     # 1 blend FMA + 2 accumulate FMAs per (candidate, PRN, sample); DESIGN.md section 4


def build_workload(name):
    import dpe_pkg
    synth = dpe_pkg.submodule("synth")
    w = WORKLOADS[name]
    prns = synth.PRNS_8 if w["prns"] == "8" else synth.PRNS_12
    sc = synth.Scenario(synth.ScenarioConfig(fs=w["fs"], prns=prns))
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


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        self.path = "/tmp/dpe_clocks_%d.csv" % os.getpid()
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
        return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import dpe_pkg
    capi = dpe_pkg.submodule("capi")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sc, grid, tg = build_workload(args.workload)
    G_total, C, S, T = grid.shape[0], sc.C, sc.S, len(tg)
    streams = WORKLOADS[args.workload].get("streams", 0)          # c5: independent receivers, sharded by stream
    if streams:
        lo, hi = 0, G_total
        my_streams = len(range(rank, streams, world))
    else:
        per = (G_total + world - 1) // world
        lo, hi = min(rank * per, G_total), min((rank + 1) * per, G_total)
    shard = np.ascontiguousarray(grid[lo:hi])
    score_mode = capi.SCORE_LOOKUP if args.path == "lookup" else capi.SCORE_BRUTE
    est_mode = capi.EST_WEIGHTED if args.estimate == "weighted" else capi.EST_ARGMAX
    sat_mode = capi.SAT_PER_TIME if est_mode == capi.EST_WEIGHTED else capi.SAT_MIDDLE

    ctx = capi.Context(fs=sc.cfg.fs, S=S, max_chan=C, G=hi - lo, time_dim=T, lag_halfwidth=args.lag_halfwidth,
                       flags=capi.FLAG_BRUTE_TILES, device=local, grid_offset=lo, G_total=G_total)
    ctx.grid_set(shard)
    n_blocks = 8 if streams else 4
    blocks_host = [torch.from_numpy(sc.block(b).copy()).pin_memory() for b in range(n_blocks)]
    blocks_dev = [b.to(dev) for b in blocks_host]                 # resident inputs for `value`
    epochs = [epoch_for_block(sc, b, tg) for b in range(n_blocks)]
    ep_structs = [capi.make_epoch(e) for e in epochs]
    sats = [np.ascontiguousarray(e["sat_states"]) for e in epochs]
    stream = torch.cuda.current_stream().cuda_stream
    aux = torch.cuda.Stream(device=dev)                            # the pair sort runs beside the sample pre-pass
    gathered = torch.zeros(world * capi.DPE_PARTIAL_LEN, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    recv = torch.empty(2 * S, dtype=torch.int16, device=dev)
    recv_u8 = recv.view(torch.uint8)                               # NCCL has no int16: broadcast the bytes

    def one_stream_epoch(b, host):
        if host:
            return ctx.epoch_run(blocks_host[b], ep_structs[b], sats[b], score_mode, est_mode, 0, stream)
        ctx.block_stage(blocks_dev[b], stream)
        ctx.epoch_set(ep_structs[b], sats[b], stream)
        if score_mode == capi.SCORE_BRUTE:
            ctx.brute_presort(sat_mode, aux.cuda_stream)
        ctx.replica_prepare(stream)
        ctx.correlogram(stream)
        ctx.score_pos(score_mode, sat_mode, stream)
        ctx.estimate(est_mode, None, 1, stream)

    def step_resident(i):
        b = i % n_blocks
        if streams:                                                # one epoch of every stream this rank owns
            for s_ in range(my_streams):
                one_stream_epoch((i + s_) % n_blocks, False)
            return
        if world > 1:
            if rank == 0:
                recv.copy_(blocks_dev[b], non_blocking=True)
            dist.broadcast(recv_u8, 0)
            ctx.block_stage(recv, stream)
        else:
            ctx.block_stage(blocks_dev[b], stream)
        ctx.epoch_set(ep_structs[b], sats[b], stream)
        if score_mode == capi.SCORE_BRUTE:
            ctx.brute_presort(sat_mode, aux.cuda_stream)
        ctx.replica_prepare(stream)
        ctx.correlogram(stream)
        ctx.score_pos(score_mode, sat_mode, stream)
        if world > 1:
            part = _as_tensor(torch, ctx.dev_ptr(capi.PTR_PARTIAL), capi.DPE_PARTIAL_LEN, dev)
            dist.all_gather_into_tensor(gathered, part)
            ctx.estimate(est_mode, gathered, world, stream)
        else:
            ctx.estimate(est_mode, None, 1, stream)

    def step_e2e(i):
        b = i % n_blocks
        if streams:
            r_ = None
            for s_ in range(my_streams):
                r_ = one_stream_epoch((i + s_) % n_blocks, True)
            return r_
        if world > 1:
            if rank == 0:
                recv.copy_(blocks_host[b], non_blocking=True)     # H2D from pinned memory, then NVLink broadcast
            dist.broadcast(recv_u8, 0)
            ctx.block_stage(recv, stream)
            ctx.epoch_set(ep_structs[b], sats[b], stream)
            if score_mode == capi.SCORE_BRUTE:
                ctx.brute_presort(sat_mode, aux.cuda_stream)
            ctx.replica_prepare(stream)
            ctx.correlogram(stream)
            ctx.score_pos(score_mode, sat_mode, stream)
            part = _as_tensor(torch, ctx.dev_ptr(capi.PTR_PARTIAL), capi.DPE_PARTIAL_LEN, dev)
            dist.all_gather_into_tensor(gathered, part)
            ctx.estimate(est_mode, gathered, world, stream)
            return ctx.result_fetch(stream)
        return ctx.epoch_run(blocks_host[b], ep_structs[b], sats[b], score_mode, est_mode, 0, stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, K, W, with_flush=True):
        for i in range(W):
            step_fn(i)
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        t0 = time.perf_counter()
        for i in range(K):
            if with_flush:
                flush.zero_()                                      # L2 flush between timed iterations (untimed)
            ev[i][0].record()
            step_fn(i)
            ev[i][1].record()
        barrier()
        wall = time.perf_counter() - t0
        ms = sum(a.elapsed_time(b) for a, b in ev)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), wall

    pairs_per_step = G_total * C * (streams if streams else 1)
    n_epochs_rank = my_streams if streams else 1
    # --- resident-input throughput (`value`) with per-stage device timing -----------------------
    sampler = ClockSampler(local) if rank == 0 else None
    ctx.profile_enable(True)
    launches0 = ctx.launch_count()
    ms_total, wall = timed(step_resident, args.steps, args.warmup)
    stage_ms, stage_cnt = ctx.profile_read()
    clocks = sampler.stop() if sampler else None
    launches = (ctx.launch_count() - launches0) // ((args.steps + args.warmup) * n_epochs_rank)
    ctx.profile_enable(False)
    res = ctx.result_fetch(stream)
    valid_pairs = ctx.brute_pairs() if score_mode == capi.SCORE_BRUTE else (hi - lo) * C
    ms_per_step = ms_total / args.steps
    value = pairs_per_step / (ms_per_step * 1e-3)

    # --- end to end: host buffers through the C-ABI epoch call ---------------------------------
    e2e_ms, _ = timed(step_e2e, args.steps, args.warmup)
    e2e_ms /= args.steps
    e2e_value = pairs_per_step / (e2e_ms * 1e-3)
    h2d = (4 * S + 8 * 8 * C * T + 2300) * n_epochs_rank         # block + sat states + dpe_epoch per epoch
    d2h = 16 * 8 * n_epochs_rank

    # --- the other path for context (lookup when the headline is brute and vice versa) ----------
    other = None
    if args.both:
        keep = score_mode
        score_mode = capi.SCORE_LOOKUP if keep == capi.SCORE_BRUTE else capi.SCORE_BRUTE
        o_ms, _ = timed(step_resident, args.steps, args.warmup)
        o_e2e, _ = timed(step_e2e, args.steps, args.warmup)
        other = dict(path="lookup" if score_mode == capi.SCORE_LOOKUP else "brute",
                     ms_per_step=o_ms / args.steps, epochs_per_s=1e3 * args.steps / o_ms,
                     value=pairs_per_step * args.steps / (o_ms * 1e-3),
                     e2e_value=pairs_per_step * args.steps / (o_e2e * 1e-3),
                     e2e_epochs_per_s=1e3 * args.steps / o_e2e)
        score_mode = keep

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # --- roofline of the dominant kernel ------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    if score_mode == capi.SCORE_BRUTE:
        fp32_peak = capi.microbench_fp32(local, True)             # FFMA2 stream, measured in this run
        n = max(int(stage_cnt[capi.STAGE_BRUTE_CORR]), 1)
        k_ms = stage_ms[capi.STAGE_BRUTE_CORR] / n                # this rank's k_brute, average per launch
        flop = FLOP_PER_SAMPLE_PAIR * S * valid_pairs
        achieved = flop / (k_ms * 1e-3) / 1e12
        roofline = dict(bound="fp32", kernel="k_brute", achieved=achieved, peak=fp32_peak, unit="TFLOP/s",
                        frac=achieved / fp32_peak, traffic=None,
                        peak_source="FFMA2 issue-limit micro-benchmark in this run (dpe_microbench_fp32: scalar multiplier, shared pair); nominal "
                                    "148 SM x 128 lanes x 2 x 1.965 GHz = 74.4",
                        algorithmic="6 FLOP x S x valid (candidate,PRN) pairs per launch",
                        kernel_ms=k_ms, kernel_share=stage_ms[capi.STAGE_BRUTE_CORR] / max(stage_ms.sum(), 1e-9))
        mhz = (clocks or {}).get("sm_mhz") or 1965.0
        nominal = 148 * 128 * 2 * mhz * 1e6 / 1e12
        try:                                                       # DRAM traffic of one launch, from the committed ncu capture
            tr = json.load(open(os.path.join(ROOT, "profiles", "k_brute_traffic.json")))
            if tr.get("workload") == args.workload and world == 1:
                roofline["traffic"] = tr["dram_bytes_per_launch"]
                roofline["traffic_source"] = tr["source"]
        except Exception:
            pass
        roofline.update(peak_nominal=nominal, frac_nominal=achieved / nominal,
                        nominal_source="148 SM x 128 FP32 lanes x 2 FLOP x %.0f MHz (median SM clock under load)" % mhz)
    else:
        n = max(int(stage_cnt[capi.STAGE_LOOKUP]), 1)
        k_ms = stage_ms[capi.STAGE_LOOKUP] / n
        byts = (32 + 8) * (hi - lo) + 16 * C * (2 * args.lag_halfwidth + 2)
        hbm = peaks.get("hbm_gbs", 6650.0)
        achieved = byts / (k_ms * 1e-3) / 1e9
        roofline = dict(bound="hbm", kernel="k_score_lookup", achieved=achieved, peak=hbm, unit="GB/s",
                        frac=achieved / hbm, traffic=None,
                        peak_source="MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                        algorithmic="32 B grid + 8 B score per candidate + correlogram window", kernel_ms=k_ms,
                        kernel_share=stage_ms[capi.STAGE_LOOKUP] / max(stage_ms.sum(), 1e-9))
    stages = {name: round(float(stage_ms[i] / max(args.steps + args.warmup, 1)), 5) for i, name in enumerate(
        ("prepare", "correlogram", "lookup", "brute_bins", "brute_corr", "brute_score", "estimate"))}

    cpu = None if args.no_cpu_baseline else cpu_baseline(args.workload, budget_s=args.cpu_budget)
    flow = None
    if world == 1 and args.flow_epochs > 0 and args.workload == "demo":
        try:
            ctx.close()
            flow = [flow_realtime(args, "brute"), flow_realtime(args, "lookup")]
        except Exception as exc:                                   # reported, never silently dropped
            flow = dict(error=repr(exc))

    line = dict(metric="DPE candidate-PRN correlations/s (20 ms epochs, %s BCM)" % args.path, value=value,
                unit="corr/s", n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_per_step,
                higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32 (f64 geometry/bins)",
                data="synthetic",
                config=dict(workload="%s: %s" % (args.workload, WORKLOADS[args.workload]["desc"]),
                            S=S, prns=C, candidates=G_total, path=args.path, estimate=args.estimate,
                            lag_halfwidth=args.lag_halfwidth,
                            sharding=("independent streams, %d per rank, no collective" % n_epochs_rank) if streams
                            else "grid candidates, contiguous index ranges",
                            l2="flushed (256 MiB memset) between timed iterations",
                            unit_of_work=("one candidate-PRN pair scored; brute = a full S-sample correlation per pair "
                                          "(6*S FLOP, the north-star kernel); the reference arm and other_path=lookup "
                                          "score a pair by interpolating a precomputed correlogram (same result)")),
                epochs_per_s=1e3 * (streams if streams else 1) / ms_per_step,
                realtime_factor=(1e3 / ms_per_step) / 50.0,
                e2e=dict(value=e2e_value, unit="corr/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                         ms_per_step=e2e_ms, epochs_per_s=1e3 / e2e_ms),
                gpu_launches=int(launches) * args.steps * n_epochs_rank, gpu_launches_per_epoch=int(launches),
                stage_ms_per_step=stages, roofline=roofline, cpu_baseline=cpu,
                clocks=clocks, other_path=other, flow=flow, wall_s=wall,
                fix=dict(z=[res.z[i] for i in range(4)], argmax=res.argmax, out_of_window=res.out_of_window))
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


_FLOW_FILES = {}


def flow_realtime(args, path):
    """BASELINE.json config 2: the same synthetic capture pushed through the rebuilt `newflow dpe /
    loadflow / startflow` flow (file reader -> BCS -> BCM -> pass-through EKF -> host channel manager ->
    CSV logger), 25^4 position + 25^4 velocity grids.  Returns the flow's own per-epoch statistics
    (the reference's `[Flow] Average ... block duration`, flow.cu:172-191) and the real-time factor."""
    import dpe_pkg
    flowapi = dpe_pkg.submodule("flowapi")
    synth = dpe_pkg.submodule("synth")
    sc, grid, tg = build_workload("demo")
    d = "/tmp/dpe_bench_flow"
    n_epochs = args.flow_epochs
    need = 4 * sc.S * (n_epochs + 34)                                    # + the reader's 32-block read-ahead
    files = _FLOW_FILES.get("files")
    if not files or os.path.getsize(files["dat"]) < need:
        files = sc.write_files(d, n_epochs + 34, grid=grid, handoff_block=0)
        _FLOW_FILES["files"] = files
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
    truth = sc.rx_state(sc.cfg.rx_time0 + n_epochs * sc.cfg.T)
    err = float(np.linalg.norm(rows[-1, :3] - truth[:3]))
    sh.close()
    return dict(path=path, epochs=st["run_count"], avg_epoch_us=st["avg_us"], min_epoch_us=st["min_us"],
                max_epoch_us=st["max_us"], epochs_per_s=1e6 / st["avg_us"], realtime_factor=20000.0 / st["avg_us"],
                final_position_error_m=err,
                note="dpe_console flow on a synthetic 2.5 MHz capture: 25^4 spread position grid + 25^4 velocity "
                     "grid (0.5 m/s), 8 PRNs, host buffers, per-epoch D2H of the fix, CSV logging")


def _as_tensor(torch, ptr, n, dev):
    """Zero-copy float64 view of a context-owned device buffer (for NCCL)."""
    class _A:
        __cuda_array_interface__ = dict(shape=(n,), typestr="<f8", data=(ptr, False), version=2)
    return torch.as_tensor(_A(), device=dev)


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
    cpu = dict(value=v, unit="corr/s", cores=procs, kind="port",
               sample="%d step(s) x %d process(es), one epoch of workload %s per process and step "
                      "(NumPy FFT correlogram + grid lookup, oracle/dpe_oracle.py)" % (args.steps, procs, args.workload),
               epochs_per_s=args.steps * procs / wall)
    line = dict(impl="reference", metric="DPE candidate-PRN correlations/s (20 ms epochs, %s BCM)" % args.path,
                value=v, unit="corr/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * wall / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None,
                dtype="f64", data="synthetic",
                config=dict(workload="%s: %s" % (args.workload, WORKLOADS[args.workload]["desc"]),
                            note="reference CPU DPE path (py3/NumPy restatement of PyGNSS / CUDARecv), "
                                 "%d process(es), each step = one epoch per process" % procs),
                cpu_baseline=cpu, e2e=dict(value=v, unit="corr/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="demo", choices=sorted(WORKLOADS))
    ap.add_argument("--path", default="brute", choices=["brute", "lookup"])
    ap.add_argument("--estimate", default="argmax", choices=["argmax", "weighted"])
    ap.add_argument("--lag-halfwidth", type=int, default=16)
    ap.add_argument("--both", action="store_true", default=True, help="also time the other scoring path")
    ap.add_argument("--no-both", dest="both", action="store_false")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
