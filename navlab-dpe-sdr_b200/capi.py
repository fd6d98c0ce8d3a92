"""ctypes binding of ``include/dpe_b200.h`` (libdpe_b200.so).

This is the only way Python code (tests, bench.py, the flow mirror) reaches the
CUDA kernels: plain pointers and sizes through the C ABI.  There is no CPU
fallback: if the shared library is missing, loading raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DPE_B200_LIB") or os.path.join(_HERE, "lib", "libdpe_b200.so")   # DPE_B200_LIB: another build of the library (A/B of kernel variants)

DPE_MAX_CHAN = 37
DPE_ABI_VERSION = 2
DPE_PARTIAL_LEN = 16
DPE_COMM_ID_BYTES = 128

DPE_OK, DPE_EINVAL, DPE_ECUDA, DPE_ENOMEM, DPE_ESTATE, DPE_EWINDOW, DPE_ECOMM = 0, -1, -2, -3, -4, -5, -6
SCORE_LOOKUP, SCORE_BRUTE = 0, 1
EST_ARGMAX, EST_WEIGHTED = 0, 1
SAT_MIDDLE, SAT_PER_TIME = 0, 1
(PTR_SAMPLES, PTR_CODE_SCORES, PTR_POS_SCORES, PTR_ZVAL, PTR_RVAL, PTR_GRID, PTR_PARTIAL,
 PTR_CHIP_IDX, PTR_XW, PTR_CARR_SCORES, PTR_VEL_SCORES, PTR_VEL_GRID, PTR_REPLICA_SIGN,
 PTR_CA_TABLE) = range(14)
FLAG_KEEP_CHIP_IDX, FLAG_BRUTE_TILES, FLAG_KEEP_BINS, FLAG_BRUTE_VEL = 1, 2, 4, 8
PART_CHANNELS, PART_GEOMETRY = 1, 2
DEBUG_BINS_EXACT = 16
(STAGE_PREPARE, STAGE_CORRELOGRAM, STAGE_LOOKUP, STAGE_BRUTE_BINS, STAGE_BRUTE_CORR, STAGE_BRUTE_SCORE,
 STAGE_ESTIMATE, STAGE_VELOCITY) = range(8)

EXPORTS = (
    "dpe_ctx_create", "dpe_ctx_destroy", "dpe_last_error", "dpe_abi_version", "dpe_grid_set",
    "dpe_vel_grid_set", "dpe_block_stage", "dpe_epoch_set", "dpe_epoch_set_part", "dpe_replica_prepare", "dpe_correlogram",
    "dpe_code_scores_set",
    "dpe_score_pos", "dpe_brute_presort", "dpe_estimate", "dpe_score_vel", "dpe_result_fetch", "dpe_epoch_run", "dpe_dev_ptr",
    "dpe_debug_channel_flags", "dpe_debug_bins", "dpe_debug_read", "dpe_launch_count", "dpe_profile_enable", "dpe_profile_read",
    "dpe_brute_pairs", "dpe_stream_create", "dpe_stream_destroy", "dpe_stream_sync", "dpe_host_alloc",
    "dpe_host_free", "dpe_device_count", "dpe_microbench_fp32",
    "dpe_microbench_hbm", "dpe_epoch_submit", "dpe_epoch_collect", "dpe_epoch_pending", "dpe_epoch_run_dist",
    "dpe_comm_get_unique_id", "dpe_comm_init", "dpe_comm_destroy", "dpe_comm_info", "dpe_ctx_stream",
    "dpe_kernel_attr", "dpe_epoch_set_device", "dpe_stream_create_on", "dpe_device_alloc",
    "dpe_device_free", "dpe_copy_h2d", "dpe_score_vel_brute", "dpe_fold_estimate", "dpe_microbench_fp64", "dpe_score_vel_est")


class DpeCfg(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("device", C.c_int32), ("fs", C.c_double), ("S", C.c_int64),
                ("max_chan", C.c_int32), ("time_dim", C.c_int32), ("G", C.c_int64), ("grid_offset", C.c_int64),
                ("G_total", C.c_int64), ("lpower", C.c_int32), ("lag_halfwidth", C.c_int32),
                ("flags", C.c_uint32), ("Gv", C.c_int64), ("n_fft", C.c_int32), ("dopp_halfwidth", C.c_int32)]


class DpeEpoch(C.Structure):
    _fields_ = [("C", C.c_int32), ("doppler_sign", C.c_int32), ("prn", C.c_uint8 * (DPE_MAX_CHAN + 3)),
                ("rc_start", C.c_double * DPE_MAX_CHAN), ("ri_start", C.c_double * DPE_MAX_CHAN),
                ("fc", C.c_double * DPE_MAX_CHAN), ("fi", C.c_double * DPE_MAX_CHAN),
                ("cp_start", C.c_int32 * DPE_MAX_CHAN), ("cp_ref", C.c_int32 * DPE_MAX_CHAN),
                ("rc_end", C.c_double * DPE_MAX_CHAN), ("cp_end", C.c_int32 * DPE_MAX_CHAN),
                ("cp_ref_tow", C.c_int32 * DPE_MAX_CHAN), ("rx_time", C.c_double),
                ("center", C.c_double * 8), ("enu2ecef", C.c_double * 9)]


class DpeResult(C.Structure):
    _fields_ = [("z", C.c_double * 8), ("max_score", C.c_double), ("sum_score", C.c_double),
                ("argmax", C.c_int64), ("out_of_window", C.c_int64), ("vel_max_score", C.c_double),
                ("vel_argmax", C.c_int64), ("vel_out_of_window", C.c_int64)]


class DpeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libdpe_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load_library(path: str | None = None):
    """dlopen libdpe_b200.so and declare the prototypes.  Raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise FileNotFoundError(
            "%s not found: build it with `make -C navlab-dpe-sdr_b200/csrc` "
            "(or __graft_entry__.build()); there is no CPU fallback" % p)
    lib = C.CDLL(p)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    lib.dpe_ctx_create.argtypes = [C.POINTER(vp), C.POINTER(DpeCfg)]
    lib.dpe_ctx_destroy.argtypes = [vp]
    lib.dpe_last_error.restype = C.c_char_p
    lib.dpe_grid_set.argtypes = [vp, vp, i64, vp]
    lib.dpe_vel_grid_set.argtypes = [vp, vp, i64, vp]
    lib.dpe_block_stage.argtypes = [vp, vp, i64, vp]
    lib.dpe_epoch_set.argtypes = [vp, C.POINTER(DpeEpoch), vp, vp]
    lib.dpe_epoch_set_part.argtypes = [vp, C.POINTER(DpeEpoch), vp, C.c_uint, vp]
    lib.dpe_replica_prepare.argtypes = [vp, vp]
    lib.dpe_correlogram.argtypes = [vp, vp]
    lib.dpe_code_scores_set.argtypes = [vp, vp, i32, vp]
    lib.dpe_score_pos.argtypes = [vp, i32, i32, vp]
    lib.dpe_brute_presort.argtypes = [vp, i32, vp]
    lib.dpe_estimate.argtypes = [vp, i32, vp, i32, vp]
    lib.dpe_fold_estimate.argtypes = [vp, i32]
    lib.dpe_score_vel.argtypes = [vp, vp]
    lib.dpe_score_vel_brute.argtypes = [vp, vp]
    lib.dpe_score_vel_est.argtypes = [vp, i32, vp]
    lib.dpe_result_fetch.argtypes = [vp, C.POINTER(DpeResult), vp]
    lib.dpe_epoch_run.argtypes = [vp, vp, C.POINTER(DpeEpoch), vp, i32, i32, i32, C.POINTER(DpeResult), vp]
    lib.dpe_dev_ptr.argtypes = [vp, i32]
    lib.dpe_dev_ptr.restype = vp
    lib.dpe_debug_channel_flags.argtypes = [vp, vp, vp, i32]
    lib.dpe_debug_bins.argtypes = [vp, i64, i64, i32, vp, vp, vp]
    lib.dpe_debug_read.argtypes = [vp, i32, C.c_size_t, vp, C.c_size_t]
    lib.dpe_launch_count.argtypes = [vp]
    lib.dpe_launch_count.restype = i64
    lib.dpe_profile_enable.argtypes = [vp, i32]
    lib.dpe_profile_read.argtypes = [vp, vp, vp]
    lib.dpe_brute_pairs.argtypes = [vp]
    lib.dpe_brute_pairs.restype = i64
    lib.dpe_epoch_submit.argtypes = [vp, vp, C.POINTER(DpeEpoch), vp, i32, i32, i32]
    lib.dpe_epoch_collect.argtypes = [vp, C.POINTER(DpeResult)]
    lib.dpe_epoch_pending.argtypes = [vp]
    lib.dpe_epoch_run_dist.argtypes = [vp, vp, C.POINTER(DpeEpoch), vp, i32, i32, i32, C.POINTER(DpeResult)]
    lib.dpe_comm_get_unique_id.argtypes = [vp]
    lib.dpe_comm_init.argtypes = [vp, i32, i32, vp]
    lib.dpe_comm_destroy.argtypes = [vp]
    lib.dpe_comm_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    lib.dpe_ctx_stream.argtypes = [vp]
    lib.dpe_ctx_stream.restype = vp
    lib.dpe_kernel_attr.argtypes = [C.c_char_p, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    lib.dpe_microbench_fp32.argtypes = [i32, i32, C.POINTER(C.c_double)]
    lib.dpe_microbench_hbm.argtypes = [i32, C.c_size_t, C.POINTER(C.c_double)]
    lib.dpe_microbench_fp64.argtypes = [i32, C.POINTER(C.c_double)]
    if path is None:
        _lib = lib
    return lib


def _check(lib, rc):
    if rc != 0:
        raise DpeError(rc, lib.dpe_last_error().decode())


def make_epoch(ep: dict) -> DpeEpoch:
    """Build a dpe_epoch from the dict produced by synth.Scenario.epoch_inputs /
    the channel manager (keys: prn, rc_start, ri_start, fc, fi, cp_start, cp_ref,
    rc_end, cp_end, cp_ref_tow, rx_time, center, enu2ecef, doppler_sign)."""
    e = DpeEpoch()
    n = len(ep["prn"])
    e.C = n
    e.doppler_sign = int(ep.get("doppler_sign", 1))
    for i in range(n):
        e.prn[i] = int(ep["prn"][i])
        e.rc_start[i] = float(ep["rc_start"][i]); e.ri_start[i] = float(ep["ri_start"][i])
        e.fc[i] = float(ep["fc"][i]); e.fi[i] = float(ep["fi"][i])
        e.cp_start[i] = int(ep["cp_start"][i]); e.cp_ref[i] = int(ep["cp_ref"][i])
        e.rc_end[i] = float(ep["rc_end"][i]); e.cp_end[i] = int(ep["cp_end"][i])
        e.cp_ref_tow[i] = int(ep["cp_ref_tow"][i])
    e.rx_time = float(ep["rx_time"])
    for i in range(8):
        e.center[i] = float(ep["center"][i])
    for i in range(9):
        e.enu2ecef[i] = float(ep["enu2ecef"][i])
    return e


def _ptr(a):
    """Pointer of a numpy array, a torch tensor, an int address or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


class Context:
    """One dpe_ctx (one flow on one GPU).  Thin: every method is one C-ABI call."""

    def __init__(self, fs, S, max_chan, G, time_dim=1, lpower=1, lag_halfwidth=32, flags=0, device=0,
                 grid_offset=0, G_total=None, Gv=0, n_fft=0, dopp_halfwidth=0):
        self.lib = load_library()
        cfg = DpeCfg(DPE_ABI_VERSION, device, float(fs), int(S), int(max_chan), int(time_dim), int(G),
                     int(grid_offset), int(G_total if G_total is not None else grid_offset + G),
                     int(lpower), int(lag_halfwidth), int(flags), int(Gv), int(n_fft), int(dopp_halfwidth))
        self.cfg = cfg
        self.S, self.G, self.W, self.T = int(S), int(G), int(lag_halfwidth), int(time_dim)
        self.h = C.c_void_p()
        _check(self.lib, self.lib.dpe_ctx_create(C.byref(self.h), C.byref(cfg)))
        self._keep = []

    def close(self):
        if self.h:
            self.lib.dpe_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- stages -------------------------------------------------------------
    def grid_set(self, grid, stream=0):
        g = grid if hasattr(grid, "data_ptr") else np.ascontiguousarray(grid, dtype=np.float64)
        self._keep = [g]
        n = g.shape[0]
        _check(self.lib, self.lib.dpe_grid_set(self.h, _ptr(g), n, C.c_void_p(stream)))

    def vel_grid_set(self, vgrid, stream=0):
        g = np.ascontiguousarray(vgrid, dtype=np.float64)
        self._vgrid = g
        _check(self.lib, self.lib.dpe_vel_grid_set(self.h, _ptr(g), g.shape[0], C.c_void_p(stream)))

    def fold_estimate(self, est_mode):
        _check(self.lib, self.lib.dpe_fold_estimate(self.h, est_mode))

    def score_vel(self, stream=0):
        _check(self.lib, self.lib.dpe_score_vel(self.h, C.c_void_p(stream)))

    def score_vel_est(self, est_mode, stream=0):
        _check(self.lib, self.lib.dpe_score_vel_est(self.h, est_mode, C.c_void_p(stream)))

    def score_vel_brute(self, stream=0):
        _check(self.lib, self.lib.dpe_score_vel_brute(self.h, C.c_void_p(stream)))

    def block_stage(self, iq, stream=0):
        if isinstance(iq, np.ndarray):
            iq = np.ascontiguousarray(iq, dtype=np.int16)
        self._iq = iq
        _check(self.lib, self.lib.dpe_block_stage(self.h, _ptr(iq), self.S, C.c_void_p(stream)))

    def epoch_set(self, ep, sat_states=None, stream=0):
        e = ep if isinstance(ep, DpeEpoch) else make_epoch(ep)
        sat = sat_states if sat_states is not None else ep["sat_states"]
        sat = np.ascontiguousarray(sat, dtype=np.float64)
        assert sat.size == e.C * self.T * 8, "sat_states must be [C][T][8]"
        self._ep, self._sat = e, sat
        _check(self.lib, self.lib.dpe_epoch_set(self.h, C.byref(e), _ptr(sat), C.c_void_p(stream)))

    def epoch_set_part(self, ep, parts, sat_states=None, stream=0):
        e = ep if isinstance(ep, DpeEpoch) else make_epoch(ep)
        sat = None
        if parts & PART_GEOMETRY:
            sat = sat_states if sat_states is not None else ep["sat_states"]
            sat = np.ascontiguousarray(sat, dtype=np.float64)
        self._ep, self._sat = e, sat
        _check(self.lib, self.lib.dpe_epoch_set_part(self.h, C.byref(e), _ptr(sat), parts, C.c_void_p(stream)))

    def replica_prepare(self, stream=0):
        _check(self.lib, self.lib.dpe_replica_prepare(self.h, C.c_void_p(stream)))

    def correlogram(self, stream=0):
        _check(self.lib, self.lib.dpe_correlogram(self.h, C.c_void_p(stream)))

    def code_scores_set(self, cs, stream=0):
        """cs: complex128 [C][2W+2] window of an externally produced correlogram."""
        cs = np.ascontiguousarray(cs, dtype=np.complex128)
        self._cs = cs
        _check(self.lib, self.lib.dpe_code_scores_set(self.h, _ptr(cs), cs.shape[0], C.c_void_p(stream)))

    def score_pos(self, score_mode=SCORE_LOOKUP, sat_mode=SAT_MIDDLE, stream=0):
        _check(self.lib, self.lib.dpe_score_pos(self.h, score_mode, sat_mode, C.c_void_p(stream)))

    def brute_presort(self, sat_mode=SAT_MIDDLE, stream=0):
        """Sort the pairs for SCORE_BRUTE ahead of time (e.g. on a second stream beside the pre-pass)."""
        _check(self.lib, self.lib.dpe_brute_presort(self.h, sat_mode, C.c_void_p(stream)))

    def estimate(self, est_mode=EST_ARGMAX, gathered=None, nranks=1, stream=0):
        _check(self.lib, self.lib.dpe_estimate(self.h, est_mode, _ptr(gathered), nranks, C.c_void_p(stream)))

    def result_fetch(self, stream=0) -> DpeResult:
        r = DpeResult()
        _check(self.lib, self.lib.dpe_result_fetch(self.h, C.byref(r), C.c_void_p(stream)))
        return r

    def epoch_run(self, iq_host, ep, sat_states=None, score_mode=SCORE_LOOKUP, est_mode=EST_ARGMAX,
                  with_vel=0, stream=0) -> DpeResult:
        e = ep if isinstance(ep, DpeEpoch) else make_epoch(ep)
        sat = sat_states if sat_states is not None else ep["sat_states"]
        sat = np.ascontiguousarray(sat, dtype=np.float64)
        r = DpeResult()
        _check(self.lib, self.lib.dpe_epoch_run(self.h, _ptr(iq_host), C.byref(e), _ptr(sat), score_mode,
                                                est_mode, with_vel, C.byref(r), C.c_void_p(stream)))
        return r

    # -- asynchronous epochs / multi-GPU ----------------------------------------
    def epoch_submit(self, iq, ep, sat_states=None, score_mode=SCORE_LOOKUP, est_mode=EST_ARGMAX, with_vel=0):
        """Enqueue one whole epoch on the context's own stream (returns at once)."""
        e = ep if isinstance(ep, DpeEpoch) else make_epoch(ep)
        sat = sat_states if sat_states is not None else (ep["sat_states"] if isinstance(ep, dict) else None)
        if sat is not None and not hasattr(sat, "data_ptr"):
            sat = np.ascontiguousarray(sat, dtype=np.float64)
        if isinstance(iq, np.ndarray):
            iq = np.ascontiguousarray(iq, dtype=np.int16)
        self._sub = (iq, e, sat)                      # keep alive until collected
        _check(self.lib, self.lib.dpe_epoch_submit(self.h, _ptr(iq), C.byref(e), _ptr(sat), score_mode, est_mode,
                                                   with_vel))

    def epoch_collect(self) -> DpeResult:
        r = DpeResult()
        _check(self.lib, self.lib.dpe_epoch_collect(self.h, C.byref(r)))
        return r

    def epoch_run_dist(self, iq, ep, sat_states=None, score_mode=SCORE_LOOKUP, est_mode=EST_ARGMAX,
                       with_vel=0) -> DpeResult:
        self.epoch_submit(iq, ep, sat_states, score_mode, est_mode, with_vel)
        return self.epoch_collect()

    def comm_init(self, nranks, rank, unique_id: bytes):
        """Collective: join the NCCL communicator of this context (one per context)."""
        assert len(unique_id) == DPE_COMM_ID_BYTES
        buf = C.create_string_buffer(bytes(unique_id), DPE_COMM_ID_BYTES)
        _check(self.lib, self.lib.dpe_comm_init(self.h, nranks, rank, buf))

    def comm_destroy(self):
        _check(self.lib, self.lib.dpe_comm_destroy(self.h))

    def comm_info(self):
        n, r, v = C.c_int(), C.c_int(), C.c_int()
        _check(self.lib, self.lib.dpe_comm_info(self.h, C.byref(n), C.byref(r), C.byref(v)))
        return n.value, r.value, v.value

    def stream(self) -> int:
        p = self.lib.dpe_ctx_stream(self.h)
        return int(p) if p else 0

    # -- access ---------------------------------------------------------------
    def dev_ptr(self, which) -> int:
        p = self.lib.dpe_dev_ptr(self.h, which)
        return int(p) if p else 0

    def launch_count(self) -> int:
        return int(self.lib.dpe_launch_count(self.h))

    def profile_enable(self, on=True):
        _check(self.lib, self.lib.dpe_profile_enable(self.h, 1 if on else 0))

    def profile_read(self):
        """(ms[8], count[8]) accumulated per stage since the last read (synchronises)."""
        ms = np.zeros(8, dtype=np.float64)
        cnt = np.zeros(8, dtype=np.int64)
        _check(self.lib, self.lib.dpe_profile_read(self.h, _ptr(ms), _ptr(cnt)))
        return ms, cnt

    def brute_pairs(self) -> int:
        return int(self.lib.dpe_brute_pairs(self.h))

    def channel_flags(self, n_chan):
        a = np.zeros(n_chan, dtype=np.int32)
        b = np.zeros(n_chan, dtype=np.int32)
        _check(self.lib, self.lib.dpe_debug_channel_flags(self.h, _ptr(a), _ptr(b), n_chan))
        return a, b

    def debug_bins(self, i0, n, n_chan, sat_mode=SAT_MIDDLE, stream=0, exact=False):
        """Bins of candidates [i0, i0+n): (floor index, lerp weight).  exact=True evaluates the reference's FP64
        chain literally; the default is the centre-relative form the scoring kernels use (bit-identical)."""
        f = np.zeros((n, n_chan), dtype=np.int64)
        a = np.zeros((n, n_chan), dtype=np.float64)
        _check(self.lib, self.lib.dpe_debug_bins(self.h, i0, n, sat_mode | (DEBUG_BINS_EXACT if exact else 0), _ptr(f),
                                                 _ptr(a), C.c_void_p(stream)))
        return f, a

    def copy_out(self, which, dtype, count, offset=0):
        """Synchronous D2H read of a context buffer (tests / smoke)."""
        out = np.empty(count, dtype=dtype)
        _check(self.lib, self.lib.dpe_debug_read(self.h, which, offset, _ptr(out), out.nbytes))
        return out


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the library (rank 0 calls it and ships the bytes to the others)."""
    lib = load_library()
    buf = C.create_string_buffer(DPE_COMM_ID_BYTES)
    _check(lib, lib.dpe_comm_get_unique_id(buf))
    return buf.raw


def kernel_attr(name: str):
    """(registers per thread, static shared bytes, max threads per block) of a library kernel."""
    lib = load_library()
    r, sm, mt = C.c_int(), C.c_int(), C.c_int()
    _check(lib, lib.dpe_kernel_attr(name.encode(), C.byref(r), C.byref(sm), C.byref(mt)))
    return r.value, sm.value, mt.value


def microbench_fp32(device=0, use_ffma2=True) -> float:
    lib = load_library()
    v = C.c_double()
    _check(lib, lib.dpe_microbench_fp32(device, 1 if use_ffma2 else 0, C.byref(v)))
    return v.value


def microbench_fp64(device=0) -> float:
    lib = load_library()
    v = C.c_double()
    _check(lib, lib.dpe_microbench_fp64(device, C.byref(v)))
    return v.value


def microbench_hbm(device=0, nbytes=1 << 30) -> float:
    lib = load_library()
    v = C.c_double()
    _check(lib, lib.dpe_microbench_hbm(device, nbytes, C.byref(v)))
    return v.value
