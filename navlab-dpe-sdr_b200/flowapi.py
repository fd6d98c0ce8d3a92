"""ctypes binding of ``include/dpe_flow.h`` (libdpe_flow.so): the console / FlowMgr / DPEFlow
mirror of the reference's ``cudarecv`` front end, driven line by line like a console user."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdpe_flow.so")
CONSOLE_PATH = os.path.join(_HERE, "lib", "dpe_console")
EXPORTS = ("dpe_shell_create", "dpe_shell_destroy", "dpe_shell_exec", "dpe_shell_run_blocking",
           "dpe_shell_flow_stats", "dpe_shell_read_port", "dpe_host_sat_position", "dpe_host_make_grid",
           "dpe_host_read_handoff", "dpe_host_read_grid")

_lib = None


def load_library():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError("%s not found: run `make -C navlab-dpe-sdr_b200/host`" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        lib.dpe_shell_create.restype = C.c_void_p
        lib.dpe_shell_destroy.argtypes = [C.c_void_p]
        lib.dpe_shell_exec.argtypes = [C.c_void_p, C.c_char_p]
        lib.dpe_shell_run_blocking.argtypes = [C.c_void_p, C.c_char_p, C.c_long]
        lib.dpe_shell_flow_stats.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        lib.dpe_shell_read_port.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_long]
        lib.dpe_shell_read_port.restype = C.c_long
        lib.dpe_host_sat_position.argtypes = [C.c_char_p, C.c_int, C.c_double, C.c_void_p]
        lib.dpe_host_make_grid.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_long]
        lib.dpe_host_read_handoff.argtypes = [C.c_char_p, C.c_void_p, C.c_long]
        lib.dpe_host_read_handoff.restype = C.c_long
        lib.dpe_host_read_grid.argtypes = [C.c_char_p, C.c_void_p, C.c_long]
        lib.dpe_host_read_grid.restype = C.c_long
        _lib = lib
    return _lib


class Shell:
    """One console session (``dpe_shell``)."""

    def __init__(self):
        self.lib = load_library()
        self.h = C.c_void_p(self.lib.dpe_shell_create())

    def close(self):
        if self.h:
            self.lib.dpe_shell_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def exec(self, line: str) -> int:
        return int(self.lib.dpe_shell_exec(self.h, line.encode()))

    def run_blocking(self, flow: str, max_epochs: int = -1) -> int:
        return int(self.lib.dpe_shell_run_blocking(self.h, flow.encode(), max_epochs))

    def stats(self, flow: str) -> dict:
        a = np.zeros(5)
        if self.lib.dpe_shell_flow_stats(self.h, flow.encode(), a.ctypes.data):
            raise RuntimeError("no such flow")
        return dict(run_count=int(a[0]), avg_us=a[1], min_us=a[2], max_us=a[3], total_s=a[4])

    def read_port(self, flow: str, module: str, port: str, cap: int = 1 << 16) -> np.ndarray:
        out = np.zeros(cap)
        n = self.lib.dpe_shell_read_port(self.h, flow.encode(), module.encode(), port.encode(), out.ctypes.data, cap)
        if n < 0:
            raise RuntimeError("port %s.%s is not a readable HOST port" % (module, port))
        return out[:n].copy()


def sat_position(rinex: str, prn: int, tx_time: float) -> np.ndarray:
    out = np.zeros(8)
    if load_library().dpe_host_sat_position(rinex.encode(), prn, tx_time, out.ctypes.data):
        raise RuntimeError("no ephemeris")
    return out


def make_grid(dims, spacing, grid_type=0) -> np.ndarray:
    d = np.asarray(dims, dtype=np.int32)
    s = np.asarray(spacing, dtype=np.float64)
    n = int(np.prod(d))
    out = np.zeros(n * 4)
    if load_library().dpe_host_make_grid(d.ctypes.data, s.ctypes.data, grid_type, out.ctypes.data, out.size) != n:
        raise RuntimeError("make_grid failed")
    return out.reshape(n, 4)


def read_handoff(path: str) -> dict:
    """The host flow's handoff-CSV reader (DPInit), as a dict of arrays."""
    out = np.zeros(12 + 8 * 64)
    n = load_library().dpe_host_read_handoff(path.encode(), out.ctypes.data, out.size)
    if n < 0:
        raise RuntimeError("handoff file not readable")
    c = int(out[0])
    rows = out[12:12 + 8 * c].reshape(8, c)
    keys = ("prn", "rc", "ri", "fc", "fi", "cp", "cp_timestamp", "TOW")
    d = dict(rxTime=out[1], bytes_read=int(out[2]), t_oe=int(out[3]), X_ECEF=out[4:12].copy())
    d.update({k: rows[i].copy() for i, k in enumerate(keys)})
    return d


def read_grid(path: str, cap: int = 1 << 22) -> np.ndarray:
    """The host flow's grid-CSV reader (BatchCorrManifold LoadPosGrid)."""
    out = np.zeros(4 * cap)
    n = load_library().dpe_host_read_grid(path.encode(), out.ctypes.data, out.size)
    if n < 0:
        raise RuntimeError("grid file not readable")
    return out[:4 * n].reshape(n, 4).copy()
