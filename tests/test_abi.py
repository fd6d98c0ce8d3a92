"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports
every symbol include/dpe_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dpe_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dpe_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_documented_surface():
    names = _declared_functions()
    for must in ("dpe_ctx_create", "dpe_block_stage", "dpe_replica_prepare", "dpe_correlogram",
                 "dpe_score_pos", "dpe_estimate", "dpe_epoch_run"):
        assert must in names


def test_library_exports_every_declared_symbol(capi):
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(lib, name), "libdpe_b200.so does not export %s" % name
    assert set(capi.EXPORTS) == set(_declared_functions())
    lib.dpe_abi_version.restype = ctypes.c_int
    assert lib.dpe_abi_version() == capi.DPE_ABI_VERSION


def test_struct_layouts_match_header(capi, tmp_path):
    """Compile the real header with gcc and compare sizes / key offsets with the ctypes mirror."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "dpe_b200.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(dpe_cfg), sizeof(dpe_epoch),'
        ' sizeof(dpe_result), offsetof(dpe_cfg, Gv), offsetof(dpe_epoch, rc_end), offsetof(dpe_epoch, rx_time),'
        ' offsetof(dpe_result, argmax));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(capi.DpeCfg), ctypes.sizeof(capi.DpeEpoch), ctypes.sizeof(capi.DpeResult),
            capi.DpeCfg.Gv.offset, capi.DpeEpoch.rc_end.offset, capi.DpeEpoch.rx_time.offset,
            capi.DpeResult.argmax.offset]
    assert got == want


def test_no_cpu_fallback_without_gpu(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.DpeError):
        capi.Context(fs=2.5e6, S=50000, max_chan=8, G=81)


def test_missing_library_fails_loudly(capi, tmp_path):
    with pytest.raises(FileNotFoundError):
        capi.load_library(str(tmp_path / "libdpe_b200.so"))
