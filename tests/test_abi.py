"""The C-ABI library loads and exports every symbol include/hmrt.h declares (no compute calls)."""
import ctypes as C
import re
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (REPO / "include" / "hmrt.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hmrt_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ["hmrt_create", "hmrt_set_heightmap", "hmrt_trace", "hmrt_trace_host", "hmrt_clear_heightmap",
                 "hmrt_scatter_las", "hmrt_scatter_xyz", "hmrt_build_mips", "hmrt_resolve_colors"]:
        assert must in syms


def test_library_exports_every_declared_symbol():
    from hmrt import _abi

    lib = _abi.load()  # raises if csrc/libhmrt.so is missing: there is no fallback
    for name in declared_symbols():
        assert hasattr(lib, name), f"libhmrt.so does not export {name}"
    assert set(declared_symbols()) == set(_abi.PROTOTYPES), "hmrt/_abi.py prototypes out of sync with include/hmrt.h"


def test_struct_layouts_match_header():
    from hmrt import _abi

    assert C.sizeof(_abi.Color) == 3
    assert C.sizeof(_abi.Camera) == 36
    assert C.sizeof(_abi.Hit) == 16
    assert C.sizeof(_abi.TraceOpts) == 40
    assert C.sizeof(_abi.LasTransform) == 96


def test_host_only_entry_points():
    """Entry points that need no device: layout tables (main.cpp:995-1003), tile bookkeeping, errors."""
    from hmrt import _abi

    lib = _abi.load()
    assert lib.hmrt_version() == 201
    res = (C.c_int * 8)()
    idx = (C.c_int64 * 8)()
    total = C.c_int64()
    assert lib.hmrt_pyramid_layout(32, 8, res, idx, C.byref(total)) == 0
    assert list(res) == [4096, 2048, 1024, 512, 256, 128, 64, 32]
    assert idx[7] == 0 and idx[6] == 32 * 32 and total.value == 32 * 32 * 21845  # stride_x = 21845
    assert lib.hmrt_pyramid_layout(0, 8, None, None, None) == _abi.E_ARG
    assert lib.hmrt_pyramid_layout(4096, 8, None, None, None) == _abi.E_SHAPE  # > 2^32 cells
    assert lib.hmrt_rows_local(2160, 0, 1) == 2160
    assert sum(lib.hmrt_rows_local(2160, r, 8) for r in range(8)) == 2160
    assert sum(lib.hmrt_rows_local(1083, r, 4) for r in range(4)) == 1083
    assert b"argument" in lib.hmrt_error_string(_abi.E_ARG)
    o = _abi.TraceOpts()
    lib.hmrt_trace_opts_default(C.byref(o), 12.5)
    assert o.max_height == 12.5 and o.tile_stride == 1 and o.shadows == 0 and o.full_frame_output == 0


def test_no_device_means_loud_failure():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import hmrt

    with pytest.raises(RuntimeError):
        hmrt.Context(0)
