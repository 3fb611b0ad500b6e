"""ctypes view of include/hmrt.h (struct layouts, prototypes) and the loader of libhmrt.so.

The library is the product: there is no Python/CPU fallback.  `load()` raises if the CUDA
shared object has not been built (run `python -c "import __graft_entry__ as g; g.build()"`).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG_ROOT = Path(__file__).resolve().parent.parent
LIB_PATH = PKG_ROOT / "csrc" / "libhmrt.so"

HMRT_MAX_LEVELS = 16
HMRT_ROW_TILE = 8

HIT_HIT = 1
HIT_MIRROR_X = 2
HIT_MIRROR_Z = 4
HIT_SHADOWED = 8
HIT_STEPS_SHIFT = 8

E_ARG, E_STATE, E_SHAPE, E_NOMEM, E_NCCL = -1, -2, -3, -4, -5


class Color(C.Structure):  # hmrt_color == CudaSpace::Color (CudaKernel.cuh:37-48)
    _pack_ = 1
    _fields_ = [("r", C.c_uint8), ("g", C.c_uint8), ("b", C.c_uint8)]


class Camera(C.Structure):  # hmrt_camera: per-frame args of CudaSpace::rayTrace (CudaKernel.cuh:49)
    _fields_ = [("frame_dim", C.c_float * 3), ("forward", C.c_float * 3), ("position", C.c_float * 3)]


class Hit(C.Structure):  # hmrt_hit
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("flags", C.c_uint32)]


class TraceOpts(C.Structure):  # hmrt_trace_opts
    _fields_ = [
        ("use_color_map", C.c_int),
        ("max_height", C.c_float),
        ("shadows", C.c_int),
        ("light_dir", C.c_float * 3),
        ("shadow_bias", C.c_float),
        ("tile_first", C.c_int),
        ("tile_stride", C.c_int),
        ("full_frame_output", C.c_int),
    ]


class LasTransform(C.Structure):  # hmrt_las_transform
    _fields_ = [
        ("scale", C.c_double * 3),
        ("offset", C.c_double * 3),
        ("min", C.c_double * 3),
        ("cell_size", C.c_float * 3),
        ("origin", C.c_float * 2),
    ]


class WindowPlacement(C.Structure):  # hmrt_window_placement (preparePointBuffer's host arithmetic, main.cpp:461-516)
    _fields_ = [("min_x", C.c_int), ("min_y", C.c_int), ("max_x", C.c_int), ("max_y", C.c_int),
                ("cell_x", C.c_int), ("cell_y", C.c_int), ("camera", C.c_float * 3)]


class WindowSections(C.Structure):  # hmrt_window_sections: [x][y], x 0 = left, y 0 = bottom
    _fields_ = [("d_pyramid", (C.c_void_p * 2) * 2), ("d_color_map", (C.c_void_p * 2) * 2)]


# every symbol include/hmrt.h declares: name -> (restype, argtypes)
_P = C.c_void_p
PROTOTYPES = {
    "hmrt_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "hmrt_destroy": (C.c_int, [_P]),
    "hmrt_set_stream": (C.c_int, [_P, _P]),
    "hmrt_get_stream": (_P, [_P]),
    "hmrt_synchronize": (C.c_int, [_P]),
    "hmrt_error_string": (C.c_char_p, [C.c_int]),
    "hmrt_version": (C.c_int, []),
    "hmrt_pyramid_layout": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "hmrt_set_heightmap": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_float]),
    "hmrt_trace_opts_default": (None, [C.POINTER(TraceOpts), C.c_float]),
    "hmrt_trace": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(Camera), C.c_int, C.POINTER(TraceOpts), _P, _P]),
    "hmrt_trace_host": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(Camera), C.c_int, C.POINTER(TraceOpts), _P]),
    "hmrt_trace_host_begin": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(Camera), C.c_int, C.POINTER(TraceOpts), _P]),
    "hmrt_trace_host_wait": (C.c_int, [_P]),
    "hmrt_rows_local": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "hmrt_clear_heightmap": (C.c_int, [_P]),
    "hmrt_clear_section": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P]),
    "hmrt_scatter_las": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, C.POINTER(LasTransform), C.c_int64, _P, C.c_int, C.c_int, _P]),
    "hmrt_set_scatter_mode": (C.c_int, [_P, C.c_int]),
    "hmrt_scatter_xyz": (C.c_int, [_P, _P, C.c_int64, C.POINTER(LasTransform), _P, C.c_int, C.c_int]),
    "hmrt_build_mips": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "hmrt_resolve_colors": (C.c_int, [_P, _P, _P, C.c_int64]),
    "hmrt_window_place": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.POINTER(WindowPlacement)]),
    "hmrt_compose_window": (C.c_int, [_P, C.POINTER(WindowSections), C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "hmrt_broadcast_heightmap": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int]),
    "hmrt_allreduce_max_heights": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int]),
    "hmrt_copy_tiles_to_frames": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "hmrt_ipc_alloc": (C.c_int, [_P, C.c_size_t, C.POINTER(_P), _P]),
    "hmrt_ipc_free": (C.c_int, [_P, _P]),
    "hmrt_ipc_open": (C.c_int, [_P, _P, C.POINTER(_P)]),
    "hmrt_ipc_close": (C.c_int, [_P, _P]),
    "hmrt_rx_create": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.POINTER(_P)]),
    "hmrt_rx_destroy": (C.c_int, [_P]),
    "hmrt_rx_region_bytes": (C.c_size_t, [_P]),
    "hmrt_rx_export": (C.c_int, [_P, _P]),
    "hmrt_rx_connect": (C.c_int, [_P, _P]),
    "hmrt_rx_begin": (C.c_int, [_P]),
    "hmrt_rx_bin": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, C.POINTER(LasTransform)]),
    "hmrt_rx_barrier": (C.c_int, [_P]),
    "hmrt_rx_apply": (C.c_int, [_P]),
    "hmrt_rx_gather_mips": (C.c_int, [_P, _P]),
    "hmrt_rx_status": (C.c_int, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "hmrt_rx_bands": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "hmrt_set_trace_variant": (C.c_int, [_P, C.c_int]),
    "hmrt_set_host_variant": (C.c_int, [_P, C.c_int]),
    "hmrt_set_window_variant": (C.c_int, [_P, C.c_int]),
    "hmrt_set_l2_persist": (C.c_int, [_P, C.c_int, C.c_float]),
    "hmrt_trace_stats": (C.c_int, [_P, C.POINTER(C.c_uint64), C.c_int]),
    "hmrt_launch_count": (C.c_int64, [_P]),
}

_lib = None


class HmrtError(RuntimeError):
    def __init__(self, code: int, what: str):
        self.code = code
        super().__init__(f"{what}: hmrt error {code} ({error_string(code)})")


def load() -> C.CDLL:
    """Load csrc/libhmrt.so (the CUDA product library).  No fallback: raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("HMRT_LIB", LIB_PATH))
    if not path.exists():
        raise RuntimeError(
            f"{path} is missing: the CUDA library is the only implementation of this path "
            "(no CPU fallback). Build it with __graft_entry__.build() or `make -C "
            f"{PKG_ROOT / 'csrc'}`."
        )
    lib = C.CDLL(str(path))
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export the ABI
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def error_string(code: int) -> str:
    try:
        return load().hmrt_error_string(code).decode()
    except Exception:  # pragma: no cover
        return "?"


def check(code: int, what: str) -> None:
    if code != 0:
        raise HmrtError(code, what)
