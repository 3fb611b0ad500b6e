"""Test-side access to the CPU checkers under oracle/ (TEST INFRASTRUCTURE ONLY).

`oracle()`  -> oracle/_build/libhmrt_oracle.so  (plain-C restatement, hmrt_oracle.c)
`ref()`     -> oracle/_ref/libhmrt_ref.so       (the reference's own code, host-compiled)
`ref_dpow()`-> oracle/_ref/libhmrt_ref_dpow.so  (same, Linux double-pow/floor meaning)
Nothing here is imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "gpu-heightmap-raytracer_b200"))

from hmrt._abi import Camera, Color, Hit, LasTransform, TraceOpts  # noqa: E402  (struct layouts only)

ORACLE_SO = REPO / "oracle" / "_build" / "libhmrt_oracle.so"
REF_SO = REPO / "oracle" / "_ref" / "libhmrt_ref.so"
REF_DPOW_SO = REPO / "oracle" / "_ref" / "libhmrt_ref_dpow.so"

_P = C.c_void_p
_TRACE_ARGS = [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Camera), C.POINTER(TraceOpts),
               C.c_int, C.c_int, C.c_int, _P, _P]
_cache = {}


def build_oracle() -> None:
    subprocess.run(["make", "-s", "-C", str(REPO / "oracle"), "_build/libhmrt_oracle.so"], check=True)


def oracle() -> C.CDLL:
    if "oracle" not in _cache:
        if not ORACLE_SO.exists() or ORACLE_SO.stat().st_mtime < (REPO / "oracle" / "hmrt_oracle.c").stat().st_mtime:
            build_oracle()
        lib = C.CDLL(str(ORACLE_SO))
        lib.hmrt_oracle_trace.restype = C.c_int
        lib.hmrt_oracle_trace.argtypes = _TRACE_ARGS
        lib.hmrt_oracle_pyramid_layout.restype = C.c_int
        lib.hmrt_oracle_pyramid_layout.argtypes = [C.c_int, C.c_int, _P, _P, _P]
        lib.hmrt_oracle_rasterise_las.restype = C.c_int
        lib.hmrt_oracle_rasterise_las.argtypes = [_P, C.c_int64, C.c_int, C.c_int, C.POINTER(LasTransform), _P, C.c_int, C.c_int, _P]
        lib.hmrt_oracle_rasterise_xyz.restype = C.c_int
        lib.hmrt_oracle_rasterise_xyz.argtypes = [_P, C.c_int64, C.POINTER(LasTransform), _P, C.c_int, C.c_int]
        lib.hmrt_oracle_build_mips.restype = C.c_int
        lib.hmrt_oracle_build_mips.argtypes = [_P, C.c_int, C.c_int]
        lib.hmrt_oracle_pdg_generate.restype = C.c_int
        lib.hmrt_oracle_pdg_generate.argtypes = [C.c_int, C.c_uint64, _P]
        lib.hmrt_oracle_window_place.restype = C.c_int
        lib.hmrt_oracle_window_place.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, _P]
        lib.hmrt_oracle_compose_window.restype = C.c_int
        lib.hmrt_oracle_compose_window.argtypes = [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]
        _cache["oracle"] = lib
    return _cache["oracle"]


def _load_ref(path: Path, key: str):
    if key not in _cache:
        if not path.exists():
            if (Path("/root/reference") / "GPUHeightmapRaytracer").exists():
                subprocess.run(["bash", str(REPO / "oracle" / "build_ref.sh")], check=True)
        if not path.exists():
            _cache[key] = None
        else:
            lib = C.CDLL(str(path))
            lib.hmrt_ref_trace.restype = C.c_int
            lib.hmrt_ref_trace.argtypes = _TRACE_ARGS
            lib.hmrt_ref_float_math.restype = C.c_int
            _cache[key] = lib
    return _cache[key]


def ref():
    return _load_ref(REF_SO, "ref")


def ref_dpow():
    return _load_ref(REF_DPOW_SO, "ref_dpow")


# ---------------------------------------------------------------------------------------------
# numpy helpers (host logic of the tests; not an implementation of the traced path)

def pyramid_layout(coarse_res: int, levels: int):
    """res[i], idx[i], total -- main.cpp:995-1003."""
    res = [0] * levels
    idx = [0] * levels
    res[levels - 1] = coarse_res
    for i in range(levels - 2, -1, -1):
        idx[i] = idx[i + 1] + res[i + 1] * res[i + 1]
        res[i] = res[i + 1] * 2
    return res, idx, idx[0] + res[0] * res[0]


def pyramid_from_finest(finest: np.ndarray, levels: int) -> np.ndarray:
    """Max-pyramid in the reference layout (coarsest first) from a [R0,R0] float32 grid
    indexed [z, x]; 'level i+1 = max of 2x2 children' (fixed point of main.cpp:227-233)."""
    r0 = finest.shape[0]
    assert finest.shape == (r0, r0) and r0 % (1 << (levels - 1)) == 0
    coarse = r0 >> (levels - 1)
    res, idx, total = pyramid_layout(coarse, levels)
    out = np.zeros(total, dtype=np.float32)
    cur = np.ascontiguousarray(finest, dtype=np.float32)
    out[idx[0]:idx[0] + r0 * r0] = cur.ravel()
    for l in range(1, levels):
        r = res[l]
        cur = cur.reshape(r, 2, r, 2).max(axis=(1, 3))
        out[idx[l]:idx[l] + r * r] = cur.ravel()
    return out


def sines_terrain(r0: int, seed: int = 0) -> np.ndarray:
    """Analytic terrain of SURVEY.md Appendix A (computed in float64, rounded once) + a seeded
    small-scale perturbation so neighbouring cells differ.  [z, x] float32, >= 0."""
    x = np.arange(r0, dtype=np.float64)[None, :]
    z = np.arange(r0, dtype=np.float64)[:, None]
    h = 40 + 25 * np.sin(0.013 * x) * np.cos(0.017 * z) + 10 * np.sin(0.11 * x + 0.07 * z) + 3 * np.sin(0.9 * x) * np.sin(0.8 * z)
    rng = np.random.default_rng(seed)
    h = h + rng.random((r0, r0)) * 0.5
    return np.maximum(h, 0).astype(np.float32)


def make_camera(position, forward, frame_dim=(32.0, 18.0, 20.0)) -> Camera:
    f = np.asarray(forward, dtype=np.float64)
    f = (f / np.linalg.norm(f)).astype(np.float32)
    cam = Camera()
    cam.frame_dim[:] = [float(v) for v in frame_dim]
    cam.forward[:] = [float(v) for v in f]
    cam.position[:] = [float(np.float32(v)) for v in position]
    return cam


def make_opts(max_height: float, use_color_map=False, shadows=False, light_dir=(0.3, 0.8, 0.52), shadow_bias=0.0,
              tile_first=0, tile_stride=1) -> TraceOpts:
    o = TraceOpts()
    o.use_color_map = int(use_color_map)
    o.max_height = float(np.float32(max_height))
    o.shadows = int(shadows)
    l = np.asarray(light_dir, dtype=np.float64)
    l = (l / np.linalg.norm(l)).astype(np.float32)
    o.light_dir[:] = [float(v) for v in l]
    o.shadow_bias = float(shadow_bias)
    o.tile_first = tile_first
    o.tile_stride = tile_stride
    return o


def cpu_trace(lib_fn, pyramid: np.ndarray, color_map, coarse_res: int, levels: int, W: int, H: int, cam: Camera,
              opts: TraceOpts, n_threads: int | None = None, rows=None, want_hits=True):
    """Run hmrt_oracle_trace / hmrt_ref_trace.  Returns (rgb[H,W,3] u8, hits structured array or None)."""
    n_threads = n_threads or os.cpu_count() or 1
    rgb = np.zeros((H, W, 3), dtype=np.uint8)
    hits = np.zeros((H, W), dtype=hit_dtype) if want_hits else None
    r0, r1 = rows if rows is not None else (0, H)
    cm_ptr = color_map.ctypes.data if color_map is not None else None
    rc = lib_fn(pyramid.ctypes.data, cm_ptr, coarse_res, levels, W, H, C.byref(cam), C.byref(opts), n_threads, r0, r1,
                rgb.ctypes.data, hits.ctypes.data if hits is not None else None)
    if rc != 0:
        raise RuntimeError(f"cpu trace failed: {rc}")
    return rgb, hits


hit_dtype = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("flags", "<u4")])
assert hit_dtype.itemsize == C.sizeof(Hit)


def hit_cells(hits: np.ndarray, r0: int):
    """Un-mirrored finest hit cell per pixel (-1 = miss), and hit mask; SURVEY.md section 8(a) R4."""
    flags = hits["flags"]
    hit = (flags & 1) != 0
    cx = np.floor(hits["x"]).astype(np.int64)
    cz = np.floor(hits["z"]).astype(np.int64)
    cx = np.where((flags & 2) != 0, r0 - 1 - cx, cx)
    cz = np.where((flags & 4) != 0, r0 - 1 - cz, cz)
    cell = np.where(hit, cx + cz * r0, -1)
    return cell, hit


def steps_of(hits: np.ndarray) -> np.ndarray:
    return hits["flags"] >> 8


# ---------------------------------------------------------------------------------------------
# hostsim: the product's ray_core.cuh compiled for the host (tests/hostsim/hostsim.cpp)

HOSTSIM_SO = REPO / "tests" / "_build" / "libhmrt_hostsim.so"


def hostsim() -> C.CDLL:
    if "hostsim" not in _cache:
        src = REPO / "tests" / "hostsim" / "hostsim.cpp"
        core = REPO / "gpu-heightmap-raytracer_b200" / "csrc" / "ray_core.cuh"
        if not HOSTSIM_SO.exists() or HOSTSIM_SO.stat().st_mtime < max(src.stat().st_mtime, core.stat().st_mtime):
            HOSTSIM_SO.parent.mkdir(parents=True, exist_ok=True)
            subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-shared", "-o", str(HOSTSIM_SO),
                            str(src), "-lpthread"], check=True)
        lib = C.CDLL(str(HOSTSIM_SO))
        lib.hostsim_trace.restype = C.c_int
        lib.hostsim_trace.argtypes = _TRACE_ARGS
        _cache["hostsim"] = lib
    return _cache["hostsim"]


def assert_same_trace(a, b, what=""):
    """Bit-exact comparison of two (rgb, hits) results; steps compared only if both sides count them."""
    rgb_a, hit_a = a
    rgb_b, hit_b = b
    bad = int((rgb_a != rgb_b).any(axis=-1).sum())
    assert bad == 0, f"{what}: {bad} pixels differ in colour"
    if hit_a is not None and hit_b is not None:
        for f in "xyz":
            # compare bit patterns so that NaN == NaN
            assert (hit_a[f].view(np.uint32) == hit_b[f].view(np.uint32)).all(), f"{what}: hit.{f} differs"
        fa, fb = hit_a["flags"], hit_b["flags"]
        assert ((fa & 0xFF) == (fb & 0xFF)).all(), f"{what}: hit flags differ"
        if (fa >> 8).any() and (fb >> 8).any():
            assert ((fa >> 8) == (fb >> 8)).all(), f"{what}: step counts differ"


SCENES = {
    # name: (R0, levels)
    "r1024_l8": (1024, 8),
    "r512_l4": (512, 4),
    "r256_l1": (256, 1),
    "r768_l7": (768, 7),   # coarse_res 12: not a power of two
}


def scene(name: str, seed: int = 0):
    r0, levels = SCENES[name]
    fin = sines_terrain(r0, seed)
    pyr = pyramid_from_finest(fin, levels)
    rng = np.random.default_rng(seed + 1)
    cmap = rng.integers(0, 256, (r0, r0, 3), dtype=np.uint8)
    return dict(r0=r0, levels=levels, coarse=r0 >> (levels - 1), finest=fin, pyramid=pyr, color_map=cmap,
                max_height=float(fin.max()))


def cameras_for(sc, n=4):
    """Deterministic camera set: down-looking, oblique, grazing and upward, four headings."""
    r0, mh = sc["r0"], sc["max_height"]
    poses = [(-0.9, 1.5, (0.0, 1.0)), (-0.3, 1.5, (1.0, 0.3)), (-0.1, 1.2, (-0.7, -0.5)), (0.2, 0.5, (0.4, -1.0)),
             (-0.5, 2.5, (-1.0, 0.05)), (-0.05, 1.05, (0.6, 0.8))]
    return [make_camera((r0 / 2 + 0.37, k * mh, r0 / 2 - 3.21), (hx, pitch, hz)) for pitch, k, (hx, hz) in poses[:n]]


def load_golden():
    """tests/golden/ray_golden.npz -> dict with the pyramid rebuilt from the stored finest level."""
    g = dict(np.load(REPO / "tests" / "golden" / "ray_golden.npz"))
    levels = int(g["levels"])
    g["pyramid"] = pyramid_from_finest(g["finest"], levels)
    g["r0"] = g["finest"].shape[0]
    g["coarse"] = g["r0"] >> (levels - 1)
    cams = []
    for row in g["cameras"]:
        c = Camera()
        c.frame_dim[:] = [float(v) for v in row[0:3]]
        c.forward[:] = [float(v) for v in row[3:6]]
        c.position[:] = [float(v) for v in row[6:9]]
        cams.append(c)
    g["cams"] = cams
    return g


GOLDEN_MODES = {"ramp": (False, False), "shadow": (False, True), "cmap": (True, False)}


def golden_opts(g, mode):
    uc, sh = GOLDEN_MODES[mode]
    o = make_opts(float(g["max_height"]), use_color_map=uc, shadows=sh)
    o.light_dir[:] = [float(v) for v in g["light_dir"]]
    return o


def golden_expected(g, mode, i):
    hmode = "ramp" if mode == "cmap" else mode  # the walk does not depend on the colouring mode
    hits = np.ascontiguousarray(g[f"hits_{hmode}_{i}"]).view(hit_dtype).reshape(int(g["H"]), int(g["W"]))
    return g[f"rgb_{mode}_{i}"], hits


# ---- camera window over the section grid (preparePointBuffer, main.cpp:459-618) ------------------------------------

def oracle_window_place(camera_position, origins, grid, coarse_res, levels):
    """(rc, WindowPlacement) from the restatement of main.cpp:461-516; origins[i][j] = (x, y)."""
    from hmrt._abi import WindowPlacement

    cam = (C.c_float * 3)(*[float(v) for v in camera_position])
    org = np.ascontiguousarray(np.asarray(origins, dtype=np.float32).reshape(grid, grid, 2))
    out = WindowPlacement()
    rc = oracle().hmrt_oracle_window_place(cam, org.ctypes.data_as(C.POINTER(C.c_float)), grid, coarse_res, levels, C.byref(out))
    return rc, out


def oracle_compose_window(pyramids, colors, coarse_res, levels, cell_x, cell_y):
    """The four memcpy loops per level of main.cpp:519-618 on host arrays.  pyramids / colors: [[lb, lt], [rb, rt]]
    (x-major: [0][*] = left sections, [*][0] = bottom sections); colors may be None."""
    res, idx, total = pyramid_layout(coarse_res, levels)
    ptrs = (C.c_void_p * 4)(*[pyramids[a][b].ctypes.data for a in range(2) for b in range(2)])
    out = np.full(total, np.nan, np.float32)
    out_c, cptrs = None, None
    if colors is not None:
        cptrs = (C.c_void_p * 4)(*[colors[a][b].ctypes.data for a in range(2) for b in range(2)])
        out_c = np.full((res[0], res[0], 3), 7, np.uint8)
    rc = oracle().hmrt_oracle_compose_window(ptrs, cptrs, coarse_res, levels, cell_x, cell_y, out.ctypes.data,
                                             out_c.ctypes.data if out_c is not None else None)
    assert rc == 0
    return out, out_c
