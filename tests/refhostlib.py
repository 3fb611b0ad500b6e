"""Test-side access to oracle/_ref/libhmrt_refhost*.so (TEST INFRASTRUCTURE ONLY): the reference's own host-side code --
main.cpp:44-618 (globals, readLASHeader, loadLASToSection, the section manager, preparePointBuffer), :745-781 (camera) and
:995-1003 (pyramid tables) -- cut verbatim by oracle/build_ref.sh and compiled against stub liblas / thread / ifstream
types (oracle/refhost_harness.cpp).  The library keeps the reference's namespace-scope globals: one user at a time."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
_P = C.c_void_p
_cache = {}


def refhost(variant: str = ""):
    """variant "" = verbatim (4 x 4 sections, 8 levels); "_g3l4" = the two compile-time constants patched to 3 and 4."""
    key = "refhost" + variant
    if key not in _cache:
        path = REPO / "oracle" / "_ref" / f"libhmrt_refhost{variant}.so"
        if not path.exists() and (Path("/root/reference") / "GPUHeightmapRaytracer").exists():
            subprocess.run(["bash", str(REPO / "oracle" / "build_ref.sh")], check=True)
        if not path.exists():
            _cache[key] = None
        else:
            lib = C.CDLL(str(path))
            lib.hmrt_refhost_config.argtypes = [C.c_int, _P, _P, _P]
            lib.hmrt_refhost_set_las.argtypes = [_P, C.c_int64, C.c_int, C.c_int, _P, _P, _P, _P]
            lib.hmrt_refhost_set_cell_size.argtypes = [C.c_float] * 3
            lib.hmrt_refhost_set_cell_size.restype = None
            lib.hmrt_refhost_read_header.argtypes = [_P] * 4
            lib.hmrt_refhost_rasterise.argtypes = [_P] * 3
            lib.hmrt_refhost_init_sections.argtypes = [_P] * 5
            lib.hmrt_refhost_manage.argtypes = [_P] * 5
            lib.hmrt_refhost_fill_section.argtypes = [C.c_int, C.c_int, _P, _P]
            lib.hmrt_refhost_read_section.argtypes = [C.c_int, C.c_int, _P, _P]
            lib.hmrt_refhost_prepare.argtypes = [_P] * 4
            lib.hmrt_refhost_move_camera.argtypes = [_P, _P, _P, C.c_float, _P, C.c_float]
            lib.hmrt_refhost_rotate_camera.argtypes = [_P, C.c_float, C.c_float, C.c_float]
            _cache[key] = RefHost(lib)
    return _cache[key]


class RefHost:
    def __init__(self, lib):
        self.lib = lib
        self.levels = lib.hmrt_refhost_levels()
        self.grid = lib.hmrt_refhost_grid()
        self.coarse = None
        self._keep = None

    def config(self, coarse):
        res = np.zeros(self.levels, np.int32)
        idx = np.zeros(self.levels, np.int64)
        total = C.c_int64(0)
        assert self.lib.hmrt_refhost_config(coarse, res.ctypes.data, idx.ctypes.data, C.byref(total)) == 0
        self.coarse, self.res, self.idx, self.total = coarse, [int(r) for r in res], [int(i) for i in idx], int(total.value)
        return self.res, self.idx, self.total

    def set_las(self, hdr, rec, cell=(2.0, 2.0, 2.0)):
        rec = np.ascontiguousarray(rec, np.uint8)
        self._keep = rec
        a = [np.array(v, np.float64) for v in (hdr.scale, hdr.offset, hdr.min, hdr.max)]
        assert self.lib.hmrt_refhost_set_las(rec.ctypes.data, rec.shape[0], rec.shape[1], hdr.point_format, *[v.ctypes.data for v in a]) == 0
        self.lib.hmrt_refhost_set_cell_size(*[float(np.float32(c)) for c in cell])

    def read_header(self):
        cam, bounds, mh, cell = np.zeros(3, np.float32), np.zeros(2, np.float32), C.c_float(0), np.zeros(3, np.float32)
        assert self.lib.hmrt_refhost_read_header(cam.ctypes.data, bounds.ctypes.data, C.byref(mh), cell.ctypes.data) == 0
        return cam, bounds, np.float32(mh.value), cell

    def rasterise(self, origin=(0.0, 0.0)):
        """allocateSection's zero-initialised section (main.cpp:259-260) through loadLASToSection (:174-238)."""
        pyr = np.empty(self.total, np.float32)
        col = np.empty((self.res[0], self.res[0], 3), np.uint8)
        org = np.array(origin, np.float32)
        assert self.lib.hmrt_refhost_rasterise(org.ctypes.data, pyr.ctypes.data, col.ctypes.data) == 0
        return pyr, col

    def _sections_call(self, fn, cam):
        g = self.grid
        cam = np.array(cam, np.float32)
        origins = np.zeros((g, g, 2), np.float32)
        slots = np.zeros((g * g, 2), np.int32)
        lorg = np.zeros((g * g, 2), np.float32)
        n = C.c_int(0)
        assert fn(cam.ctypes.data, origins.ctypes.data, slots.ctypes.data, lorg.ctypes.data, C.byref(n)) == 0
        return origins, slots[:n.value].copy(), lorg[:n.value].copy()

    def init_sections(self, cam):
        return self._sections_call(self.lib.hmrt_refhost_init_sections, cam)

    def manage(self, cam):
        return self._sections_call(self.lib.hmrt_refhost_manage, cam)

    def fill_section(self, i, j, pyr=None, col=None):
        if pyr is not None:
            pyr = np.ascontiguousarray(pyr, np.float32)
            assert pyr.size == self.total
        if col is not None:
            col = np.ascontiguousarray(col, np.uint8)
        assert self.lib.hmrt_refhost_fill_section(i, j, pyr.ctypes.data if pyr is not None else None, col.ctypes.data if col is not None else None) == 0

    def read_section(self, i, j):
        pyr = np.empty(self.total, np.float32)
        col = np.empty((self.res[0], self.res[0], 3), np.uint8)
        assert self.lib.hmrt_refhost_read_section(i, j, pyr.ctypes.data, col.ctypes.data) == 0
        return pyr, col

    def prepare(self, cam, want_buffers=True):
        cam = np.array(cam, np.float32)
        cpb = np.zeros(3, np.float32)
        pyr = np.empty(self.total, np.float32) if want_buffers else None
        col = np.empty((self.res[0], self.res[0], 3), np.uint8) if want_buffers else None
        assert self.lib.hmrt_refhost_prepare(cam.ctypes.data, cpb.ctypes.data, pyr.ctypes.data if want_buffers else None,
                                             col.ctypes.data if want_buffers else None) == 0
        return cpb, pyr, col

    def move_camera(self, position, forward, move_rht_up_fwd, dt, bounds, max_height):
        p = np.array(position, np.float32)
        f = np.array(forward, np.float32)
        m = np.array(move_rht_up_fwd, np.float32)
        b = np.array(bounds, np.float32)
        assert self.lib.hmrt_refhost_move_camera(p.ctypes.data, f.ctypes.data, m.ctypes.data, float(dt), b.ctypes.data, float(max_height)) == 0
        return p

    def rotate_camera(self, forward, rot_up, rot_right, dt):
        f = np.array(forward, np.float32)
        assert self.lib.hmrt_refhost_rotate_camera(f.ctypes.data, float(rot_up), float(rot_right), float(dt)) == 0
        return f
