"""Section grid bookkeeping of the C++ host layer (hmrt_host::SectionLayout = initializeSections / manageSections /
rearrangeSections*, main.cpp:276-448) against an independent Python restatement of the same reference lines, along random
camera walks; plus the invariants the reference relies on (the window never leaves the grid after manageSections)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oraclelib as ol

REPO = Path(__file__).resolve().parent.parent
HOST = REPO / "gpu-heightmap-raytracer_b200" / "host"


@pytest.fixture(scope="module")
def hostlib():
    subprocess.run(["make", "-s", "-C", str(HOST), "libhmrt_host.so"], check=True)
    lib = C.CDLL(str(HOST / "libhmrt_host.so"))
    lib.hmrt_host_section_layout_step.argtypes = [C.c_int] * 4 + [C.c_void_p] * 5 + [C.c_int]
    return lib


def step(lib, init, grid, coarse, levels, cam, origins, tags):
    cam = np.asarray(cam, np.float32)
    loads = np.zeros((4 * grid, 3), np.int32)
    lorg = np.zeros((4 * grid, 2), np.float32)
    if init:
        loads = np.zeros((grid * grid, 3), np.int32)
        lorg = np.zeros((grid * grid, 2), np.float32)
    n = lib.hmrt_host_section_layout_step(int(init), grid, coarse, levels, cam.ctypes.data, origins.ctypes.data, tags.ctypes.data,
                                          loads.ctypes.data, lorg.ctypes.data, len(loads))
    assert n >= 0
    return loads[:n], lorg[:n]


def py_manage(origins, tags, cam, size):
    """manageSections (main.cpp:407-448) with rearrangeSectionsX/Y (:329-402); origins[i][j] = (x, y)."""
    g = origins.shape[0]
    f32 = np.float32
    loads = []
    if cam[0] < origins[1, 0, 0]:
        freed = tags[g - 1, :].copy()
        for i in range(g - 1, 0, -1):
            origins[i], tags[i] = origins[i - 1].copy(), tags[i - 1].copy()
        for j in range(g):
            origins[0, j] = (f32(origins[1, j, 0] - f32(size)), origins[1, j, 1])
            tags[0, j] = freed[j]
            loads.append((0, j, int(freed[j])))
    if cam[0] >= origins[g - 1, g - 1, 0]:
        freed = tags[0, :].copy()
        for i in range(g - 1):
            origins[i], tags[i] = origins[i + 1].copy(), tags[i + 1].copy()
        for j in range(g):
            origins[g - 1, j] = (f32(origins[g - 2, j, 0] + f32(size)), origins[g - 2, j, 1])
            tags[g - 1, j] = freed[j]
            loads.append((g - 1, j, int(freed[j])))
    if cam[2] < origins[0, 1, 1]:
        freed = tags[:, g - 1].copy()
        for j in range(g - 1, 0, -1):
            origins[:, j], tags[:, j] = origins[:, j - 1].copy(), tags[:, j - 1].copy()
        for i in range(g):
            origins[i, 0] = (origins[i, 1, 0], f32(origins[i, 1, 1] - f32(size)))
            tags[i, 0] = freed[i]
            loads.append((i, 0, int(freed[i])))
    if cam[2] >= origins[0, g - 1, 1]:
        freed = tags[:, 0].copy()
        for j in range(g - 1):
            origins[:, j], tags[:, j] = origins[:, j + 1].copy(), tags[:, j + 1].copy()
        for i in range(g):
            origins[i, g - 1] = (origins[i, g - 2, 0], f32(origins[i, g - 2, 1] + f32(size)))
            tags[i, g - 1] = freed[i]
            loads.append((i, g - 1, int(freed[i])))
    return loads


@pytest.mark.parametrize("grid,coarse,levels", [(4, 32, 8), (4, 8, 4), (3, 5, 3)])
def test_section_layout_follows_the_reference_rules(hostlib, grid, coarse, levels):
    size = coarse << (levels - 1)
    cam = np.array([5000.25, 40.0, -1234.5], np.float32)
    origins = np.zeros((grid, grid, 2), np.float32)
    tags = np.zeros((grid, grid), np.int32)
    loads, lorg = step(hostlib, True, grid, coarse, levels, cam, origins, tags)
    # initializeSections, main.cpp:276-288
    assert len(loads) == grid * grid and sorted(tags.ravel()) == list(range(grid * grid))
    for i in range(grid):
        for j in range(grid):
            assert origins[i, j, 0] == np.float32(cam[0] + np.float32((i - grid / 2.0) * size))
            assert origins[i, j, 1] == np.float32(cam[2] + np.float32((j - grid / 2.0) * size))
    rng = np.random.default_rng(grid * 100 + coarse)
    po, pt = origins.copy(), tags.copy()
    shifts = 0
    for k in range(400):
        cam = cam + np.array([rng.normal(0, size * 0.35), 0, rng.normal(0, size * 0.35)], np.float32)
        loads, lorg = step(hostlib, False, grid, coarse, levels, cam, origins, tags)
        want = py_manage(po, pt, cam, size)
        assert np.array_equal(origins.view(np.uint32), po.view(np.uint32)) and np.array_equal(tags, pt), k
        assert [tuple(int(v) for v in l) for l in loads] == want, k
        last = {int(t): o for (i, j, t), o in zip(loads, lorg)}  # a later shift of the same call may move or recycle a section
        for t, o in last.items():
            i, j = np.argwhere(tags == t)[0]
            assert np.array_equal(o, origins[i, j]), "a (re)filled section must be filled for the origin it ends up with"
        shifts += bool(len(loads))
        assert sorted(tags.ravel()) == list(range(grid * grid))  # section objects are recycled, never lost or duplicated
        # the origins stay a regular lattice
        assert np.all(np.diff(origins[:, :, 0], axis=0) == np.float32(size)) and np.all(np.diff(origins[:, :, 1], axis=1) == np.float32(size))
        if grid == 4 and abs(cam[0] - (origins[0, 0, 0] + 2 * size)) < size and abs(cam[2] - (origins[0, 0, 1] + 2 * size)) < size:
            # camera inside the inner 2 x 2 sections: the window fits the grid (preparePointBuffer's precondition)
            rc, pl = ol.oracle_window_place(cam, origins, grid, coarse, levels)
            assert rc == 0 and 0 <= pl.min_x <= pl.max_x < grid and 0 <= pl.min_y <= pl.max_y < grid
    assert shifts > 20
