"""Pins every host-side restatement to the REFERENCE'S OWN TEXT: oracle/_ref/libhmrt_refhost.so is main.cpp:44-618,
:745-781 and :995-1003 cut verbatim and compiled for the host (oracle/build_ref.sh, oracle/refhost_harness.cpp).

  hmrt_oracle_pyramid_layout / hmrt.pyramid_layout   == main.cpp:995-1003
  hmrt_oracle_rasterise_las                          == allocateSection's zero init + loadLASToSection (:174-238, :259-260)
  hmrt_host read_las_header                          == readLASHeader (:124-168)
  SectionLayout (C++ host) and the Python py_manage  == initializeSections / manageSections / rearrangeSections* (:276-448)
  hmrt_oracle_window_place / hmrt_window_place       == preparePointBuffer's host arithmetic (:461-516)
  hmrt_oracle_compose_window                         == preparePointBuffer's copy loops (:519-618)
  hmrt_host Camera::move / rotate                    == moveCamera / rotateCamera (:753-781), glm::rotate included

Only libLAS's decode (raw * scale + offset, class = low 5 bits) stays restated: its .cpp is not in the reference tree.
CPU only.  The GPU tests compare the CUDA path with these restatements (and, for a few cases, with this library directly)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oraclelib as ol
import rasterlib as rl
import refhostlib
from hmrt import las

REPO = Path(__file__).resolve().parent.parent
HOST = REPO / "gpu-heightmap-raytracer_b200" / "host"

pytestmark = pytest.mark.skipif(refhostlib.refhost() is None, reason="oracle/_ref/libhmrt_refhost.so not built (reference tree absent)")


@pytest.fixture(scope="module")
def hostlib():
    subprocess.run(["make", "-s", "-C", str(HOST), "libhmrt_host.so"], check=True)
    lib = C.CDLL(str(HOST / "libhmrt_host.so"))
    lib.hmrt_host_section_layout_step.argtypes = [C.c_int] * 4 + [C.c_void_p] * 5 + [C.c_int]
    lib.hmrt_host_las_scene.argtypes = [C.c_char_p, C.c_void_p]
    lib.hmrt_host_camera_step.argtypes = [C.c_void_p] + [C.c_float] * 6 + [C.c_void_p, C.c_float]
    lib.hmrt_host_camera_step.restype = None
    return lib


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


# ------------------------------------------------------------------------------------------------ tables (S2 / R9)
@pytest.mark.parametrize("variant,coarse", [("", 32), ("", 3), ("", 128), ("_g3l4", 5), ("_g3l4", 64)])
def test_pyramid_tables_equal_reference(variant, coarse):
    import hmrt

    rh = refhostlib.refhost(variant)
    res, idx, total = rh.config(coarse)
    assert (res, idx, total) == ol.pyramid_layout(coarse, rh.levels)
    ores, oidx, ototal = np.zeros(rh.levels, np.int32), np.zeros(rh.levels, np.int64), C.c_int64()
    assert ol.oracle().hmrt_oracle_pyramid_layout(coarse, rh.levels, ores.ctypes.data, oidx.ctypes.data, C.byref(ototal)) == 0
    assert (list(ores), list(oidx), ototal.value) == (res, idx, total)
    pres, pidx, ptotal = hmrt.pyramid_layout(coarse, rh.levels)     # host arithmetic of the C ABI (no device)
    assert (list(pres), list(pidx), ptotal) == (res, idx, total)
    rh.config(1)


# ------------------------------------------------------------------------------------------------ rasteriser (S1 / S3)
RASTER_CASES = [
    # variant, coarse, n, fmt, record_len, origin, cell, seed
    ("", 2, 20000, 2, None, (0.0, 0.0), (2.0, 2.0, 2.0), 1),
    ("", 2, 20000, 0, None, (0.0, 0.0), (2.0, 2.0, 2.0), 2),
    ("", 2, 20000, 1, 31, (0.0, 0.0), (2.0, 2.0, 2.0), 3),           # odd record length with user bytes
    ("", 2, 20000, 3, 37, (17.0, -9.5), (2.0, 2.0, 2.0), 4),          # section origin off the file minimum, fractional
    ("", 3, 30000, 2, None, (0.0, 0.0), (1.5, 1.5, 0.75), 5),         # non-power-of-two resolution and cell size
    ("", 1, 5000, 2, None, (100.0, 60.0), (2.0, 2.0, 2.0), 6),       # section covering a corner of the cloud
    ("_g3l4", 16, 20000, 2, None, (0.0, 0.0), (2.0, 2.0, 2.0), 7),
    ("_g3l4", 7, 20000, 3, None, (3.0, 3.0), (3.0, 3.0, 1.0), 8),
]


@pytest.mark.parametrize("variant,coarse,n,fmt,record_len,origin,cell,seed", RASTER_CASES)
def test_rasteriser_restatement_equals_reference_loader(variant, coarse, n, fmt, record_len, origin, cell, seed):
    rh = refhostlib.refhost(variant)
    res, idx, total = rh.config(coarse)
    hdr, rec = rl.synthetic_las(n, res[0], cell=cell[0], point_format=fmt, seed=seed, record_len=record_len)
    rh.set_las(hdr, rec, cell)
    want_pyr, want_col = rh.rasterise(origin)
    got_pyr, got_col = rl.oracle_rasterise(hdr, rec, coarse, rh.levels, cell=cell, origin=origin)
    assert want_pyr.max() > 0 and (want_pyr[idx[0]:] == 0).any()      # the case exercises both filled and empty cells
    assert (bits(got_pyr) == bits(want_pyr)).all()
    assert (got_col == want_col).all()
    # the property the CUDA path builds on (SURVEY 8(a) equivalence (i)): finest = max over points, level i+1 = 2x2 max
    assert (bits(ol.pyramid_from_finest(want_pyr[idx[0]:].reshape(res[0], res[0]), rh.levels)) == bits(want_pyr)).all()
    rh.config(1)


def test_rasteriser_heights_below_the_header_minimum():
    """Points below header.GetMinZ(): fZ < 0 never passes `buffer <= fZ` (main.cpp:229) on a zero-initialised section, so
    they leave no trace -- except fZ == -0.0f, which the reference stores over a +0.0f cell (and a later +0.0f point
    stores back): the sign of an all-zero cell follows file order there.  The restatement keeps that; the CUDA scatter's
    integer atomicMax cannot (INT_MIN loses), which is the one defined deviation: it leaves +0.0f.  Same value, same
    traversal (every use of a height is a comparison or `h - y`)."""
    rh = refhostlib.refhost("")
    coarse = 1
    res, idx, total = rh.config(coarse)
    r0 = res[0]
    scale, offset = (1.0, 1.0, 1e-3), (0.0, 0.0, 0.0)
    # header minimum z = 0: raw Z < 0 is below it; raw Z = 0 gives +0.0f; cell size 2
    X = np.array([1, 1, 5, 5, 9, 9, 13, 13, 13, 17], np.int32)
    Y = np.array([1, 1, 1, 1, 1, 1, 1, 1, 1, 1], np.int32)
    Z = np.array([-5000, 3000, 4000, -1, -7, -9, 0, 0, 0, -2000], np.int32)
    rec = las.encode_points(X, Y, Z, 2, np.zeros(10, np.uint8), np.full((10, 3), 65535, np.uint16))
    hdr = las.LasHeader(2, rec.shape[1], len(X), scale, offset, (0.0, 0.0, 0.0), (2.0 * r0, 2.0 * r0, 4.0))
    rh.set_las(hdr, rec)
    want_pyr, want_col = rh.rasterise((0.0, 0.0))
    got_pyr, got_col = rl.oracle_rasterise(hdr, rec, coarse, rh.levels)
    assert (bits(got_pyr) == bits(want_pyr)).all() and (got_col == want_col).all()
    fin = want_pyr[idx[0]:].reshape(r0, r0)
    assert fin[0, 0] == np.float32(1.5) and fin[0, 2] == np.float32(2.0)          # the negative point left no trace
    assert fin[0, 4] == 0 and fin[0, 8] == 0 and not np.signbit(fin).any()        # only-below-minimum cells stay +0
    assert (want_col[0, 4] == 255).all()                                          # ...but their colour IS written (main.cpp:223-224 precede the test)

    # -0.0f: a denormal-range negative height, (float)(-1e-50) == -0.0f
    hdr2 = las.LasHeader(2, rec.shape[1], 3, (1.0, 1.0, 1e-50), offset, (0.0, 0.0, 0.0), (2.0 * r0, 2.0 * r0, 4.0))
    rec2 = las.encode_points(np.array([1, 5, 5], np.int32), np.array([1, 1, 1], np.int32), np.array([-1, -1, 0], np.int32), 2,
                             np.zeros(3, np.uint8), np.zeros((3, 3), np.uint16))
    rh.set_las(hdr2, rec2)
    want2, _ = rh.rasterise((0.0, 0.0))
    got2, _ = rl.oracle_rasterise(hdr2, rec2, coarse, rh.levels, with_colors=False)
    assert (bits(got2) == bits(want2)).all()
    fin2 = want2[idx[0]:].reshape(r0, r0)
    assert np.signbit(fin2[0, 0]) and fin2[0, 0] == 0          # -0.0f stored by the reference
    assert not np.signbit(fin2[0, 2])                          # a later +0.0f point stores back
    rh.config(1)


def test_read_las_header_equals_reference(hostlib, tmp_path):
    rh = refhostlib.refhost("")
    rh.config(2)
    for seed, fmt in [(3, 2), (11, 0), (12, 3)]:
        hdr, rec = rl.synthetic_las(1000, 256, seed=seed, point_format=fmt)
        rh.set_las(hdr, rec)
        cam, bounds, max_h, cell = rh.read_header()
        path = tmp_path / f"t{seed}.las"
        las.write_las(path, hdr, rec)
        scene = np.zeros(9, np.float32)
        assert hostlib.hmrt_host_las_scene(str(path).encode(), scene.ctypes.data) == 0
        assert (bits(scene[0:3]) == bits(cell)).all()
        assert (bits(scene[3:5]) == bits(bounds)).all()
        assert (bits(scene[5:8]) == bits(cam)).all()
        assert bits(scene[8:9])[0] == bits(np.array([max_h]))[0]
    rh.config(1)


# ------------------------------------------------------------------------------------------------ section manager (f3)
def _host_step(lib, init, grid, coarse, levels, cam, origins, tags):
    cam = np.asarray(cam, np.float32)
    n_max = grid * grid if init else 4 * grid
    loads = np.zeros((n_max, 3), np.int32)
    lorg = np.zeros((n_max, 2), np.float32)
    n = lib.hmrt_host_section_layout_step(int(init), grid, coarse, levels, cam.ctypes.data, origins.ctypes.data, tags.ctypes.data,
                                          loads.ctypes.data, lorg.ctypes.data, len(loads))
    assert n >= 0
    return loads[:n], lorg[:n]


@pytest.mark.parametrize("variant,coarse", [("", 32), ("", 2), ("_g3l4", 8)])
def test_section_walk_equals_reference(hostlib, variant, coarse):
    from test_section_layout import py_manage

    rh = refhostlib.refhost(variant)
    rh.config(coarse)
    grid, levels = rh.grid, rh.levels
    size = float(coarse << (levels - 1))
    rng = np.random.default_rng(coarse)
    cam = np.array([1234.5, 80.0, -321.25], np.float32)
    want_org, want_slots, want_lorg = rh.init_sections(cam)
    origins = np.zeros((grid, grid, 2), np.float32)
    tags = np.zeros((grid, grid), np.int32)
    loads, lorg = _host_step(hostlib, True, grid, coarse, levels, cam, origins, tags)
    assert (bits(origins) == bits(want_org)).all()
    assert [tuple(l[:2]) for l in loads] == [tuple(s) for s in want_slots] and (bits(lorg) == bits(want_lorg)).all()
    py_org, py_tags = origins.copy(), tags.copy()
    shifts = 0
    for step in range(400):
        # mostly small moves, sometimes a jump of more than a section (two shifts in one call cannot happen per axis,
        # but x and y shifts can coincide)
        d = rng.uniform(-0.45, 0.45, 2) * size if step % 7 else rng.uniform(-0.99, 0.99, 2) * size
        cam = np.array([cam[0] + d[0], cam[1], cam[2] + d[1]], np.float32)
        want_org, want_slots, want_lorg = rh.manage(cam)
        loads, lorg = _host_step(hostlib, False, grid, coarse, levels, cam, origins, tags)
        py_loads = py_manage(py_org, py_tags, cam, size)
        assert (bits(origins) == bits(want_org)).all(), step
        assert (bits(py_org) == bits(want_org)).all(), step
        # When an x shift and a y shift coincide the reference allocates a column, shifts it along y (dropping its last
        # section again) and allocates a row: the loaders that survive the call are what must be loaded.  Reference side:
        # allocations whose buffer still sits in a slot, (final slot, loader origin).  Host side: the last load per
        # recycled section object (tag), looked up at the slot that holds the tag after the call.
        want_final = {(int(s[0]), int(s[1])): tuple(bits(o)) for s, o in zip(want_slots, want_lorg)}
        last = {int(l[2]): tuple(bits(o)) for l, o in zip(loads, lorg)}
        got_final = {tuple(int(v) for v in np.argwhere(tags == t)[0]): o for t, o in last.items()}
        assert got_final == want_final, step
        for (i, j), o in want_final.items():
            assert tuple(bits(want_org[i, j])) == o
        assert len({(i, j) for i, j, _ in py_loads}) >= len(want_final) > 0 or not py_loads, step
        shifts += len(want_slots) > 0
    assert shifts > 20
    rh.config(1)


# ------------------------------------------------------------------------------------------------ window (f1)
def _tagged_sections(rh, rng):
    """Every section gets content that identifies (section, level, cell): a wrong source cell cannot go unnoticed."""
    g = rh.grid
    secs = {}
    for i in range(g):
        for j in range(g):
            pyr = (rng.random(rh.total, dtype=np.float32) + np.float32(10 * (i * g + j))).astype(np.float32)
            col = rng.integers(0, 256, (rh.res[0], rh.res[0], 3), dtype=np.uint8)
            col[..., 0] = i * g + j
            rh.fill_section(i, j, pyr, col)
            secs[i, j] = (pyr, col)
    return secs


@pytest.mark.parametrize("variant,coarse", [("", 4), ("", 3), ("_g3l4", 8), ("_g3l4", 5)])
def test_window_restatement_equals_reference_prepare(variant, coarse):
    import hmrt

    rh = refhostlib.refhost(variant)
    rh.config(coarse)
    grid, levels = rh.grid, rh.levels
    size = float(coarse << (levels - 1))
    rng = np.random.default_rng(100 + coarse)
    cam0 = np.array([517.25, 33.0, -90.5], np.float32)
    origins, _, _ = rh.init_sections(cam0)
    secs = _tagged_sections(rh, rng)
    lo = origins[grid // 2 - 1, grid // 2 - 1] if grid % 2 == 0 else origins[0, 0] + np.float32(size / 2)
    cams = [cam0]
    # anywhere manageSections leaves the camera (window inside the grid), incl. exactly on section and coarse-cell borders
    span = size * (grid - 2) if grid > 2 else size
    base = origins[1, 1] if grid % 2 == 0 else origins[0, 0] + np.float32(size / 2)
    for _ in range(40):
        cams.append(np.array([base[0] + rng.uniform(0, span * 0.999), 20.0, base[1] + rng.uniform(0, span * 0.999)], np.float32))
    top = float(1 << (levels - 1))
    cams.append(np.array([origins[grid // 2, 0, 0], 5.0, origins[0, grid // 2, 1]], np.float32))
    cams.append(np.array([origins[grid // 2, 0, 0] + 3 * top, 5.0, origins[0, grid // 2, 1] - 2 * top], np.float32))
    checked_full = 0
    for k, cam in enumerate(cams):
        rc, pl = ol.oracle_window_place(cam, origins, grid, coarse, levels)
        if rc != 0:
            continue
        full = k < 12
        want_cpb, want_pyr, want_col = rh.prepare(cam, want_buffers=full)
        assert (bits(pl.camera) == bits(want_cpb)).all(), cam
        got = hmrt.window_place(cam, origins, grid, coarse, levels)
        assert (bits(got.camera) == bits(want_cpb)).all()
        assert (got.min_x, got.min_y, got.max_x, got.max_y, got.cell_x, got.cell_y) == (pl.min_x, pl.min_y, pl.max_x, pl.max_y, pl.cell_x, pl.cell_y)
        if not full:
            continue
        xs, ys = (pl.min_x, pl.max_x), (pl.min_y, pl.max_y)
        pyr = [[secs[xs[a], ys[b]][0] for b in range(2)] for a in range(2)]
        col = [[secs[xs[a], ys[b]][1] for b in range(2)] for a in range(2)]
        got_pyr, got_col = ol.oracle_compose_window(pyr, col, coarse, levels, pl.cell_x, pl.cell_y)
        assert (bits(got_pyr) == bits(want_pyr)).all(), cam
        assert (got_col == want_col).all(), cam
        checked_full += 1
    assert checked_full >= 8
    rh.config(1)


# ------------------------------------------------------------------------------------------------ camera (f4)
def test_camera_move_and_rotate_equal_reference(hostlib):
    """Camera::move == moveCamera (main.cpp:753-772), Camera::rotate == rotateCamera (:776-781, two glm::rotate calls from
    the reference's own GLM 0.9.8.3), bit for bit, along a random fly-through that hits every clamp."""
    rh = refhostlib.refhost("")
    rng = np.random.default_rng(42)
    bounds = np.array([5000.0, 3000.0], np.float32)
    max_h = 120.0
    pos = np.array([2500.0, 100.0, 1500.0], np.float32)
    fwd = np.array([0.0, -0.6689647, 0.7432941], np.float32)
    clamped = 0
    for step in range(600):
        dt = float(rng.choice([0.004, 0.016, 0.1, 0.7, 1.5]))
        mv_fwd, mv_rht, mv_up = [float(rng.choice([0.0, 250.0, -250.0, 2500.0])) for _ in range(3)]
        yaw, pitch = float(rng.choice([0.0, 1.0, -1.0])), float(rng.choice([0.0, 1.0, -1.0]))
        if abs(fwd[1]) > 0.95:
            pitch = -np.sign(fwd[1]) * 1.0 if pitch else 0.0    # stay away from the pole (cross(forward, up) -> 0: NaN in both)
            dt = min(dt, 0.1)
        want_pos = rh.move_camera(pos, fwd, (mv_rht, mv_up, mv_fwd), dt, bounds, max_h)
        want_fwd = rh.rotate_camera(fwd, yaw, pitch, dt)
        st = np.concatenate([pos, fwd]).astype(np.float32)
        hostlib.hmrt_host_camera_step(st.ctypes.data, mv_fwd, mv_rht, mv_up, yaw, pitch, dt, bounds.ctypes.data, max_h)
        assert (bits(st[0:3]) == bits(want_pos)).all(), (step, st[0:3], want_pos)
        assert (bits(st[3:6]) == bits(want_fwd)).all(), (step, st[3:6], want_fwd)
        clamped += int(want_pos[0] in (0.0,) or want_pos[1] in (0.0, np.float32(4 * max_h)) or want_pos[2] == 0.0)
        pos, fwd = want_pos, want_fwd
        assert np.isfinite(fwd).all()
    assert clamped > 10
