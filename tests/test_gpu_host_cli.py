"""End to end through the C++ host layer: hmrt_render (LAS / PointdataGenerator input -> GPU rasterisation ->
max-mipmap -> ray traversal -> PPM) against the same pipeline run entirely on the CPU oracle."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oraclelib as ol
import rasterlib as rl
from hmrt import LasTransform, las

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent
HOST = REPO / "gpu-heightmap-raytracer_b200" / "host"


def _read_ppm(path):
    data = Path(path).read_bytes()
    magic, dims, maxv, rest = data.split(b"\n", 3)
    w, h = map(int, dims.split())
    assert magic == b"P6" and maxv == b"255"
    return np.frombuffer(rest, np.uint8).reshape(h, w, 3)[::-1]  # the tool writes the GL bottom row last


def _cli(*args):
    subprocess.run(["make", "-s", "-C", str(HOST)], check=True)
    return subprocess.run([str(HOST / "hmrt_render"), *map(str, args)], check=True, capture_output=True, text=True).stdout


def _default_camera(r0, max_height):
    cam = ol.Camera()
    cam.frame_dim[:] = [32.0, 18.0, 20.0]
    cam.forward[:] = [0.0, float(np.float32(-0.6689647)), float(np.float32(0.7432941))]
    cam.position[:] = [float(np.float32(r0 * 0.5)), float(np.float32(1.5) * np.float32(max_height)), float(np.float32(r0 * 0.5))]
    return cam


def test_generate_rasterise_render(tmp_path):
    """BASELINE config 1 shape, scaled: seeded PointdataGenerator terrain -> 256^2 grid, 8 levels, one 160x120 frame."""
    n, W, H = 256, 160, 120
    out = _cli("--generate", n, "--seed", 5, "--width", W, "--height", H, "--shadows", "--out", tmp_path / "g")
    assert "rasterised" in out and "FPS" in out
    got = _read_ppm(tmp_path / "g_0000.ppm")
    # the same pipeline on the oracle
    xyz = np.zeros(((n + 1) ** 2, 3), np.float32)
    assert ol.oracle().hmrt_oracle_pdg_generate(n, 5, xyz.ctypes.data) == 0
    xf = LasTransform()
    xf.scale[:] = (1.0, 1.0, 1.0)
    xf.cell_size[:] = (1.0, 1.0, 1.0)
    res, idx, total = ol.pyramid_layout(n >> 7, 8)
    pyr = np.zeros(total, np.float32)
    assert ol.oracle().hmrt_oracle_rasterise_xyz(xyz.ctypes.data, len(xyz), C.byref(xf), pyr.ctypes.data, n >> 7, 8) == 0
    mh = float(pyr[: (n >> 7) ** 2].max())
    opts = ol.make_opts(mh, shadows=True)
    opts.light_dir[:] = [float(np.float32(v)) for v in (0.3244, 0.8111, 0.4867)]
    want, _ = ol.cpu_trace(ol.oracle().hmrt_oracle_trace, pyr, None, n >> 7, 8, W, H, _default_camera(n, mh), opts)
    assert (got == want).all()


def test_las_file_to_frames(tmp_path):
    hdr, rec = rl.synthetic_las(300_000, 512, seed=8, frac_outside=0.0)
    path = tmp_path / "cloud.las"
    las.write_las(path, hdr, rec)
    W, H = 128, 96
    out = _cli("--las", path, "--width", W, "--height", H, "--frames", 2, "--colors", "--out", tmp_path / "l")
    assert "LAS File Loaded" in out
    got = _read_ppm(tmp_path / "l_0000.ppm")
    assert (tmp_path / "l_0001.ppm").exists()
    # oracle: the tool sizes the grid to the next power of two >= the header extent (1024 cells of 2.0 here -> 512 cells)
    r0 = 512
    pyr, cmap = rl.oracle_rasterise(hdr, rec, r0 >> 7, 8)
    # readLASHeader (main.cpp:153-164): camera on the first point at height (maxZ - minZ) / cell, that value is max_height
    X0, Y0 = int(rec[0, 0:4].view("<i4")[0]), int(rec[0, 4:8].view("<i4")[0])
    mh = np.float32(np.float32(hdr.max[2] - hdr.min[2]) / np.float32(2.0))
    cam = ol.Camera()
    cam.frame_dim[:] = [32.0, 18.0, 20.0]
    cam.forward[:] = [0.0, float(np.float32(-0.6689647)), float(np.float32(0.7432941))]
    cam.position[:] = [float(np.float32(((X0 * hdr.scale[0] + hdr.offset[0]) - hdr.min[0]) / 2.0)),
                       float(np.float32((hdr.max[2] - hdr.min[2]) / 2.0)),
                       float(np.float32(((Y0 * hdr.scale[1] + hdr.offset[1]) - hdr.min[1]) / 2.0))]
    opts = ol.make_opts(float(mh), use_color_map=True)
    want, _ = ol.cpu_trace(ol.oracle().hmrt_oracle_trace, pyr, cmap, r0 >> 7, 8, W, H, cam, opts)
    assert (got == want).all()


def test_las_section_grid_window_frames(tmp_path):
    """The reference's own main loop (main.cpp:947-966) headless: 4 x 4 resident sections around the camera, manageSections,
    one composed window per frame, rayTrace in window coordinates -- frame 0 against the same flow on the CPU oracle."""
    import hmrt

    hdr, rec = rl.synthetic_las(400_000, 512, seed=12, frac_outside=0.0)
    path = tmp_path / "cloud.las"
    las.write_las(path, hdr, rec)
    W, H, coarse, levels, grid = 128, 96, 1, 8, 4  # sections of 128^2 cells
    out = _cli("--las", path, "--sections", grid, "--coarse", coarse, "--levels", levels, "--width", W, "--height", H, "--frames", 3,
               "--colors", "--out", tmp_path / "s")
    assert "sections of 128^2 cells loaded" in out and (tmp_path / "s_0002.ppm").exists()
    got = _read_ppm(tmp_path / "s_0000.ppm")
    size = coarse << (levels - 1)
    X0, Y0 = int(rec[0, 0:4].view("<i4")[0]), int(rec[0, 4:8].view("<i4")[0])
    mh = np.float32(np.float32(hdr.max[2] - hdr.min[2]) / np.float32(2.0))
    cam_world = (np.float32(((X0 * hdr.scale[0] + hdr.offset[0]) - hdr.min[0]) / 2.0), np.float32((hdr.max[2] - hdr.min[2]) / 2.0),
                 np.float32(((Y0 * hdr.scale[1] + hdr.offset[1]) - hdr.min[1]) / 2.0))
    org = np.array([[[np.float32(cam_world[0] + np.float32((i - grid / 2.0) * size)), np.float32(cam_world[2] + np.float32((j - grid / 2.0) * size))]
                     for j in range(grid)] for i in range(grid)], np.float32)  # initializeSections, main.cpp:276-288
    rc, pl = ol.oracle_window_place(cam_world, org, grid, coarse, levels)
    assert rc == 0
    sec = {(i, j): rl.oracle_rasterise(hdr, rec, coarse, levels, origin=tuple(org[i, j])) for i in {pl.min_x, pl.max_x} for j in {pl.min_y, pl.max_y}}
    pick = lambda k: [[sec[pl.min_x, pl.min_y][k], sec[pl.min_x, pl.max_y][k]], [sec[pl.max_x, pl.min_y][k], sec[pl.max_x, pl.max_y][k]]]  # noqa: E731
    win, wcol = ol.oracle_compose_window(pick(0), pick(1), coarse, levels, pl.cell_x, pl.cell_y)
    cam = ol.Camera()
    cam.frame_dim[:] = [32.0, 18.0, 20.0]
    cam.forward[:] = [0.0, float(np.float32(-0.6689647)), float(np.float32(0.7432941))]
    cam.position[:] = list(pl.camera)
    want, _ = ol.cpu_trace(ol.oracle().hmrt_oracle_trace, win, wcol, coarse, levels, W, H, cam, ol.make_opts(float(mh), use_color_map=True))
    assert (got == want).all()
