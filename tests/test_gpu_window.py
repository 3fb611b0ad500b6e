"""Camera window over resident sections on the GPU (hmrt_compose_window, through the C ABI) against the oracle's restatement
of preparePointBuffer's copy loops (main.cpp:519-618): bit-exact, every level and the colour map; then the reference's
whole per-frame flow -- place the window, compose it, trace it -- against the CPU oracle on the oracle-composed window."""
import numpy as np
import pytest
import torch

import oraclelib as ol

pytestmark = pytest.mark.gpu


def _host_sections(coarse, levels, seed, aliased=False):
    rng = np.random.default_rng(seed)
    res, idx, total = ol.pyramid_layout(coarse, levels)
    pyr = [[rng.random(total, dtype=np.float32) * 50 for _ in range(2)] for _ in range(2)]
    col = [[rng.integers(0, 256, (res[0], res[0], 3), dtype=np.uint8) for _ in range(2)] for _ in range(2)]
    if aliased:  # the window lies inside one section: all four entries are the same buffers
        pyr = [[pyr[0][0]] * 2] * 2
        col = [[col[0][0]] * 2] * 2
    return res, idx, total, pyr, col


def _compose_gpu(ctx, pyr, col, coarse, levels, cx, cy, with_colors):
    res, idx, total = ol.pyramid_layout(coarse, levels)
    cache = {}

    def dev(a):
        if id(a) not in cache:
            cache[id(a)] = torch.from_numpy(a).cuda()
        return cache[id(a)]

    dp = [[dev(pyr[a][b]) for b in range(2)] for a in range(2)]
    dc = [[dev(col[a][b]) for b in range(2)] for a in range(2)] if with_colors else None
    out = torch.full((total,), float("nan"), dtype=torch.float32, device="cuda")
    out_c = torch.full((res[0], res[0], 3), 9, dtype=torch.uint8, device="cuda") if with_colors else None
    ctx.compose_window(dp, dc, coarse, levels, cx, cy, out, out_c)
    torch.cuda.synchronize()
    return out.cpu().numpy(), (out_c.cpu().numpy() if with_colors else None)


@pytest.fixture(params=[0, 1], ids=["tma", "per-thread"])
def window_variant(request, cuda_ctx):
    """Both formulations of hmrt_compose_window: TMA bulk copies (default) and the per-thread 128-bit gather."""
    cuda_ctx.set_window_variant(request.param)
    yield request.param
    cuda_ctx.set_window_variant(0)


@pytest.mark.parametrize("coarse,levels", [(8, 4), (4, 1), (3, 3), (5, 2), (16, 6), (2, 8), (64, 7)])
def test_compose_window_equals_oracle(cuda_ctx, window_variant, coarse, levels):
    res, idx, total, pyr, col = _host_sections(coarse, levels, seed=100 + coarse + levels)
    cells = {(0, 0), (0, coarse - 1), (coarse - 1, 0), (coarse // 2, coarse // 3), (coarse - 1, coarse - 1), (1 % coarse, 1 % coarse)}
    for cx, cy in sorted(cells):
        for with_colors in (True, False):
            got, got_c = _compose_gpu(cuda_ctx, pyr, col, coarse, levels, cx, cy, with_colors)
            want, want_c = ol.oracle_compose_window(pyr, col if with_colors else None, coarse, levels, cx, cy)
            assert (got.view(np.uint32) == want.view(np.uint32)).all(), (cx, cy)
            if with_colors:
                assert (got_c == want_c).all(), (cx, cy)


def test_compose_window_aliased_sections_and_errors(cuda_ctx):
    import hmrt

    coarse, levels = 8, 5
    res, idx, total, pyr, col = _host_sections(coarse, levels, seed=7, aliased=True)
    got, got_c = _compose_gpu(cuda_ctx, pyr, col, coarse, levels, 3, 5, True)
    want, want_c = ol.oracle_compose_window(pyr, col, coarse, levels, 3, 5)
    assert (got.view(np.uint32) == want.view(np.uint32)).all() and (got_c == want_c).all()
    d = torch.from_numpy(pyr[0][0]).cuda()
    secs = [[d, d], [d, d]]
    out = torch.empty_like(d)
    with pytest.raises(hmrt.HmrtError):  # cell_position outside the section
        cuda_ctx.compose_window(secs, None, coarse, levels, coarse, 0, out)
    with pytest.raises(hmrt.HmrtError):  # in place
        cuda_ctx.compose_window(secs, None, coarse, levels, 1, 1, d)
    # odd byte offsets take the scalar paths and still match
    big = torch.zeros(total + 1, dtype=torch.float32, device="cuda")
    cuda_ctx.compose_window(secs, None, coarse, levels, 2, 6, big[1:])
    torch.cuda.synchronize()
    want, _ = ol.oracle_compose_window(pyr, None, coarse, levels, 2, 6)
    assert (big[1:].cpu().numpy().view(np.uint32) == want.view(np.uint32)).all()


def test_reference_frame_flow_place_compose_trace(cuda_ctx):
    """main.cpp:947-966 for one frame: manage/prepare the window around the camera, then rayTrace in window coordinates.
    Sections are resident device pyramids; the composed window + camera_point_buffer traced on the GPU must equal the CPU
    oracle tracing the oracle-composed window."""
    import gpulib
    import hmrt

    coarse, levels, grid = 8, 5, 4  # sections of 128^2 cells
    size = coarse << (levels - 1)
    res, idx, total = ol.pyramid_layout(coarse, levels)
    cam0 = (1000.0, 0.0, 2000.0)
    org = np.array([[[cam0[0] + (i - grid / 2.0) * size, cam0[2] + (j - grid / 2.0) * size] for j in range(grid)] for i in range(grid)], np.float32)
    # every section: its own terrain (continuous across sections) as a max-pyramid with a colour map
    secs, cols = {}, {}
    for i in range(grid):
        for j in range(grid):
            xs = org[i, j, 0] + np.arange(size, dtype=np.float32)
            zs = org[i, j, 1] + np.arange(size, dtype=np.float32)
            fin = (30 + 18 * np.sin(xs[None, :] * 0.031) * np.cos(zs[:, None] * 0.027) + 6 * np.sin(xs[None, :] * 0.21 + zs[:, None] * 0.13)).astype(np.float32)
            secs[i, j] = ol.pyramid_from_finest(np.maximum(fin, 0), levels)
            c = np.zeros((size, size, 3), np.uint8)
            c[..., 0], c[..., 1], c[..., 2] = (i * 60 + 20), (j * 60 + 20), (fin * 3).astype(np.uint8)
            cols[i, j] = c
    dsecs = {k: torch.from_numpy(v).cuda() for k, v in secs.items()}
    dcols = {k: torch.from_numpy(v).cuda() for k, v in cols.items()}
    win = torch.empty(total, dtype=torch.float32, device="cuda")
    wcol = torch.empty((size, size, 3), dtype=torch.uint8, device="cuda")
    mh = float(max(v.max() for v in secs.values()))
    W, H = 160, 96
    for cam_world, fwd in [((1000.0, 70.0, 2000.0), (0.3, -0.5, 0.8)), ((1037.3, 55.0, 1890.6), (-0.6, -0.4, 0.2)), ((911.0, 90.0, 2100.5), (0.1, -0.9, -0.4))]:
        pl = hmrt.window_place(cam_world, org, grid, coarse, levels)
        pick = lambda d: [[d[pl.min_x, pl.min_y], d[pl.min_x, pl.max_y]], [d[pl.max_x, pl.min_y], d[pl.max_x, pl.max_y]]]  # noqa: E731
        cuda_ctx.compose_window(pick(dsecs), pick(dcols), coarse, levels, pl.cell_x, pl.cell_y, win, wcol)
        cuda_ctx.set_heightmap(win, wcol, coarse, levels, mh)
        cam = ol.make_camera(tuple(pl.camera), fwd)
        opts = ol.make_opts(mh, use_color_map=True)
        rgb, hits = gpulib.gpu_trace(cuda_ctx, W, H, [cam], opts)
        want_win, want_col = ol.oracle_compose_window(pick(secs), pick(cols), coarse, levels, pl.cell_x, pl.cell_y)
        assert (win.cpu().numpy().view(np.uint32) == want_win.view(np.uint32)).all() and (wcol.cpu().numpy() == want_col).all()
        want = ol.cpu_trace(ol.oracle().hmrt_oracle_trace, want_win, want_col, coarse, levels, W, H, cam, opts)
        ol.assert_same_trace((rgb[0], hits[0]), want, f"window frame at {cam_world}")
        assert (hits[0]["flags"] & 1).mean() > 0.5


def test_compose_window_default_size_bandwidth(cuda_ctx, window_variant):
    """The reference's default window (coarse 32, 8 levels: 4096^2 finest, 89.5 MB of floats + 50.3 MB of colours):
    bit-exact against the oracle and, as a sanity bound, faster than 1 ms (the reference uploads it over PCIe every frame)."""
    coarse, levels = 32, 8
    res, idx, total, pyr, col = _host_sections(coarse, levels, seed=11)
    dp = [[torch.from_numpy(pyr[a][b]).cuda() for b in range(2)] for a in range(2)]
    dc = [[torch.from_numpy(col[a][b]).cuda() for b in range(2)] for a in range(2)]
    out = torch.empty(total, dtype=torch.float32, device="cuda")
    out_c = torch.empty((res[0], res[0], 3), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        cuda_ctx.compose_window(dp, dc, coarse, levels, 13, 22, out, out_c)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        cuda_ctx.compose_window(dp, dc, coarse, levels, 13, 22, out, out_c)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    want, want_c = ol.oracle_compose_window(pyr, col, coarse, levels, 13, 22)
    assert (out.cpu().numpy().view(np.uint32) == want.view(np.uint32)).all() and (out_c.cpu().numpy() == want_c).all()
    assert ms < 1.0, f"compose_window took {ms:.3f} ms"


def test_section_grid_follows_camera_and_composes_window():
    """C++ host layer end to end (hmrt_host::SectionGrid): initializeSections at the first camera, manageSections along a
    walk that crosses section borders in both axes, GPU rasterisation of every (re)loaded section, then the window of the
    last camera -- against numpy sections + the oracle's preparePointBuffer restatement."""
    import ctypes as C
    import subprocess
    from pathlib import Path

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    host = Path(__file__).resolve().parent.parent / "gpu-heightmap-raytracer_b200" / "host"
    subprocess.run(["make", "-s", "-C", str(host), "libhmrt_host.so"], check=True)
    lib = C.CDLL(str(host / "libhmrt_host.so"))
    lib.hmrt_host_section_grid_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 5
    coarse, levels, grid = 8, 5, 4
    size = coarse << (levels - 1)
    res, idx, total = ol.pyramid_layout(coarse, levels)
    # steps shorter than a section (128 cells): manageSections shifts by one section per call, like the reference per frame
    cams = np.array([[4000.0, 60.0, 9000.0], [4100.0, 60.0, 9050.0], [4200.0, 60.0, 8950.0], [4300.0, 60.0, 8850.0], [4290.5, 61.0, 8750.25],
                     [4200.0, 61.0, 8650.0], [4100.0, 50.0, 8661.0], [4003.0, 50.0, 8661.0], [3921.75, 50.0, 8661.5]], np.float32)
    win = np.zeros(total, np.float32)
    pb = np.zeros(3, np.float32)
    origins = np.zeros((grid, grid, 2), np.float32)
    tags = np.zeros((grid, grid), np.int32)
    loaded = C.c_int(0)
    rc = lib.hmrt_host_section_grid_run(coarse, levels, grid, cams.ctypes.data, len(cams), win.ctypes.data, pb.ctypes.data,
                                        origins.ctypes.data, tags.ctypes.data, C.byref(loaded))
    assert rc == 0
    assert loaded.value > grid * grid  # the walk forced reloads
    rc, pl = ol.oracle_window_place(cams[-1], origins, grid, coarse, levels)
    assert rc == 0 and np.array_equal(np.array(pl.camera, np.float32), pb)

    def section(i, j):
        ox, oz = int(np.floor(origins[i, j, 0])), int(np.floor(origins[i, j, 1]))
        wx = (ox + np.arange(size, dtype=np.int64))[None, :].astype(np.uint64)
        wz = (oz + np.arange(size, dtype=np.int64))[:, None].astype(np.uint64)
        fin = (((wx * np.uint64(73856093)) ^ (wz * np.uint64(19349663))) & np.uint64(1023)).astype(np.float32) / np.float32(8.0)
        return ol.pyramid_from_finest(fin, levels)

    secs = [[section(pl.min_x, pl.min_y), section(pl.min_x, pl.max_y)], [section(pl.max_x, pl.min_y), section(pl.max_x, pl.max_y)]]
    want, _ = ol.oracle_compose_window(secs, None, coarse, levels, pl.cell_x, pl.cell_y)
    assert (win.view(np.uint32) == want.view(np.uint32)).all()
