"""Parity of the CUDA rasterisation (scatter + max-mipmap build + colour resolve, through the C ABI)
against the restatement of loadLASToSection (main.cpp:193-234): heightmap and colour map bit-exact."""
import ctypes as C

import numpy as np
import pytest
import torch

import oraclelib as ol
import rasterlib as rl

pytestmark = pytest.mark.gpu


def _gpu_rasterise(ctx, hdr, rec, coarse, levels, cell=(2.0, 2.0, 2.0), origin=(0.0, 0.0), chunks=1, with_colors=True):
    res, idx, total = ol.pyramid_layout(coarse, levels)
    pyr = torch.empty(total, dtype=torch.float32, device="cuda").fill_(-1.0)  # clear_section must zero it
    keys = torch.empty(res[0] * res[0], dtype=torch.int64, device="cuda").fill_(-1) if with_colors else None
    cmap = torch.empty((res[0], res[0], 3), dtype=torch.uint8, device="cuda").fill_(7) if with_colors else None
    ctx.clear_section(pyr, coarse, levels, keys, cmap)
    xf = hdr.transform(cell, origin)
    n = len(rec)
    bounds = np.linspace(0, n, chunks + 1).astype(np.int64)
    for a, b in zip(bounds[:-1], bounds[1:]):
        d = torch.from_numpy(np.ascontiguousarray(rec[a:b])).cuda()
        ctx.scatter_las(d, int(b - a), rec.shape[1], hdr.point_format, xf, pyr, coarse, levels, first_index=int(a), color_keys=keys)
    ctx.build_mips(pyr, coarse, levels)
    if with_colors:
        ctx.resolve_colors(keys, cmap, res[0] * res[0])
    torch.cuda.synchronize()
    return pyr.cpu().numpy(), (cmap.cpu().numpy() if with_colors else None)


@pytest.mark.parametrize("fmt,record_len", [(0, None), (1, None), (2, None), (3, None), (2, 31), (3, 40)])
def test_las_formats_bit_exact(cuda_ctx, fmt, record_len):
    hdr, rec = rl.synthetic_las(150_001, 256, point_format=fmt, seed=fmt, record_len=record_len)
    want_p, want_c = rl.oracle_rasterise(hdr, rec, 256 >> 5, 6)
    got_p, got_c = _gpu_rasterise(cuda_ctx, hdr, rec, 256 >> 5, 6)
    assert (got_p.view(np.uint32) == want_p.view(np.uint32)).all()
    assert (got_c == want_c).all()


@pytest.mark.parametrize("r0,levels", [(128, 8), (384, 8), (256, 3), (96, 4), (64, 1), (512, 9)])
def test_grid_shapes_fused_and_generic_mips(cuda_ctx, r0, levels):
    """128-multiples take the fused single-pass mip kernel, others the per-level kernel; levels > 8 use both."""
    coarse = r0 >> (levels - 1)
    assert coarse << (levels - 1) == r0
    hdr, rec = rl.synthetic_las(60_000, r0, seed=r0)
    want_p, want_c = rl.oracle_rasterise(hdr, rec, coarse, levels)
    got_p, got_c = _gpu_rasterise(cuda_ctx, hdr, rec, coarse, levels)
    assert (got_p.view(np.uint32) == want_p.view(np.uint32)).all()
    assert (got_c == want_c).all()


def test_chunked_and_shuffled_input(cuda_ctx):
    """Heights do not depend on point order or chunking; colours follow FILE order (first_index)."""
    hdr, rec = rl.synthetic_las(120_000, 256, seed=11)
    want_p, want_c = rl.oracle_rasterise(hdr, rec, 8, 6)
    got_p, got_c = _gpu_rasterise(cuda_ctx, hdr, rec, 8, 6, chunks=5)
    assert (got_p == want_p).all() and (got_c == want_c).all()
    perm = np.random.default_rng(3).permutation(len(rec))
    got_p2, _ = _gpu_rasterise(cuda_ctx, hdr, rec[perm], 8, 6, with_colors=False)
    assert (got_p2 == want_p).all()


def test_cell_size_and_section_origin(cuda_ctx):
    """Non-power-of-two cell size (true fp32 divide) and a section origin (main.cpp:174,205-206)."""
    hdr, rec = rl.synthetic_las(80_000, 256, cell=1.5, seed=4)
    for origin in [(0.0, 0.0), (64.0, 32.0), (-10.5, 3.25)]:
        want_p, want_c = rl.oracle_rasterise(hdr, rec, 4, 6, cell=(1.5, 1.5, 0.75), origin=origin)
        got_p, got_c = _gpu_rasterise(cuda_ctx, hdr, rec, 4, 6, cell=(1.5, 1.5, 0.75), origin=origin)
        assert (got_p.view(np.uint32) == want_p.view(np.uint32)).all()
        assert (got_c == want_c).all()


def test_empty_and_all_rejected_inputs(cuda_ctx):
    hdr, rec = rl.synthetic_las(1000, 64, seed=1)
    res, idx, total = ol.pyramid_layout(64 >> 3, 4)
    got_p, got_c = _gpu_rasterise(cuda_ctx, hdr, rec[:0], 8, 4)
    assert (got_p.view(np.uint32) == 0).all() and (got_c == 0).all()
    rec7 = rec.copy()
    rec7[:, 15] = 7
    got_p, got_c = _gpu_rasterise(cuda_ctx, hdr, rec7, 8, 4)
    assert (got_p.view(np.uint32) == 0).all() and (got_c == 0).all()


def test_pdg_xyz_points(cuda_ctx):
    """BASELINE config 1 input: seeded PointdataGenerator terrain, 1025^2 points at integer (i, j),
    cell size 1, origin 0 -> 1024^2 grid, 8 levels."""
    n = 1024
    xyz = np.zeros(((n + 1) ** 2, 3), np.float32)
    assert ol.oracle().hmrt_oracle_pdg_generate(n, 2024, xyz.ctypes.data) == 0
    from hmrt import LasTransform

    xf = LasTransform()
    xf.scale[:] = (1.0, 1.0, 1.0)
    xf.cell_size[:] = (1.0, 1.0, 1.0)
    res, idx, total = ol.pyramid_layout(8, 8)
    want = np.zeros(total, np.float32)
    assert ol.oracle().hmrt_oracle_rasterise_xyz(xyz.ctypes.data, len(xyz), C.byref(xf), want.ctypes.data, 8, 8) == 0
    pyr = torch.empty(total, dtype=torch.float32, device="cuda")
    cuda_ctx.clear_section(pyr, 8, 8)
    cuda_ctx.scatter_xyz(torch.from_numpy(xyz).cuda(), len(xyz), xf, pyr, 8, 8)
    cuda_ctx.build_mips(pyr, 8, 8)
    torch.cuda.synchronize()
    assert (pyr.cpu().numpy().view(np.uint32) == want.view(np.uint32)).all()
    assert want[idx[0]:].max() > 1.0


def test_full_size_16384_properties(cuda_ctx):
    """BASELINE config 4 shape on one GPU (scaled to 40 M points): 16384^2 grid.  Properties: sharded scatter
    + max == single scatter (what the NCCL max all-reduce computes), mip levels are 2x2 maxima, oracle on a subset."""
    r0, levels, coarse = 16384, 8, 128
    res, idx, total = ol.pyramid_layout(coarse, levels)
    n = 40_000_000
    g = torch.Generator(device="cuda").manual_seed(7)
    ext_raw = int(r0 * 2.0 / 0.01)
    X = torch.randint(0, ext_raw, (n,), device="cuda", generator=g, dtype=torch.int32)
    Y = torch.randint(0, ext_raw, (n,), device="cuda", generator=g, dtype=torch.int32)
    Z = (2000 + 1500 * torch.sin(X.float() * 1e-5) + torch.randint(0, 300, (n,), device="cuda", generator=g)).to(torch.int32)
    rec = torch.zeros((n, 20), dtype=torch.uint8, device="cuda")
    rec[:, 0:4] = X.view(torch.uint8).view(n, 4)
    rec[:, 4:8] = Y.view(torch.uint8).view(n, 4)
    rec[:, 8:12] = Z.view(torch.uint8).view(n, 4)
    from hmrt import las

    hdr = las.LasHeader(0, 20, n, (0.01, 0.01, 0.01), (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), (r0 * 2.0, r0 * 2.0, 40.0))
    xf = hdr.transform()
    single = torch.empty(total, dtype=torch.float32, device="cuda")
    cuda_ctx.clear_section(single, coarse, levels)
    cuda_ctx.scatter_las(rec, n, 20, 0, xf, single, coarse, levels)
    # two shards into private grids, combined with an integer max on the bit patterns
    merged = None
    for lo, hi in [(0, n // 2), (n // 2, n)]:
        part = torch.empty(total, dtype=torch.float32, device="cuda")
        cuda_ctx.clear_section(part, coarse, levels)
        cuda_ctx.scatter_las(rec[lo:hi], hi - lo, 20, 0, xf, part, coarse, levels, first_index=lo)
        fin = part[idx[0]:].view(torch.int32)
        merged = fin.clone() if merged is None else torch.maximum(merged, fin)
        del part
    torch.cuda.synchronize()
    assert torch.equal(merged, single[idx[0]:].view(torch.int32))
    del merged
    cuda_ctx.build_mips(single, coarse, levels)
    torch.cuda.synchronize()
    for l in range(1, levels):
        fine = single[idx[l - 1]: idx[l - 1] + res[l - 1] ** 2].view(res[l], 2, res[l], 2)
        coarse_l = single[idx[l]: idx[l] + res[l] ** 2].view(res[l], res[l])
        assert torch.equal(fine.amax(dim=(1, 3)), coarse_l), f"level {l} is not the 2x2 max of level {l - 1}"
    assert float(single[0:coarse * coarse].max()) == float(single[idx[0]:].max())
    # oracle on a subset of the same records (bit-exact finest level)
    sub = rec[:300_000].cpu().numpy()
    want = np.zeros(total, np.float32)
    assert ol.oracle().hmrt_oracle_rasterise_las(sub.ctypes.data, len(sub), 20, 0, C.byref(xf), want.ctypes.data, coarse, levels, None) == 0
    got = torch.empty(total, dtype=torch.float32, device="cuda")
    cuda_ctx.clear_section(got, coarse, levels)
    cuda_ctx.scatter_las(rec[:300_000], 300_000, 20, 0, xf, got, coarse, levels)
    cuda_ctx.build_mips(got, coarse, levels)
    torch.cuda.synchronize()
    assert torch.equal(got.cpu().view(torch.int32), torch.from_numpy(want).view(torch.int32))


def _device_records(n, r0, seed, skew=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    ext_raw = int(r0 * 2.0 / 0.01)
    hi = ext_raw // 16 if skew else ext_raw  # skew: every point in the first tile column/row -> bucket overflow
    X = torch.randint(0, hi, (n,), device="cuda", generator=g, dtype=torch.int32)
    Y = torch.randint(0, hi, (n,), device="cuda", generator=g, dtype=torch.int32)
    Z = torch.randint(0, 90000, (n,), device="cuda", generator=g, dtype=torch.int32)
    rec = torch.zeros((n, 20), dtype=torch.uint8, device="cuda")
    rec[:, 0:4] = X.view(torch.uint8).view(n, 4)
    rec[:, 4:8] = Y.view(torch.uint8).view(n, 4)
    rec[:, 8:12] = Z.view(torch.uint8).view(n, 4)
    rec[:, 15] = torch.randint(0, 12, (n,), device="cuda", generator=g, dtype=torch.int32).to(torch.uint8)  # some class 7
    return rec


@pytest.mark.parametrize("skew", [False, True])
def test_binned_scatter_equals_direct(cuda_ctx, skew):
    """hmrt_set_scatter_mode: the tile-binned path (unordered clouds) and the direct path must give the same bits,
    also when every bucket overflows (skewed input) and in auto mode; oracle on a subset."""
    from hmrt import las

    r0, levels, coarse = 8192, 8, 64
    res, idx, total = ol.pyramid_layout(coarse, levels)
    n = 6_000_000
    rec = _device_records(n, r0, 21, skew)
    hdr = las.LasHeader(0, 20, n, (0.01, 0.01, 0.01), (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), (r0 * 2.0, r0 * 2.0, 900.0))
    xf = hdr.transform()
    out = {}
    try:
        for mode in (1, 2, 0):
            cuda_ctx.set_scatter_mode(mode)
            pyr = torch.empty(total, dtype=torch.float32, device="cuda")
            cuda_ctx.clear_section(pyr, coarse, levels)
            cuda_ctx.scatter_las(rec, n, 20, 0, xf, pyr, coarse, levels)
            torch.cuda.synchronize()
            out[mode] = pyr[idx[0]:].view(torch.int32).clone()
            del pyr
    finally:
        cuda_ctx.set_scatter_mode(0)
    assert torch.equal(out[1], out[2]) and torch.equal(out[1], out[0])
    assert int((out[1] != 0).sum()) > 1000
    # oracle on the same first 200k records
    sub = rec[:200_000].cpu().numpy()
    want = np.zeros(total, np.float32)
    assert ol.oracle().hmrt_oracle_rasterise_las(sub.ctypes.data, len(sub), 20, 0, C.byref(xf), want.ctypes.data, coarse, levels, None) == 0
    # binned mode needs >= 4 M points to engage, so compare the direct result of the subset
    pyr = torch.empty(total, dtype=torch.float32, device="cuda")
    cuda_ctx.clear_section(pyr, coarse, levels)
    cuda_ctx.scatter_las(rec[:200_000], 200_000, 20, 0, xf, pyr, coarse, levels)
    torch.cuda.synchronize()
    assert torch.equal(pyr[idx[0]:].cpu().view(torch.int32), torch.from_numpy(want[idx[0]:]).view(torch.int32))


def test_baseline_config_1_end_to_end(cuda_ctx):
    """BASELINE configs[0] exactly: seeded PointdataGenerator terrain -> 1024^2 heightmap (8 levels) -> one 640x480
    primary-ray frame, camera at the grid centre at 1.5 x max height looking along normalize(0, -0.9, 1) (main.cpp:56).
    GPU pipeline (scatter_xyz + build_mips + trace) against the same pipeline on the CPU oracle, and -- when the
    reference's own code is available -- against oracle/_ref."""
    import gpulib
    from hmrt import LasTransform

    n, W, H = 1024, 640, 480
    xyz = np.zeros(((n + 1) ** 2, 3), np.float32)
    assert ol.oracle().hmrt_oracle_pdg_generate(n, 1, xyz.ctypes.data) == 0
    xf = LasTransform()
    xf.scale[:] = (1.0, 1.0, 1.0)
    xf.cell_size[:] = (1.0, 1.0, 1.0)
    res, idx, total = ol.pyramid_layout(8, 8)
    want_pyr = np.zeros(total, np.float32)
    assert ol.oracle().hmrt_oracle_rasterise_xyz(xyz.ctypes.data, len(xyz), C.byref(xf), want_pyr.ctypes.data, 8, 8) == 0
    pyr = torch.empty(total, dtype=torch.float32, device="cuda")
    cuda_ctx.clear_section(pyr, 8, 8)
    cuda_ctx.scatter_xyz(torch.from_numpy(xyz).cuda(), len(xyz), xf, pyr, 8, 8)
    cuda_ctx.build_mips(pyr, 8, 8)
    torch.cuda.synchronize()
    assert (pyr.cpu().numpy().view(np.uint32) == want_pyr.view(np.uint32)).all()
    mh = float(want_pyr[:64].max())
    cuda_ctx.set_heightmap(pyr, None, 8, 8, mh)
    cam = ol.make_camera((512.0, 1.5 * mh, 512.0), (0.0, -0.9, 1.0))
    opts = ol.make_opts(mh)
    got = gpulib.gpu_trace(cuda_ctx, W, H, [cam], opts)
    want = ol.cpu_trace(ol.oracle().hmrt_oracle_trace, want_pyr, None, 8, 8, W, H, cam, opts)
    ol.assert_same_trace((got[0][0], got[1][0]), want, "config 1 vs oracle")
    if ol.ref() is not None:
        ref = ol.cpu_trace(ol.ref().hmrt_ref_trace, want_pyr, None, 8, 8, W, H, cam, opts)
        ol.assert_same_trace((got[0][0], got[1][0]), ref, "config 1 vs the reference's own code")
    assert (got[1][0]["flags"] & 1).mean() > 0.9


def test_gpu_rasteriser_and_window_equal_the_reference_host_code_directly(cuda_ctx):
    """No restatement in between: the CUDA scatter + mip build + colour resolve against the reference's OWN loader
    (allocateSection + loadLASToSection, main.cpp:174-269) and hmrt_compose_window against the reference's OWN
    preparePointBuffer (main.cpp:459-618), both from oracle/_ref/libhmrt_refhost.so (verbatim text, see tests/test_refhost.py)."""
    import refhostlib

    rh = refhostlib.refhost("")
    if rh is None:
        pytest.skip("oracle/_ref/libhmrt_refhost.so not built")
    coarse, levels = 2, rh.levels
    res, idx, total = rh.config(coarse)
    for fmt, origin in [(2, (0.0, 0.0)), (3, (17.0, -9.5)), (0, (0.0, 0.0))]:
        hdr, rec = rl.synthetic_las(60_000, res[0], point_format=fmt, seed=70 + fmt)
        rh.set_las(hdr, rec)
        want_p, want_c = rh.rasterise(origin)
        got_p, got_c = _gpu_rasterise(cuda_ctx, hdr, rec, coarse, levels, origin=origin)
        assert (got_p.view(np.uint32) == want_p.view(np.uint32)).all(), (fmt, origin)
        assert (got_c == want_c).all(), (fmt, origin)
    # window: the reference's 4 x 4 sections filled with tagged content, the camera somewhere in the inner sections
    import hmrt

    rng = np.random.default_rng(9)
    cam0 = np.array([517.25, 33.0, -90.5], np.float32)
    origins, _, _ = rh.init_sections(cam0)
    secs = {}
    for i in range(rh.grid):
        for j in range(rh.grid):
            pyr = (rng.random(total, dtype=np.float32) + np.float32(10 * (i * rh.grid + j))).astype(np.float32)
            col = rng.integers(0, 256, (res[0], res[0], 3), dtype=np.uint8)
            rh.fill_section(i, j, pyr, col)
            secs[i, j] = (torch.from_numpy(pyr).cuda(), torch.from_numpy(col).cuda())
    size = float(coarse << (levels - 1))
    for k in range(6):
        cam = np.array([origins[1, 1, 0] + rng.uniform(0, 2 * size * 0.999), 20.0, origins[1, 1, 1] + rng.uniform(0, 2 * size * 0.999)], np.float32)
        want_cpb, want_pyr, want_col = rh.prepare(cam)
        pl = hmrt.window_place(cam, origins, rh.grid, coarse, levels)
        assert np.array_equal(np.array(pl.camera, np.float32).view(np.uint32), want_cpb.view(np.uint32))
        xs, ys = (pl.min_x, pl.max_x), (pl.min_y, pl.max_y)
        dp = [[secs[xs[a], ys[b]][0] for b in range(2)] for a in range(2)]
        dc = [[secs[xs[a], ys[b]][1] for b in range(2)] for a in range(2)]
        out = torch.empty(total, dtype=torch.float32, device="cuda")
        out_c = torch.empty((res[0], res[0], 3), dtype=torch.uint8, device="cuda")
        cuda_ctx.compose_window(dp, dc, coarse, levels, pl.cell_x, pl.cell_y, out, out_c)
        torch.cuda.synchronize()
        assert (out.cpu().numpy().view(np.uint32) == want_pyr.view(np.uint32)).all(), k
        assert (out_c.cpu().numpy() == want_col).all(), k
    rh.config(1)


def test_heights_below_the_header_minimum_and_negative_zero(cuda_ctx):
    """Points below header.GetMinZ() leave no height (main.cpp:229 on a zero-initialised section) but DO write their colour
    (:223-224 precede the test); `-0.0f` is the one defined deviation: the reference stores it (sign follows file order),
    the integer atomicMax leaves +0.0f -- equal values, no sign bit anywhere (tests/test_refhost.py pins the reference side)."""
    from hmrt import las

    coarse, levels = 1, 8
    res, idx, total = ol.pyramid_layout(coarse, levels)
    r0 = res[0]
    X = np.array([1, 1, 5, 5, 9, 9, 13, 13, 13, 17], np.int32)
    Y = np.ones(10, np.int32)
    Z = np.array([-5000, 3000, 4000, -1, -7, -9, 0, 0, 0, -2000], np.int32)
    rec = las.encode_points(X, Y, Z, 2, np.zeros(10, np.uint8), np.full((10, 3), 65535, np.uint16))
    hdr = las.LasHeader(2, rec.shape[1], len(X), (1.0, 1.0, 1e-3), (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), (2.0 * r0, 2.0 * r0, 4.0))
    want_p, want_c = rl.oracle_rasterise(hdr, rec, coarse, levels)
    got_p, got_c = _gpu_rasterise(cuda_ctx, hdr, rec, coarse, levels)
    assert (got_p.view(np.uint32) == want_p.view(np.uint32)).all() and (got_c == want_c).all()
    assert (got_c[0, 4] == 255).all() and got_p[idx[0]:].reshape(r0, r0)[0, 4] == 0
    # -0.0f: (float)(-1e-50) == -0.0f
    hdr2 = las.LasHeader(2, rec.shape[1], 3, (1.0, 1.0, 1e-50), (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), (2.0 * r0, 2.0 * r0, 4.0))
    rec2 = las.encode_points(np.array([1, 5, 5], np.int32), np.ones(3, np.int32), np.array([-1, -1, 0], np.int32), 2, np.zeros(3, np.uint8),
                             np.zeros((3, 3), np.uint16))
    want2, _ = rl.oracle_rasterise(hdr2, rec2, coarse, levels, with_colors=False)
    got2, _ = _gpu_rasterise(cuda_ctx, hdr2, rec2, coarse, levels, with_colors=False)
    assert np.signbit(want2).any() and not np.signbit(got2).any()
    assert (got2 == want2).all()  # -0.0 == +0.0: the same heights
