"""The tile-binned rasterisation and its peer-memory exchange (csrc/rasterx.cu, hmrt_rx_*) through the C ABI.

One GPU (world 1, every kernel of the exchange path: TMA-staged binning, apply into the owned band, gather + mip build):
bit-exact against the restatement of loadLASToSection (main.cpp:193-234), which tests/test_refhost.py pins to the
reference's own text.  N GPUs: tests/multigpu/rx_worker.py under torchrun (skipped on a single-GPU box)."""
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

import oraclelib as ol
import rasterlib as rl

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent


def _rx_rasterise(ctx, hdr, rec, coarse, levels, cell=(2.0, 2.0, 2.0), origin=(0.0, 0.0), pieces=1, budget=None):
    lib = ctx.lib
    res, idx, total = ol.pyramid_layout(coarse, levels)
    pyr = torch.empty(total, dtype=torch.float32, device="cuda").fill_(-3.0)  # every float of it must be overwritten
    xf = hdr.transform(cell, origin)
    rx = C.c_void_p()
    rc = lib.hmrt_rx_create(ctx._h, coarse, levels, 0, 1, budget if budget is not None else len(rec), C.byref(rx))
    assert rc == 0, rc
    try:
        ctx._bind_stream()
        assert lib.hmrt_rx_begin(rx) == 0
        bounds = np.linspace(0, len(rec), pieces + 1).astype(np.int64)
        keep = []
        for a, b in zip(bounds[:-1], bounds[1:]):
            d = torch.from_numpy(np.ascontiguousarray(rec[a:b])).cuda()
            keep.append(d)
            assert lib.hmrt_rx_bin(rx, C.c_void_p(d.data_ptr()) if b > a else None, int(b - a), rec.shape[1], hdr.point_format, C.byref(xf)) == 0
        assert lib.hmrt_rx_barrier(rx) == 0
        assert lib.hmrt_rx_apply(rx) == 0
        assert lib.hmrt_rx_barrier(rx) == 0
        assert lib.hmrt_rx_gather_mips(rx, C.c_void_p(pyr.data_ptr())) == 0
        ov, er = C.c_uint32(), C.c_uint32()
        assert lib.hmrt_rx_status(rx, C.byref(ov), C.byref(er)) == 0
        return pyr.cpu().numpy(), ov.value, er.value
    finally:
        lib.hmrt_rx_destroy(rx)


@pytest.mark.parametrize("fmt,record_len,pieces", [(0, None, 1), (2, None, 1), (1, 28, 3), (3, 34, 2), (0, 32, 1), (2, 48, 1), (1, 64, 2), (3, 62, 1)])
def test_rx_world1_bit_exact(cuda_ctx, fmt, record_len, pieces):
    """20 / 26 / 28-byte records take 8 records per thread and step, 32..48-byte ones 4, the longest (up to 64 bytes) 2; aligned and
    2-mod-4 record lengths; several hmrt_rx_bin calls continue the slices."""
    r0, levels = 2048, 8
    coarse = r0 >> (levels - 1)
    n = 700_001 if pieces == 1 else 300_007  # not a multiple of the step size: exercises the partial last chunk and its tail bytes
    hdr, rec = rl.synthetic_las(n, r0, point_format=fmt, seed=10 + fmt, record_len=record_len)
    want, _ = rl.oracle_rasterise(hdr, rec, coarse, levels, with_colors=False)
    got, overflow, err = _rx_rasterise(cuda_ctx, hdr, rec, coarse, levels, pieces=pieces)
    assert overflow == 0 and err == 0
    assert (got.view(np.uint32) == want.view(np.uint32)).all()
    assert want[-r0 * r0:].max() > 1.0


def test_rx_origin_cell_size_and_levels(cuda_ctx):
    r0, levels = 4096, 6
    coarse = r0 >> (levels - 1)
    hdr, rec = rl.synthetic_las(400_000, r0, cell=1.5, seed=77)
    for origin in [(0.0, 0.0), (-10.5, 3.25)]:
        want, _ = rl.oracle_rasterise(hdr, rec, coarse, levels, cell=(1.5, 1.5, 0.75), origin=origin, with_colors=False)
        got, overflow, err = _rx_rasterise(cuda_ctx, hdr, rec, coarse, levels, cell=(1.5, 1.5, 0.75), origin=origin)
        assert overflow == 0 and err == 0
        assert (got.view(np.uint32) == want.view(np.uint32)).all()


def test_non_power_of_two_grid_in_every_scatter_mode(cuda_ctx):
    """coarse_res 21 (2688^2 cells: 6 x 6 grid tiles of 512 cells, the last ones ragged): direct, forced tile-binned and the
    exchange path agree with the oracle (the old binned path's shared-memory table grew with the tile count)."""
    r0, levels = 2688, 8
    coarse = r0 >> (levels - 1)
    assert coarse == 21
    hdr, rec = rl.synthetic_las(900_000, r0, point_format=0, seed=21)
    want, _ = rl.oracle_rasterise(hdr, rec, coarse, levels, with_colors=False)
    res, idx, total = ol.pyramid_layout(coarse, levels)
    d = torch.from_numpy(rec).cuda()
    xf = hdr.transform()
    try:
        for mode in (1, 2):
            cuda_ctx.set_scatter_mode(mode)
            pyr = torch.empty(total, dtype=torch.float32, device="cuda")
            cuda_ctx.clear_section(pyr, coarse, levels)
            cuda_ctx.scatter_las(d, len(rec), rec.shape[1], 0, xf, pyr, coarse, levels)
            cuda_ctx.build_mips(pyr, coarse, levels)
            torch.cuda.synchronize()
            assert (pyr.cpu().numpy().view(np.uint32) == want.view(np.uint32)).all(), f"scatter mode {mode}"
    finally:
        cuda_ctx.set_scatter_mode(0)
    # odd coarse resolution: level 0 starts at an odd float offset -> the exchange path declines, the host falls back
    rx = C.c_void_p()
    assert cuda_ctx.lib.hmrt_rx_create(cuda_ctx._h, coarse, levels, 0, 1, 1000, C.byref(rx)) == -3
    # 20 x 128 = 2560 cells (5 x 5 tiles of 512): the exchange path takes it
    hdr2, rec2 = rl.synthetic_las(700_000, 2560, point_format=0, seed=22)
    want2, _ = rl.oracle_rasterise(hdr2, rec2, 20, levels, with_colors=False)
    got, overflow, err = _rx_rasterise(cuda_ctx, hdr2, rec2, 20, levels)
    assert overflow == 0 and err == 0 and (got.view(np.uint32) == want2.view(np.uint32)).all()


@pytest.mark.parametrize("fmt,record_len", [(2, None), (3, None), (2, 32)])
def test_binned_scatter_carries_colour_keys(cuda_ctx, fmt, record_len):
    """A coloured unordered cloud through the tile-binned path (scatter mode 2): heights AND the colour map -- last writer in
    FILE order wins (main.cpp:223-224), also for points whose height loses or lies below the floor -- equal the oracle and
    the direct path, with the input in two chunks (first_index) and a shuffled copy of it."""
    r0, levels = 2048, 8
    coarse = r0 >> (levels - 1)
    res, idx, total = ol.pyramid_layout(coarse, levels)
    hdr, rec = rl.synthetic_las(600_000, r0, point_format=fmt, seed=40 + fmt, record_len=record_len)
    want_p, want_c = rl.oracle_rasterise(hdr, rec, coarse, levels)
    xf = hdr.transform()
    out = {}
    try:
        for mode in (1, 2):
            cuda_ctx.set_scatter_mode(mode)
            pyr = torch.empty(total, dtype=torch.float32, device="cuda")
            keys = torch.empty(r0 * r0, dtype=torch.int64, device="cuda")
            cmap = torch.empty((r0, r0, 3), dtype=torch.uint8, device="cuda")
            cuda_ctx.clear_section(pyr, coarse, levels, keys, cmap)
            half = len(rec) // 2 // 8 * 8  # second chunk 16-byte aligned for every record length used here
            for a, b in ((0, half), (half, len(rec))):
                d = torch.from_numpy(np.ascontiguousarray(rec[a:b])).cuda()
                cuda_ctx.scatter_las(d, b - a, rec.shape[1], fmt, xf, pyr, coarse, levels, first_index=a, color_keys=keys)
            cuda_ctx.build_mips(pyr, coarse, levels)
            cuda_ctx.resolve_colors(keys, cmap, r0 * r0)
            torch.cuda.synchronize()
            out[mode] = (pyr.cpu().numpy(), cmap.cpu().numpy())
    finally:
        cuda_ctx.set_scatter_mode(0)
    for mode in (1, 2):
        assert (out[mode][0].view(np.uint32) == want_p.view(np.uint32)).all(), f"heights, mode {mode}"
        assert (out[mode][1] == want_c).all(), f"colours, mode {mode}"


def test_rx_empty_input_and_overflow_report(cuda_ctx):
    r0, levels = 2048, 8
    coarse = r0 >> (levels - 1)
    hdr, rec = rl.synthetic_las(1000, r0, seed=1)
    got, overflow, err = _rx_rasterise(cuda_ctx, hdr, rec[:0], coarse, levels)
    assert overflow == 0 and err == 0 and (got.view(np.uint32) == 0).all()
    # every point in one tile and a point budget far too small: the slices overflow, the path reports it (the host then
    # takes the dense all-reduce route), and never writes out of bounds
    n = 400_000
    rng = np.random.default_rng(5)
    from hmrt import las
    X = rng.integers(0, 6000, n).astype(np.int32)
    Y = rng.integers(0, 6000, n).astype(np.int32)
    Z = rng.integers(0, 50000, n).astype(np.int32)
    rec2 = las.encode_points(X, Y, Z, 0)
    hdr2 = las.LasHeader(0, 20, n, (0.01, 0.01, 0.01), (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), (r0 * 2.0, r0 * 2.0, 500.0))
    got, overflow, err = _rx_rasterise(cuda_ctx, hdr2, rec2, coarse, levels, budget=1000)
    assert overflow > 0 and err == 0


def test_rx_shape_checks(cuda_ctx):
    lib = cuda_ctx.lib
    rx = C.c_void_p()
    assert lib.hmrt_rx_create(cuda_ctx._h, 8, 6, 0, 1, 1000, C.byref(rx)) == -3   # 256^2: tiles smaller than a mip tile
    assert lib.hmrt_rx_create(cuda_ctx._h, 16, 8, 2, 2, 1000, C.byref(rx)) == -1  # rank out of range
    assert lib.hmrt_rx_create(cuda_ctx._h, 16, 8, 0, 1, 1000, C.byref(rx)) == 0
    rows = (C.c_int * 2)()
    assert lib.hmrt_rx_bands(rx, rows) == 0 and list(rows) == [0, 2048]
    from hmrt import dist as hd
    assert hd.band_rows(2048, 1) == [0, 2048]
    assert lib.hmrt_rx_destroy(rx) == 0
    # the Python mirror of the band split agrees with the library for every world size (tile rows: 8 per axis when the
    # world size divides 8, else 16)
    for coarse, levels in ((16, 8), (32, 8), (24, 8), (128, 8)):
        for world in (1, 2, 3, 4, 5, 8, 16):
            rx = C.c_void_p()
            rc = lib.hmrt_rx_create(cuda_ctx._h, coarse, levels, 0, world, 1000, C.byref(rx))
            if rc != 0:
                assert rc == -3  # tile rows smaller than a mip tile
                continue
            rows = (C.c_int * (world + 1))()
            assert lib.hmrt_rx_bands(rx, rows) == 0
            assert list(rows) == hd.band_rows(coarse << (levels - 1), world), (coarse, levels, world)
            assert lib.hmrt_rx_destroy(rx) == 0


def test_raster_pipeline_single_modes_agree(cuda_ctx):
    """hmrt.dist.RasterPipeline on one GPU: the direct/auto path and the forced peer-exchange path give the same pyramid."""
    from hmrt import dist as hd

    r0, levels = 2048, 8
    coarse = r0 >> (levels - 1)
    hdr, rec = rl.synthetic_las(500_000, r0, point_format=0, seed=3)
    res, idx, total = ol.pyramid_layout(coarse, levels)
    d = torch.from_numpy(rec).cuda()
    xf = hdr.transform()
    out = []
    for force in (None, "peer"):
        rp = hd.RasterPipeline(cuda_ctx, coarse, levels, single=True, force_mode=force)
        pyr = torch.empty(total, dtype=torch.float32, device="cuda")
        t = rp.run(d, len(rec), rec.shape[1], 0, xf, pyr, timed=True)
        assert set(t) == set(rp.PHASES) and rp.mode == ("peer" if force else "single")
        out.append(pyr.cpu().numpy())
        rp.close()
    want, _ = rl.oracle_rasterise(hdr, rec, coarse, levels, with_colors=False)
    assert (out[0].view(np.uint32) == want.view(np.uint32)).all()
    assert (out[1].view(np.uint32) == want.view(np.uint32)).all()


def test_rx_multi_gpu_exchange():
    """2 (or more) ranks under torchrun: peer exchange == NCCL all-reduce path == one GPU, bit for bit."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    n = min(torch.cuda.device_count(), 8)
    env = dict(os.environ)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                          "--master-port", "29611", str(REPO / "tests" / "multigpu" / "rx_worker.py")], capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "rx_worker: ok" in out.stdout
