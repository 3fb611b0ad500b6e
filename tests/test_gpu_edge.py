"""Edge cases of the traversal on the GPU (through the C ABI) against the CPU oracle, bit-exact:
empty map, max_height below the terrain, NaN / inf cells, cameras outside the grid, degenerate directions,
and the largest configuration of BASELINE.json (32768^2)."""
import numpy as np
import pytest
import torch

import oraclelib as ol

pytestmark = pytest.mark.gpu


def _both(cuda_ctx, sc, W, H, cams, opts):
    import gpulib

    keep = gpulib.upload_scene(cuda_ctx, sc)  # noqa: F841
    rgb, hits = gpulib.gpu_trace(cuda_ctx, W, H, cams, opts)
    for i, cam in enumerate(cams):
        want = ol.cpu_trace(ol.oracle().hmrt_oracle_trace, sc["pyramid"], sc.get("color_map"), sc["coarse"], sc["levels"], W, H, cam, opts)
        ol.assert_same_trace((rgb[i], hits[i]), want, f"cam{i}")
    return rgb, hits


def _scene_from(fin, levels):
    r0 = fin.shape[0]
    return dict(r0=r0, levels=levels, coarse=r0 >> (levels - 1), finest=fin, pyramid=ol.pyramid_from_finest(fin, levels), color_map=None,
                max_height=float(np.nanmax(np.where(np.isfinite(fin), fin, 0))))


def test_empty_heightmap(cuda_ctx):
    """All cells +0 (a section no point fell into, main.cpp:259): looking down every ray lands on the floor."""
    sc = _scene_from(np.zeros((256, 256), np.float32), 6)
    cams = [ol.make_camera((128.3, 40.0, 128.7), (0.2, -0.8, 0.5)), ol.make_camera((128.3, 40.0, 128.7), (0.2, 0.3, 0.5))]
    rgb, hits = _both(cuda_ctx, sc, 128, 96, cams, ol.make_opts(10.0, shadows=True))
    assert (hits[0]["flags"] & 1).all() and not (hits[1]["flags"] & 1).any()


def test_max_height_below_terrain(cuda_ctx):
    """'-' key of the reference (main.cpp:842-844): max_height below the real maximum changes ray termination (:153)
    and pushes the colour ramp out of range (float -> unsigned char wrap of the host build)."""
    sc = ol.scene("r512_l4", seed=8)
    cams = ol.cameras_for(sc, 4)
    for factor in (0.5, 0.25):
        _both(cuda_ctx, sc, 160, 120, cams, ol.make_opts(sc["max_height"] * factor, shadows=True))


def test_nan_and_inf_cells(cuda_ctx):
    """A NaN cell never intersects (`y <= NaN` is false); an +inf column always does.  The air phase must notice
    both through its top-level maximum (NaN disables it, +inf makes every ray descend)."""
    base = ol.sines_terrain(256, seed=3)
    for poison in (np.nan, np.inf):
        fin = base.copy()
        fin[100:103, 120:124] = poison
        sc = _scene_from(fin, 6)
        # numpy's max propagates NaN like the reference's `buf <= z` would not; rebuild the pyramid with the reference rule
        pyr = np.zeros_like(sc["pyramid"])
        res, idx, total = ol.pyramid_layout(sc["coarse"], 6)
        pyr[idx[0]:] = fin.ravel()
        assert ol.oracle().hmrt_oracle_build_mips(pyr.ctypes.data, sc["coarse"], 6) == 0
        sc["pyramid"] = pyr
        sc["max_height"] = float(base.max())
        cams = [ol.make_camera((128.4, 2.0 * sc["max_height"], 60.2), (0.0, -0.6, 1.0)), ol.make_camera((20.5, 1.2 * sc["max_height"], 110.5), (1.0, -0.1, 0.0))]
        _both(cuda_ctx, sc, 160, 120, cams, ol.make_opts(sc["max_height"], shadows=True))


def test_cameras_outside_and_degenerate_directions(cuda_ctx):
    sc = ol.scene("r256_l1", seed=1)
    sc["color_map"] = None
    mh = sc["max_height"]
    cams = [ol.make_camera((-40.0, mh, 128.0), (1.0, -0.2, 0.0)),      # outside, looking in along +x (dir.z == 0)
            ol.make_camera((128.0, mh, 300.0), (0.0, -0.2, -1.0)),     # outside on the far side, looking back (dir.x == 0)
            ol.make_camera((128.0, 3 * mh, 128.0), (0.0, -1.0, 1e-7)),  # straight down
            ol.make_camera((400.0, mh, 400.0), (1.0, -0.1, 1.0)),      # outside, looking away
            ol.make_camera((128.5, 0.5 * mh, 128.5), (0.3, 0.0, 0.7))]  # horizontal (dir.y == 0 at the image centre)
    _both(cuda_ctx, sc, 161, 121, cams, ol.make_opts(mh, shadows=True))


def test_frame_dimensions_fast_and_fallback_setup(cuda_ctx):
    """The pixel -> ray set-up divides by (W - 1) and (H - 1) through host reciprocals when every frame_dimension of the
    launch lies in [2^-40, 2^40], and takes the IEEE divisions otherwise: both must be the reference's bits.  Wide / narrow /
    negative (mirrored image plane) frame dimensions, one launch mixing an ordinary and an out-of-range camera (the whole
    launch falls back), and frame sizes whose (W - 1), (H - 1) are not powers of two."""
    sc = ol.scene("r512_l4", seed=5)
    mh = sc["max_height"]
    pos, fwd = (256.3, 1.6 * mh, 250.7), (0.4, -0.5, 0.77)
    dims = [(32.0, 18.0, 20.0), (3.0, 40.0, 7.0), (-32.0, 18.0, 20.0), (1e-3, 2e-3, 1.5e-3), (5e11, 3e11, 4e11),   # fast path
            (1e-13, 1e-13, 1e-13), (3e12, 18.0, 20.0)]                                                            # fallback
    for W, H in ((160, 120), (97, 61)):
        for fd in dims:
            _both(cuda_ctx, sc, W, H, [ol.make_camera(pos, fwd, fd)], ol.make_opts(mh, shadows=True))
        _both(cuda_ctx, sc, W, H, [ol.make_camera(pos, fwd, dims[0]), ol.make_camera(pos, fwd, dims[6]), ol.make_camera(pos, fwd, dims[1])],
              ol.make_opts(mh))


def test_largest_configuration_32768(cuda_ctx):
    """BASELINE configs[4] shape: 32768^2 heightmap (5.73 GB pyramid), 4K, primary + shadow rays.  Sampled rows against
    the CPU oracle at full size; sharded == whole."""
    import gpulib
    from hmrt import dist as hd

    r0, levels = 32768, 8
    coarse = r0 >> (levels - 1)
    res, idx, total = ol.pyramid_layout(coarse, levels)
    pyr = torch.zeros(total, dtype=torch.float32, device="cuda")
    fin = pyr[idx[0]:].view(r0, r0)
    xs = torch.arange(r0, device="cuda", dtype=torch.float32)
    for z0 in range(0, r0, 4096):
        fin[z0:z0 + 4096].copy_(500 + 300 * torch.sin(xs[None, :] * 0.0007) * torch.cos(xs[z0:z0 + 4096, None] * 0.0005)
                                + 80 * torch.sin(xs[None, :] * 0.011 + xs[z0:z0 + 4096, None] * 0.009))
    fin.clamp_(min=0)
    cuda_ctx.build_mips(pyr, coarse, levels)
    torch.cuda.synchronize()
    mh = float(pyr[: coarse * coarse].max())
    cuda_ctx.set_heightmap(pyr, None, coarse, levels, mh)
    W, H = 3840, 2160
    cam = ol.make_camera((r0 * 0.3, mh + 2500.0, r0 * 0.35), (0.7, -0.25, 0.6))
    opts = ol.make_opts(mh, shadows=True)
    whole, hits = gpulib.gpu_trace(cuda_ctx, W, H, [cam], opts)
    parts = []
    for r in range(8):
        p, _ = gpulib.gpu_trace(cuda_ctx, W, H, [cam], ol.make_opts(mh, shadows=True, tile_first=r, tile_stride=8), hits=False)
        parts.append(torch.from_numpy(p)[0])
    assert (hd.assemble_frame(parts, H, W).numpy() == whole[0]).all()
    host = pyr.cpu().numpy()
    sc = dict(pyramid=host, coarse=coarse, levels=levels)
    for r0_, r1_ in [(0, 3), (1080, 1084), (2157, 2160)]:
        ergb, ehits = ol.cpu_trace(ol.oracle().hmrt_oracle_trace, host, None, coarse, levels, W, H, cam, opts, rows=(r0_, r1_))
        ol.assert_same_trace((whole[0][r0_:r1_], hits[0][r0_:r1_]), (ergb[r0_:r1_], ehits[r0_:r1_]), f"32768 rows {r0_}")
    assert (hits[0]["flags"] & 1).mean() > 0.3


def test_full_frame_output_layout(cuda_ctx):
    """hmrt_trace_opts.full_frame_output: tile-sharded calls store their tiles at their place in ONE whole-frame buffer
    (what the ranks of a multi-GPU render do into rank 0's frame); together they reproduce the unsharded call, rows of
    other shards are never touched."""
    import gpulib
    import hmrt

    sc = ol.scene("r512_l4", seed=9)
    gpulib.upload_scene(cuda_ctx, sc)
    W, H = 200, 83  # ragged: 11 tiles, the last one 3 rows; W % 16 != 0 -> byte stores
    cams = ol.cameras_for(sc, 2)
    whole, _ = cuda_ctx.trace(W, H, cams, hmrt.trace_opts(sc["max_height"], shadows=True))
    buf = torch.full((2, H, W, 3), 9, dtype=torch.uint8, device="cuda")
    for r in (2, 0):
        cuda_ctx.trace(W, H, cams, hmrt.trace_opts(sc["max_height"], shadows=True, tile_first=r, tile_stride=3, full_frame_output=True), out=buf)
    torch.cuda.synchronize()
    rows_1 = [y for t in range(1, 11, 3) for y in range(t * 8, min(H, t * 8 + 8))]
    assert (buf[:, rows_1] == 9).all(), "rows of the missing shard were written"
    cuda_ctx.trace(W, H, cams, hmrt.trace_opts(sc["max_height"], shadows=True, tile_first=1, tile_stride=3, full_frame_output=True), out=buf)
    torch.cuda.synchronize()
    assert torch.equal(buf, whole)
    with pytest.raises(hmrt.HmrtError):
        cuda_ctx.trace_host(W, H, cams, hmrt.trace_opts(sc["max_height"], full_frame_output=True), np.zeros((2, H, W, 3), np.uint8))


def test_copy_tiles_to_frames(cuda_ctx):
    """hmrt_copy_tiles_to_frames: the compact output of every shard copied to its place reassembles the unsharded frames
    (ragged last tile owned by one of the shards, several frames, W not a multiple of 16)."""
    import gpulib
    import hmrt

    sc = ol.scene("r512_l4", seed=5)
    gpulib.upload_scene(cuda_ctx, sc)
    for W, H in ((200, 83), (320, 96)):
        cams = ol.cameras_for(sc, 3)
        whole, _ = cuda_ctx.trace(W, H, cams, hmrt.trace_opts(sc["max_height"]))
        buf = torch.full((3, H, W, 3), 9, dtype=torch.uint8, device="cuda")
        for r in range(3):
            part, _ = cuda_ctx.trace(W, H, cams, hmrt.trace_opts(sc["max_height"], tile_first=r, tile_stride=3))
            cuda_ctx.copy_tiles_to_frames(part, buf, W, H, 3, r, 3)
        torch.cuda.synchronize()
        assert torch.equal(buf, whole), (W, H)
