"""The TIMED configuration, whole frames, against the reference's own code (oracle/_ref = CudaKernel.cu host-compiled):
bench.py's terrain (same host generator, same bytes), bench.py's pose families, 3840x2160 over the 16384^2 map.

  exact mode (default)      every pixel, hit point, flag and iteration count equals the reference's  -> "pixels_differ == 0"
  tolerance mode (variant 2) measured against the acceptance bars BASELINE.json states for the path (hit cell >= 99.9 %, hit
                            distance within 1e-4 relative, colour within 1/255, >= 95 % pixel-exact).  It meets the colour
                            and pixel-exact bars everywhere but NOT the hit-cell bar on grazing views (99.8 % over a bench
                            pose batch: the reference's own accumulated rounding along ~50 air steps is what it cannot
                            reproduce), so it stays an opt-in mode that no reported number uses; the asserts below pin what
                            it does deliver.
One whole frame of each family through the reference costs ~3 s on 16 host cores."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

import oraclelib as ol

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))


@pytest.fixture(scope="module")
def bench_scene(cuda_ctx):
    import bench

    if ol.ref() is None:
        pytest.skip("oracle/_ref/libhmrt_ref.so not built")
    fin = bench.build_terrain_host()
    host_pyr = ol.pyramid_from_finest(fin, bench.LEVELS)
    pyr = torch.from_numpy(host_pyr).cuda()
    # the product's own mip kernel over the same finest level gives the same pyramid
    dev = pyr.clone()
    dev[: host_pyr.size - bench.R0 * bench.R0] = -1.0
    cuda_ctx.build_mips(dev, bench.COARSE, bench.LEVELS)
    torch.cuda.synchronize()
    assert torch.equal(dev.view(torch.int32), pyr.view(torch.int32))
    del dev
    mh = float(fin.max())
    cuda_ctx.set_heightmap(pyr, None, bench.COARSE, bench.LEVELS, mh)
    yield dict(bench=bench, host_pyr=host_pyr, pyr=pyr, mh=mh)
    cuda_ctx.set_trace_variant(0)


def _frame_pair(cuda_ctx, sc, family, step, pose):
    """(gpu rgb, gpu hits, ref rgb, ref hits, camera) of one whole 4K frame of a bench pose."""
    import gpulib

    bench = sc["bench"]
    pos, fwd = bench.pose_batch(step, sc["mh"], family)[pose]
    cam = ol.make_camera(pos, fwd, bench.FRAME_DIM)
    opts = ol.make_opts(sc["mh"])
    rgb, hits = gpulib.gpu_trace(cuda_ctx, bench.W, bench.H, [cam], opts)
    ergb, ehits = ol.cpu_trace(ol.ref().hmrt_ref_trace, sc["host_pyr"], None, bench.COARSE, bench.LEVELS, bench.W, bench.H, cam, opts)
    return rgb[0], hits[0], ergb, ehits, cam


@pytest.mark.parametrize("family,step,pose", [("high", 5, 3), ("high", 6, 11), ("low", 0, 2), ("low", 1, 9)])
def test_whole_4k_frames_of_the_bench_poses_equal_the_reference(cuda_ctx, bench_scene, family, step, pose):
    cuda_ctx.set_trace_variant(0)
    rgb, hits, ergb, ehits, _ = _frame_pair(cuda_ctx, bench_scene, family, step, pose)
    assert (hits["flags"] & 1).mean() > 0.2, "the pose must look at the terrain"
    ol.assert_same_trace((rgb, hits), (ergb, ehits), f"{family} pose {step}/{pose}")


@pytest.mark.parametrize("family,step,pose", [("high", 5, 3), ("high", 7, 0), ("low", 0, 2)])
def test_tolerance_mode_meets_the_baseline_bars(cuda_ctx, bench_scene, family, step, pose):
    bench = bench_scene["bench"]
    import gpulib

    cuda_ctx.set_trace_variant(2)
    try:
        rgb, hits, ergb, ehits, cam = _frame_pair(cuda_ctx, bench_scene, family, step, pose)
    finally:
        cuda_ctx.set_trace_variant(0)
    _, exact_hits = gpulib.gpu_trace(cuda_ctx, bench.W, bench.H, [cam], ol.make_opts(bench_scene["mh"]))  # exact walk: the iteration counts
    n = rgb.shape[0] * rgb.shape[1]
    cell_g, hit_g = ol.hit_cells(hits, bench.R0)
    cell_r, hit_r = ol.hit_cells(ehits, bench.R0)
    cell_match = (cell_g == cell_r).sum() / n
    assert cell_match >= 0.995, f"hit cell agreement {cell_match:.5f}"
    # hit distance from the camera, on pixels that hit the same cell (un-mirror first)
    def world(h):
        x = np.where((h["flags"] & 2) != 0, np.float32(bench.R0) - h["x"], h["x"]).astype(np.float64)
        z = np.where((h["flags"] & 4) != 0, np.float32(bench.R0) - h["z"], h["z"]).astype(np.float64)
        return np.stack([x, h["y"].astype(np.float64), z], axis=-1)
    c = np.array(list(cam.position), np.float64)
    both = hit_g & hit_r & (cell_g == cell_r)
    dg = np.linalg.norm(world(hits) - c, axis=-1)[both]
    dr = np.linalg.norm(world(ehits) - c, axis=-1)[both]
    rel = np.abs(dg - dr) / dr
    # same cell, possibly entered through another face: bounded by one cell over the hit distance; 1e-4 on >= 99.9 % of them
    assert (rel <= 1e-4).mean() >= 0.999 and rel.max() <= 1e-2, (float((rel <= 1e-4).mean()), float(rel.max()))
    col = np.abs(rgb.astype(np.int16) - ergb.astype(np.int16)).max(axis=-1)
    within = (col <= 1).sum() / n
    exact = (col == 0).sum() / n
    assert within >= 0.999 and exact >= 0.95, (within, exact)
    # the iteration statistics stay the reference algorithm's (one per boundary crossed): within 1 % in total
    sg, sr = ol.steps_of(hits).astype(np.int64).sum(), ol.steps_of(exact_hits[0]).astype(np.int64).sum()
    assert abs(sg - sr) / sr < 1e-2, (sg, sr)
    print(f"tolerance mode {family} {step}/{pose}: hit cell {100 * cell_match:.4f} %, distance within 1e-4 on {100 * (rel <= 1e-4).mean():.4f} % "
          f"(max {rel.max():.2e}), colour within 1/255 {100 * within:.4f} %, pixel-exact {100 * exact:.4f} %")
