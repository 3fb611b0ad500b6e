#!/usr/bin/env python
"""BASELINE.json configs[0], end to end on identical inputs: seeded PointdataGenerator terrain (N = 1024, 1025^2 points)
-> 1024^2 heightmap with 8 levels -> one 640x480 primary-ray frame, camera at the grid centre at 1.5 x max height looking
along normalize(0, -0.9, 1) (main.cpp:56-57).

  reference CPU path   the reference's own code compiled for the host (oracle/_ref; the plain-C oracle if absent): rasterisation
                       by the restated loader loop, traversal on 1 thread and on all host threads
  new GPU path         libhmrt.so through the C ABI: scatter + fused max-mipmap build + traversal, frame copied back

Prints one JSON line with the timings and whether the two frames (and the two heightmaps) are bit-identical.
"""
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "gpu-heightmap-raytracer_b200"))
sys.path.insert(0, str(REPO / "tests"))

import torch  # noqa: E402

import hmrt  # noqa: E402
import oraclelib as ol  # noqa: E402  (checker + reported CPU baseline only)


def main():
    n, levels, W, H = 1024, 8, 640, 480
    coarse = n >> (levels - 1)
    res, idx, total = ol.pyramid_layout(coarse, levels)
    xyz = np.zeros(((n + 1) ** 2, 3), np.float32)
    assert ol.oracle().hmrt_oracle_pdg_generate(n, 7, xyz.ctypes.data) == 0
    xf = hmrt.LasTransform()
    xf.scale[:] = (1.0, 1.0, 1.0)
    xf.cell_size[:] = (1.0, 1.0, 1.0)

    # ---- reference CPU path ----
    pyr = np.zeros(total, np.float32)
    t0 = time.perf_counter()
    assert ol.oracle().hmrt_oracle_rasterise_xyz(xyz.ctypes.data, len(xyz), C.byref(xf), pyr.ctypes.data, coarse, levels) == 0
    cpu_raster_ms = 1e3 * (time.perf_counter() - t0)
    mh = float(pyr[: coarse * coarse].max())
    cam = ol.make_camera((n * 0.5, 1.5 * mh, n * 0.5), (0.0, -0.9, 1.0))
    opts = ol.make_opts(mh)
    lib = ol.ref()
    kind, fn = ("reference", lib.hmrt_ref_trace) if lib is not None else ("port", ol.oracle().hmrt_oracle_trace)
    cpu = {}
    for threads in (1, os.cpu_count() or 1):
        t0 = time.perf_counter()
        want, _ = ol.cpu_trace(fn, pyr, None, coarse, levels, W, H, cam, opts, n_threads=threads, want_hits=False)
        cpu[threads] = 1e3 * (time.perf_counter() - t0)

    # ---- new GPU path ----
    ctx = hmrt.Context(0)
    d_xyz = torch.from_numpy(xyz).cuda()
    d_pyr = torch.empty(total, dtype=torch.float32, device="cuda")
    host_fb = torch.empty((1, H, W, 3), dtype=torch.uint8).pin_memory()

    def gpu_pass():
        ctx.clear_section(d_pyr, coarse, levels)
        ctx.scatter_xyz(d_xyz, len(xyz), xf, d_pyr, coarse, levels)
        ctx.build_mips(d_pyr, coarse, levels)
        ctx.set_heightmap(d_pyr, None, coarse, levels, mh)
        ctx.trace_host(W, H, [cam], opts, host_fb)

    for _ in range(3):
        gpu_pass()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 50
    for _ in range(reps):
        gpu_pass()
    torch.cuda.synchronize()
    gpu_ms = 1e3 * (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.trace_host(W, H, [cam], opts, host_fb)
    gpu_trace_ms = 1e3 * (time.perf_counter() - t0) / reps
    same_map = bool((d_pyr.cpu().numpy().view(np.uint32) == pyr.view(np.uint32)).all())
    same_frame = bool((host_fb.numpy()[0] == want).all())
    print(json.dumps({
        "workload": f"PointdataGenerator N={n} ({len(xyz)} points) -> {n}^2 heightmap, {levels} levels, one {W}x{H} frame",
        "cpu_reference_path": {"kind": kind, "rasterise_ms": cpu_raster_ms, "trace_ms_1_thread": cpu[1],
                               f"trace_ms_{max(cpu)}_threads": cpu[max(cpu)], "Mrays_per_s_1_thread": W * H / cpu[1] / 1e3},
        "gpu_path": {"rasterise_plus_mips_plus_trace_to_host_ms": gpu_ms, "trace_to_host_ms": gpu_trace_ms,
                     "Mrays_per_s_trace_to_host": W * H / gpu_trace_ms / 1e3},
        "heightmap_bit_identical": same_map, "frame_bit_identical": same_frame}))
    return 0 if same_map and same_frame else 1


if __name__ == "__main__":
    sys.exit(main())
