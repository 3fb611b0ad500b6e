#!/usr/bin/env python
"""Interactive-use latency on the reference's DEFAULT workload (main.cpp:52,71,83): one 1920x1080 frame over a
4096^2 window with 8 levels, one synchronous call per frame like the reference's draw loop.  Reports ms/frame for the
new library (hmrt_trace + synchronize) and, when oracle/_ref/libhmrt_ref_gpu.so is present, for the reference's own
CUDA kernel recompiled for sm_100a.  Also BASELINE config 1's GPU side (640x480 over 1024^2)."""
import ctypes as C
import json
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "gpu-heightmap-raytracer_b200"))

import torch  # noqa: E402

import hmrt  # noqa: E402


def run(ctx, r0, W, H, frames=200):
    levels = 8
    coarse = r0 >> (levels - 1)
    res, idx, total = hmrt.pyramid_layout(coarse, levels)
    pyr = torch.zeros(total, dtype=torch.float32, device="cuda")
    xs = torch.arange(r0, device="cuda", dtype=torch.float32)
    fin = pyr[idx[0]:].view(r0, r0)
    fin.copy_(60 + 35 * torch.sin(xs[None, :] * 0.013) * torch.cos(xs[:, None] * 0.017) + 12 * torch.sin(xs[None, :] * 0.11 + xs[:, None] * 0.07))
    fin.clamp_(min=0)
    ctx.build_mips(pyr, coarse, levels)
    mh = float(fin.max())
    ctx.set_heightmap(pyr, None, coarse, levels, mh)
    opts = hmrt.trace_opts(mh)
    fb = torch.empty((1, H, W, 3), dtype=torch.uint8, device="cuda")
    cams = [hmrt.camera((r0 / 2 + 30 * (i % 7), 1.5 * mh, r0 / 2 - 20 * (i % 5)), (0.05 * (i % 9), -0.9 + 0.08 * (i % 10), 1.0)) for i in range(frames)]
    for c in cams[:10]:
        ctx.trace(W, H, c, opts, out=fb)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for c in cams:
        ctx.trace(W, H, c, opts, out=fb)
        ctx.synchronize()  # the reference synchronises every frame (CudaKernel.cu:307)
    ours = (time.perf_counter() - t0) / frames * 1e3
    ref = None
    so = REPO / "oracle" / "_ref" / "libhmrt_ref_gpu.so"
    if so.exists():
        lib = C.CDLL(str(so))
        lib.hmrt_refgpu_trace.restype = C.c_int
        lib.hmrt_refgpu_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_int,
                                          C.POINTER(C.c_float)]
        ms, tot = C.c_float(), 0.0
        for c in cams[:50]:
            assert lib.hmrt_refgpu_trace(pyr.data_ptr(), None, coarse, levels, W, H, C.byref(c), 0, C.c_float(mh), fb.data_ptr(), 1, C.byref(ms)) == 0
            tot += ms.value
        ref = tot / 50
    return {"workload": f"{W}x{H} over {r0}^2, 8 levels, one synchronous call per frame", "ms_per_frame": ours, "fps": 1e3 / ours,
            "Mrays_per_s": W * H / ours / 1e3, "reference_cuda_kernel_ms_per_frame_device": ref,
            "reference_cuda_kernel_Mrays_per_s": (W * H / ref / 1e3) if ref else None}


def run_windowed(ctx, W=1920, H=1080, frames=200):
    """The reference's whole per-frame loop at its defaults (main.cpp:947-966): place the camera window over the 4 x 4 grid of
    resident 4096^2 sections, compose it on the device (preparePointBuffer + copyPointBuffer), rayTrace in window coordinates,
    synchronise -- per frame, while the camera walks across section borders."""
    coarse, levels, grid = 32, 8, 4
    res, idx, total = hmrt.pyramid_layout(coarse, levels)
    r0 = res[0]
    xs = torch.arange(r0, device="cuda", dtype=torch.float32)
    secs, cols = {}, {}
    for i in range(grid):  # 16 resident sections: 16 x 139.8 MB = 2.2 GB of HBM
        for j in range(grid):
            pyr = torch.zeros(total, dtype=torch.float32, device="cuda")
            fin = pyr[idx[0]:].view(r0, r0)
            x, z = xs[None, :] + i * r0, xs[:, None] + j * r0
            fin.copy_(60 + 35 * torch.sin(x * 0.013) * torch.cos(z * 0.017) + 12 * torch.sin(x * 0.11 + z * 0.07))
            fin.clamp_(min=0)
            ctx.build_mips(pyr, coarse, levels)
            secs[i, j] = pyr
            cols[i, j] = torch.full((r0, r0, 3), 40 * (i + j) % 255, dtype=torch.uint8, device="cuda")
    org = [[(float(i * r0), float(j * r0)) for j in range(grid)] for i in range(grid)]
    win = torch.empty(total, dtype=torch.float32, device="cuda")
    wcol = torch.empty((r0, r0, 3), dtype=torch.uint8, device="cuda")
    mh = float(max(float(s[: coarse * coarse].max()) for s in secs.values()))
    opts = hmrt.trace_opts(mh, use_color_map=True)
    fb = torch.empty((1, H, W, 3), dtype=torch.uint8, device="cuda")

    def frame(k):
        cam_world = (2.0 * r0 + 37.0 * k % (0.9 * r0), 1.5 * mh, 2.0 * r0 - 23.0 * k % (0.9 * r0))
        pl = hmrt.window_place(cam_world, org, grid, coarse, levels)
        pick = lambda d: [[d[pl.min_x, pl.min_y], d[pl.min_x, pl.max_y]], [d[pl.max_x, pl.min_y], d[pl.max_x, pl.max_y]]]  # noqa: E731
        ctx.compose_window(pick(secs), pick(cols), coarse, levels, pl.cell_x, pl.cell_y, win, wcol)
        ctx.set_heightmap(win, wcol, coarse, levels, mh)
        ctx.trace(W, H, hmrt.camera(tuple(pl.camera), (0.05 * (k % 9), -0.9 + 0.08 * (k % 10), 1.0)), opts, out=fb)
        ctx.synchronize()

    for k in range(10):
        frame(k)
    t0 = time.perf_counter()
    for k in range(frames):
        frame(k)
    ms = (time.perf_counter() - t0) / frames * 1e3
    return {"workload": f"{W}x{H}, 4 x 4 resident sections of {r0}^2 cells, window composed + traced + synchronised per frame (colour-map shading)",
            "ms_per_frame": ms, "fps": 1e3 / ms}


if __name__ == "__main__":
    ctx = hmrt.Context(0)
    print(json.dumps({"reference_default": run(ctx, 4096, 1920, 1080), "config1_gpu": run(ctx, 1024, 640, 480),
                      "reference_main_loop": run_windowed(ctx)}))
