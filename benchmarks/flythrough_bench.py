#!/usr/bin/env python
"""BASELINE.json configs[4]: 32768^2 heightmap (5.73 GB pyramid, replicated per GPU), 64-frame 4K camera
fly-through with primary + shadow rays, row tiles interleaved over the GPUs of one box.

  python benchmarks/flythrough_bench.py                      # 1 GPU
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 benchmarks/flythrough_bench.py

Rays = primary rays (one per pixel) + shadow rays (one per primary hit whose biased origin is inside the grid);
both are counted from the instrumented kernel's flags on a separate, untimed pass.  One JSON line on rank 0.
"""
import argparse
import json
import math
import os
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "gpu-heightmap-raytracer_b200"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import hmrt  # noqa: E402
from hmrt import dist as hd  # noqa: E402


def terrain(ctx, r0, levels):
    coarse = r0 >> (levels - 1)
    res, idx, total = hmrt.pyramid_layout(coarse, levels)
    pyr = torch.zeros(total, dtype=torch.float32, device="cuda")
    fin = pyr[idx[0]:].view(r0, r0)
    xs = torch.arange(r0, device="cuda", dtype=torch.float32)
    band = 4096  # fill in row bands: the temporaries of a 32768^2 expression would need tens of GB
    for z0 in range(0, r0, band):
        x, z = xs[None, :], xs[z0:z0 + band, None]
        fin[z0:z0 + band].copy_(520 + 300 * torch.sin(x * 0.00061) * torch.cos(z * 0.00049) + 140 * torch.sin(x * 0.0023 + z * 0.0019)
                                + 60 * torch.sin(x * 0.0095) * torch.sin(z * 0.0115) + 14 * torch.sin(x * 0.055 + z * 0.035))
    fin.clamp_(min=0)
    ctx.build_mips(pyr, coarse, levels)
    torch.cuda.synchronize()
    return pyr, coarse, float(pyr[: coarse * coarse].max())


def spline(frames, r0, mh):
    """Deterministic fly-through: a slow arc over the map, descending, yawing, pitch -0.15..-0.45."""
    cams = []
    for i in range(frames):
        s = i / max(1, frames - 1)
        ang = 0.4 + 1.9 * s
        pos = (r0 * (0.5 + 0.28 * math.cos(ang)), mh + 3500.0 - 2200.0 * s, r0 * (0.5 + 0.28 * math.sin(ang)))
        heading = ang + 2.2
        fwd = (math.cos(heading), -0.15 - 0.3 * s, math.sin(heading))
        cams.append(hmrt.camera(pos, fwd))
    return cams


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=32768)
    ap.add_argument("--levels", type=int, default=8)
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--no-shadows", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = hmrt.Context(local)
    W, H, F = args.width, args.height, args.frames
    if rank == 0:
        pyr, coarse, mh = terrain(ctx, args.res, args.levels)
    else:
        coarse = args.res >> (args.levels - 1)
        pyr, mh = torch.empty(hmrt.pyramid_layout(coarse, args.levels)[2], dtype=torch.float32, device="cuda"), 0.0
    bcast_ms = None
    if world > 1:
        t = torch.tensor([mh], device="cuda")
        dist.broadcast(t, 0)
        mh = float(t.item())
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        hd.broadcast_pyramid(pyr, 0)
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
    ctx.set_heightmap(pyr, None, coarse, args.levels, mh)
    first, stride = hd.tiles_for_rank(rank, world)
    opts = hmrt.trace_opts(mh, shadows=not args.no_shadows, light_dir=(0.35, 0.6, 0.72), tile_first=first, tile_stride=stride)
    rows = hmrt.rows_local(H, first, stride)
    cams = hmrt.context._cam_array(spline(F, args.res, mh))
    fb = torch.empty((F, rows, W, 3), dtype=torch.uint8, device="cuda")

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    ctx.trace(W, H, cams, opts, out=fb)  # warm-up
    best = None
    for _ in range(args.reps):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.trace(W, H, cams, opts, out=fb)
        e1.record()
        sync()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    # end to end: every frame lands in pinned host memory
    host = torch.empty((F, rows, W, 3), dtype=torch.uint8).pin_memory()
    ctx.trace_host(W, H, cams, opts, host)
    sync()
    t0 = time.perf_counter()
    ctx.trace_host(W, H, cams, opts, host)
    e2e_s = time.perf_counter() - t0
    # ray census + iterations (untimed, chunks of 8 frames to bound the hit buffer)
    hits = torch.empty((8, rows, W, 4), dtype=torch.int32, device="cuda")
    shadow_rays = iters = hit_px = 0
    for f0 in range(0, F, 8):
        sub = hmrt.context._cam_array([cams[i] for i in range(f0, min(F, f0 + 8))])
        ctx.trace(W, H, sub, opts, out=fb[f0:f0 + len(sub)], hits=hits)
        fl = hits[: len(sub), ..., 3]
        hit_px += int((fl & 1).sum().item())
        iters += int((fl >> 8).to(torch.int64).sum().item())
    shadow_rays = hit_px if not args.no_shadows else 0
    vals = torch.tensor([best, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([shadow_rays, iters, hit_px], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms, e2e_ms = float(vals[0]), float(vals[1])
    shadow_rays, iters, hit_px = (int(v) for v in cnt)
    primary = F * W * H
    if rank == 0:
        print(json.dumps({
            "workload": f"{args.res}^2 heightmap ({args.levels} levels, {pyr.numel() * 4 / 1e9:.2f} GB pyramid), {F}-frame {W}x{H} fly-through, "
                        f"primary + {'shadow' if shadow_rays else 'no shadow'} rays, {world} GPU(s)",
            "frames_per_s": F / (ms * 1e-3), "ms_per_frame": ms / F,
            "Mrays_per_s": (primary + shadow_rays) / (ms * 1e-3) / 1e6, "primary_Mrays_per_s": primary / (ms * 1e-3) / 1e6,
            "e2e_frames_per_s": F / (e2e_ms * 1e-3), "e2e_Mrays_per_s": (primary + shadow_rays) / (e2e_ms * 1e-3) / 1e6,
            "primary_rays": primary, "shadow_rays_upper_bound": shadow_rays, "hit_fraction": hit_px / primary,
            "iterations_per_pixel": iters / primary, "pyramid_broadcast_ms": bcast_ms}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
