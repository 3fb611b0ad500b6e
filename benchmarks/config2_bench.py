#!/usr/bin/env python
"""BASELINE.json configs[1]: 4096^2 synthetic heightmap, 1920x1080 primary rays + shading, 1 B200, in the two
traversal modes SURVEY.md section 8(d) names: L = 8 (max-mipmap) and L = 1 ("plain DDA": the same walk with a
single level, CudaKernel.cu.rej).  The 4096^2 map (64 MiB, 89.5 MB pyramid) is L2-resident, so the roofline
denominator is the L2 read bandwidth measured by benchmarks/l2_bandwidth.cu on the same box (pass --l2-gbps)."""
import argparse
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "gpu-heightmap-raytracer_b200"))

import torch  # noqa: E402

import hmrt  # noqa: E402

R0, W, H, POSES = 4096, 1920, 1080, 16


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--l2-gbps", type=float, default=None,
                    help="L2 read bandwidth of the box; default: the committed measurement of benchmarks/l2_bandwidth.cu (profiles/raw_r01/l2_bandwidth.json)")
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    l2_source = "--l2-gbps"
    if args.l2_gbps is None:
        f = REPO / "profiles" / "raw_r01" / "l2_bandwidth.json"
        if f.exists():
            args.l2_gbps = float(json.loads(f.read_text())["l2_read_GBps"])
            l2_source = "profiles/raw_r01/l2_bandwidth.json (benchmarks/l2_bandwidth.cu, 48 MiB resident set, same pool)"
    ctx = hmrt.Context(0)
    xs = torch.arange(R0, device="cuda", dtype=torch.float32)
    fin0 = (160 + 90 * torch.sin(xs[None, :] * 0.0049) * torch.cos(xs[:, None] * 0.0039) + 35 * torch.sin(xs[None, :] * 0.019 + xs[:, None] * 0.016)
            + 8 * torch.sin(xs[None, :] * 0.11) * torch.sin(xs[:, None] * 0.09)).clamp_(min=0)
    out = {}
    for levels in (8, 1):
        coarse = R0 >> (levels - 1)
        res, idx, total = hmrt.pyramid_layout(coarse, levels)
        pyr = torch.zeros(total, dtype=torch.float32, device="cuda")
        pyr[idx[0]:].view(R0, R0).copy_(fin0)
        ctx.build_mips(pyr, coarse, levels)
        mh = float(fin0.max())
        ctx.set_heightmap(pyr, None, coarse, levels, mh)
        opts = hmrt.trace_opts(mh)
        batches = []
        for s in range(args.steps + 2):
            cams = []
            for i in range(POSES):
                k = s * POSES + i
                u, v = (k * 0.618034) % 1.0, (k * 0.754878) % 1.0
                import math
                hd = 2 * math.pi * ((k * 0.56984) % 1.0)
                cams.append(hmrt.camera((R0 * (0.25 + 0.5 * u), mh + 500 + 700 * ((k * 0.24512) % 1.0), R0 * (0.25 + 0.5 * v)),
                                        (math.cos(hd), -0.15 - 0.5 * ((k * 0.38197) % 1.0), math.sin(hd))))
            batches.append(hmrt.context._cam_array(cams))
        fb = torch.empty((POSES, H, W, 3), dtype=torch.uint8, device="cuda")
        hits = torch.empty((POSES, H, W, 4), dtype=torch.int32, device="cuda")
        for b in batches[:2]:
            ctx.trace(W, H, b, opts, out=fb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for b in batches[2:]:
            ctx.trace(W, H, b, opts, out=fb)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        iters = 0
        for b in batches[2:]:
            ctx.trace(W, H, b, opts, out=fb, hits=hits)
            iters += int((hits[..., 3] >> 8).to(torch.int64).sum().item())
        rays = POSES * W * H * args.steps
        algo = 4 * iters + 3 * rays
        out[f"levels_{levels}"] = {"Mrays_per_s": rays / (ms * 1e-3) / 1e6, "ms_per_step": ms / args.steps, "iterations_per_ray": iters / rays,
                                   "algorithmic_GBps": algo / (ms * 1e-3) / 1e9,
                                   "frac_of_l2_peak": (algo / (ms * 1e-3) / 1e9 / args.l2_gbps) if args.l2_gbps else None}
        del pyr, fb, hits
    print(json.dumps({"workload": f"{R0}^2 heightmap, {W}x{H}, {POSES} poses per step, 1 GPU", "l2_read_GBps": args.l2_gbps, "l2_read_GBps_source": l2_source if args.l2_gbps else None, **out}))


if __name__ == "__main__":
    main()
