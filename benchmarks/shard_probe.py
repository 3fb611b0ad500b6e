#!/usr/bin/env python
"""One GPU doing ONE RANK'S share of an N-GPU bench step (tile_first = 0, tile_stride = N): how close does a 1/N-frame
step come to 1/N of the whole-frame step, for the launcher's tail settings (development probe; not part of bench.py)."""
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "gpu-heightmap-raytracer_b200"))
sys.path.insert(0, str(REPO))

import torch  # noqa: E402

import bench  # noqa: E402
import hmrt  # noqa: E402


def main():
    torch.cuda.set_device(0)
    ctx = hmrt.Context(0)
    res, idx, total = hmrt.pyramid_layout(bench.COARSE, bench.LEVELS)
    pyr = torch.zeros(total, dtype=torch.float32, device="cuda")
    R0 = bench.R0
    xs = torch.arange(R0, device="cuda", dtype=torch.float32)
    x, z = xs[None, :], xs[:, None]
    fin = pyr[idx[0]:].view(R0, R0)
    fin.copy_(420 + 260 * torch.sin(x * 0.00121) * torch.cos(z * 0.00097) + 110 * torch.sin(x * 0.0047 + z * 0.0039)
              + 45 * torch.sin(x * 0.019) * torch.sin(z * 0.023) + 12 * torch.sin(x * 0.11 + z * 0.07))
    g = torch.Generator(device="cuda").manual_seed(1234)
    fin.add_(torch.rand((R0, R0), device="cuda", generator=g) * 3.0).clamp_(min=0)
    ctx.build_mips(pyr, bench.COARSE, bench.LEVELS)
    mh = float(fin.max())
    ctx.set_heightmap(pyr, None, bench.COARSE, bench.LEVELS, mh)
    K = 20
    cams = [hmrt.context._cam_array(bench.make_cameras(hmrt, s, mh)) for s in range(K)]
    out = {}
    for stride in (1, 8):
        opts = hmrt.trace_opts(mh, tile_first=0, tile_stride=stride)
        rows = hmrt.rows_local(bench.H, 0, stride)
        fbs = [torch.empty((bench.POSES, rows, bench.W, 3), dtype=torch.uint8, device="cuda") for _ in range(2)]
        run = bench.StepRunner(torch, None, ctx, 1, fbs)
        for tailed, per in ((-1, 0), (0, 0), (1, 1), (1, 2), (1, 4), (1, 16)):
            ctx.lib.hmrt_debug_trace_knob(0, tailed)
            ctx.lib.hmrt_debug_trace_knob(1, per)
            for c in cams[:3]:
                run.issue(c, opts)
            reps = 6 if stride == 1 else 48
            ms, _, _ = run.timed(cams, opts, reps)
            out[f"stride{stride}_tailed{tailed}_per{per}"] = ms / (K * reps)
        del fbs
    ctx.lib.hmrt_debug_trace_knob(0, -1)
    ctx.lib.hmrt_debug_trace_knob(1, 0)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
