#!/usr/bin/env python
"""hmrt_trace_host_begin / _wait on the bench workload (16 4K frames per step over 16384^2), two calls in flight, for the launch
granularities of the per-group schedule (development knob: frame groups per middle launch).  One GPU; development probe."""
import argparse
import ctypes as C
import json
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "gpu-heightmap-raytracer_b200"))

import torch  # noqa: E402

import bench  # noqa: E402
import hmrt  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mid", default="1,2,4,8")
    ap.add_argument("--steps", type=int, default=40)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    ctx = hmrt.Context(0)
    W, H = bench.W, bench.H
    POSES = bench.POSES
    res, idx, total = hmrt.pyramid_layout(bench.COARSE, bench.LEVELS)
    pyr = torch.empty(total, dtype=torch.float32, device="cuda")
    fin = bench.build_terrain_host()
    mh = float(fin.max())
    pyr[idx[0]:].view(bench.R0, bench.R0).copy_(torch.from_numpy(fin))
    ctx.build_mips(pyr, bench.COARSE, bench.LEVELS)
    ctx.set_heightmap(pyr, None, bench.COARSE, bench.LEVELS, mh)
    opts = hmrt.trace_opts(mh)
    cams = [hmrt.context._cam_array(bench.make_cameras(hmrt, s, mh)) for s in range(args.steps)]
    host = [torch.empty((POSES, H, W, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
    ctx.lib.hmrt_debug_host_mid_groups.argtypes = [C.c_void_p, C.c_int]
    out = {}
    for mid in [int(x) for x in args.mid.split(",")]:
        assert ctx.lib.hmrt_debug_host_mid_groups(ctx._h, mid) == 0
        for pipelined in (False, True):
            for rep in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for n, c in enumerate(cams):
                    if pipelined:
                        ctx.trace_host_begin(W, H, c, opts, host[n & 1])
                        if n:
                            ctx.trace_host_wait()
                    else:
                        ctx.trace_host(W, H, c, opts, host[n & 1])
                if pipelined:
                    ctx.trace_host_wait()
                torch.cuda.synchronize()
                ms = 1e3 * (time.perf_counter() - t0) / len(cams)
            out[f"mid{mid}_{'pipelined' if pipelined else 'sync'}_ms_per_step"] = ms
    ctx.lib.hmrt_debug_host_mid_groups(ctx._h, 0)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
