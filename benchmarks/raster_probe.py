#!/usr/bin/env python
"""Per-kernel timing of the tile-binned rasterisation on one GPU (development probe; not part of bench.py):
bin / apply / gather+mips of the hmrt_rx_* path with world = 1, the direct-atomics path, and hmrt_scatter_las in auto mode."""
import argparse
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "gpu-heightmap-raytracer_b200"))
sys.path.insert(0, str(REPO / "benchmarks"))

import torch  # noqa: E402

import hmrt  # noqa: E402
from hmrt import dist as hd  # noqa: E402
from hmrt import las  # noqa: E402
import raster_pipeline as rpl  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=500_000_000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--modes", default="peer,single,direct")
    ap.add_argument("--variants", default="", help="records per thread and step of the bin pass to time (development knob): 8,4,2 (peer mode only)")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    ctx = hmrt.Context(0)
    res, idx, total = hmrt.pyramid_layout(rpl.COARSE, rpl.LEVELS)
    rec = rpl.make_records(torch, 0, args.points)
    hdr = las.LasHeader(rpl.FMT, rpl.REC_LEN, args.points, (0.01, 0.01, 0.01), (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), (rpl.R0 * 2.0, rpl.R0 * 2.0, 700.0))
    xf = hdr.transform()
    pyr = torch.empty(total, dtype=torch.float32, device="cuda")
    out = {"points": args.points}
    hashes = {}
    for mode in args.modes.split(","):
        force = {"peer": "peer", "single": None, "direct": None}[mode]
        ctx.set_scatter_mode(1 if mode == "direct" else 0)
        rp = hd.RasterPipeline(ctx, rpl.COARSE, rpl.LEVELS, single=True, force_mode=force)
        best = None
        for _ in range(args.reps):
            t = rp.run(rec, args.points, rpl.REC_LEN, rpl.FMT, xf, pyr, timed=True)
            if best is None or sum(t.values()) < sum(best.values()):
                best = t
        out[mode] = {**best, "total_ms": sum(best.values())}
        hashes[mode] = rpl.finest_hash(torch, pyr[idx[0]:])
        rp.close()
    ctx.set_scatter_mode(0)
    for v in [v for v in args.variants.split(",") if v]:
        assert ctx.lib.hmrt_debug_raster_knob(1, int(v)) == 0
        rp = hd.RasterPipeline(ctx, rpl.COARSE, rpl.LEVELS, single=True, force_mode="peer")
        best = None
        for _ in range(args.reps):
            t = rp.run(rec, args.points, rpl.REC_LEN, rpl.FMT, xf, pyr, timed=True)
            if best is None or sum(t.values()) < sum(best.values()):
                best = t
        out["variant_" + v] = {**best, "total_ms": sum(best.values())}
        hashes["variant_" + v] = rpl.finest_hash(torch, pyr[idx[0]:])
        rp.close()
    ctx.lib.hmrt_debug_raster_knob(1, 0)
    out["hashes_equal"] = len(set(hashes.values())) == 1
    print(json.dumps(out))


if __name__ == "__main__":
    main()
