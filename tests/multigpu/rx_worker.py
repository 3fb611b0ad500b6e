"""torchrun worker of tests/test_gpu_rx.py::test_rx_multi_gpu_exchange: one rank per GPU over NCCL.
Every rank rasterises its contiguous share of one seeded cloud; the peer-memory exchange (hmrt_rx_*), the NCCL max
all-reduce path and a single-GPU rasterisation of the whole cloud must give the same pyramid bit for bit, on every rank."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

REPO = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(REPO / "gpu-heightmap-raytracer_b200"))
sys.path.insert(0, str(REPO / "tests"))

import hmrt  # noqa: E402
from hmrt import dist as hd  # noqa: E402

import oraclelib as ol  # noqa: E402
import rasterlib as rl  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = hmrt.Context(local)
    for r0, levels, n, fmt in [(2048, 8, 900_001, 0), (4096, 7, 1_500_000, 2)]:
        coarse = r0 >> (levels - 1)
        hdr, rec = rl.synthetic_las(n, r0, point_format=fmt, seed=r0)
        res, idx, total = ol.pyramid_layout(coarse, levels)
        xf = hdr.transform()
        lo, hi = hd.shard_range(n, rank, world)
        d = torch.from_numpy(np.ascontiguousarray(rec[lo:hi])).cuda()
        results = {}
        for force in ("peer", "allreduce"):
            rp = hd.RasterPipeline(ctx, coarse, levels, force_mode=force)
            pyr = torch.empty(total, dtype=torch.float32, device="cuda").fill_(-1.0)
            for rep in range(3):  # repeated runs reuse the exchange region: the barriers must keep the ranks apart
                t = rp.run(d, hi - lo, rec.shape[1], fmt, xf, pyr, first_index=lo, timed=True)
            assert rp.mode == force, (rp.mode, rp.peer_failure)
            results[force] = pyr.cpu().numpy()
            rp.close()
        # the whole cloud on one GPU, and the CPU oracle on rank 0
        rp1 = hd.RasterPipeline(ctx, coarse, levels, single=True)
        whole = torch.empty(total, dtype=torch.float32, device="cuda")
        rp1.run(torch.from_numpy(rec).cuda(), n, rec.shape[1], fmt, xf, whole)
        torch.cuda.synchronize()
        whole = whole.cpu().numpy()
        assert (results["peer"].view(np.uint32) == whole.view(np.uint32)).all(), f"rank {rank}: peer exchange != one GPU ({r0})"
        assert (results["allreduce"].view(np.uint32) == whole.view(np.uint32)).all(), f"rank {rank}: all-reduce path != one GPU ({r0})"
        if rank == 0:
            want, _ = rl.oracle_rasterise(hdr, rec, coarse, levels, with_colors=False)
            assert (whole.view(np.uint32) == want.view(np.uint32)).all(), "one GPU != oracle"
    # ---- row-tile-sharded render assembled on rank 0 through a peer mapping (hmrt.dist.SharedFrames) -----------------
    sc = ol.scene("r1024_l8", seed=2)
    pyr = torch.from_numpy(sc["pyramid"]).cuda()
    ctx.set_heightmap(pyr, None, sc["coarse"], sc["levels"], sc["max_height"])
    W, H = 640, 363  # ragged last tile
    cams = ol.cameras_for(sc, 3)
    shared = hd.SharedFrames(ctx, len(cams), H, W, root=0)
    if rank == 0:
        shared.tensor.fill_(7)
    torch.cuda.synchronize()
    dist.barrier()
    tf, ts = hd.tiles_for_rank(rank, world)
    ctx.trace(W, H, cams, hmrt.trace_opts(sc["max_height"], shadows=True, tile_first=tf, tile_stride=ts, full_frame_output=True), out=shared.tensor)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        whole, _ = ctx.trace(W, H, cams, hmrt.trace_opts(sc["max_height"], shadows=True))
        torch.cuda.synchronize()
        assert torch.equal(whole, shared.tensor), "frames assembled over NVLink differ from a single-GPU render"
    shared.close()
    dist.barrier()
    if rank == 0:
        print("rx_worker: ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
