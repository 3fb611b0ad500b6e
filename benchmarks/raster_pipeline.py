"""Rasterisation sub-record of bench.py (BASELINE.json configs[3]): N synthetic LAS points -> 16384^2 grid at the bench's
world size.  Points are a pure function of their GLOBAL index (integer hash), so every sharding rasterises the same cloud and
the finest level must hash to the same 64-bit value at every N; at N > 1 rank 0 also rasterises the whole cloud alone and the
two hashes are compared in the same run."""
from __future__ import annotations

import time

R0, LEVELS = 16384, 8
COARSE = R0 >> (LEVELS - 1)
FMT, REC_LEN = 0, 20
_M64 = (1 << 64) - 1


def _s64(v):  # python int -> the int64 with the same 64 bits
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


def _mix(torch, v):
    """splitmix64 finaliser on int64 tensors (wrap-around multiplies, logical shifts)."""
    v = (v ^ ((v >> 30) & ((1 << 34) - 1))) * _s64(0xBF58476D1CE4E5B9)
    v = (v ^ ((v >> 27) & ((1 << 37) - 1))) * _s64(0x94D049BB133111EB)
    return v ^ ((v >> 31) & ((1 << 33) - 1))


def make_records(torch, lo, hi, out=None, chunk=1 << 26):
    """LAS format-0 records (20 B) of the points with global indices [lo, hi): uniformly scattered x, y over the grid, terrain-like z."""
    n = hi - lo
    rec = out if out is not None else torch.zeros((n, REC_LEN), dtype=torch.uint8, device="cuda")
    ext_raw = int(R0 * 2.0 / 0.01)
    for a in range(lo, hi, chunk):
        b = min(hi, a + chunk)
        i = torch.arange(a, b, device="cuda", dtype=torch.int64)
        h1 = _mix(torch, i * 2 + 1)
        h2 = _mix(torch, h1 + _s64(0x9E3779B97F4A7C15))
        h3 = _mix(torch, h2 + _s64(0x9E3779B97F4A7C15))
        X = ((h1 >> 1) & ((1 << 62) - 1)) % ext_raw
        Y = ((h2 >> 1) & ((1 << 62) - 1)) % ext_raw
        Z = (40000 + 26000 * torch.sin(X.float() * 6e-6) * torch.cos(Y.float() * 5e-6)).to(torch.int64) + ((h3 >> 1) & ((1 << 62) - 1)) % 300
        xyz = torch.stack([X, Y, Z], dim=1).to(torch.int32)
        rec[a - lo:b - lo, 0:12] = xyz.view(torch.uint8).view(b - a, 12)
        del i, h1, h2, h3, X, Y, Z, xyz
    return rec


def finest_hash(torch, finest):
    """64-bit position-weighted hash of the finest level's bit patterns."""
    total = torch.zeros((), dtype=torch.int64, device=finest.device)
    bits = finest.view(torch.int32)
    chunk = 1 << 26
    for a in range(0, bits.numel(), chunk):
        b = min(bits.numel(), a + chunk)
        w = torch.arange(a, b, device=finest.device, dtype=torch.int64) * 2 + 1
        total += (bits[a:b].to(torch.int64) * w).sum()
    return int(total.item()) & _M64


def measure(torch, dist, hmrt, ctx, rank, world, points, hbm_peak, reps=3):
    from hmrt import dist as hd
    from hmrt import las

    res, idx, total = hmrt.pyramid_layout(COARSE, LEVELS)
    lo, hi = hd.shard_range(points, rank, world)
    n = hi - lo
    hdr = las.LasHeader(FMT, REC_LEN, points, (0.01, 0.01, 0.01), (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), (R0 * 2.0, R0 * 2.0, 700.0))
    xf = hdr.transform()
    pyr = torch.empty(total, dtype=torch.float32, device="cuda")
    finest = pyr[idx[0]:]
    t0 = time.perf_counter()
    rec = make_records(torch, lo, hi)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0

    rp = hd.RasterPipeline(ctx, COARSE, LEVELS)
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = rp.run(rec, n, REC_LEN, FMT, xf, pyr, first_index=lo, timed=True)
        torch.cuda.synchronize()
        tt = torch.tensor([t[k] for k in rp.PHASES], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tt = [float(v) for v in tt]
        if best is None or sum(tt) < sum(best):
            best = tt
    phases = dict(zip(rp.PHASES, best))
    end_to_end_ms = sum(best)
    my_hash = finest_hash(torch, finest)
    top_max = float(pyr[:COARSE * COARSE].max().item())

    # every rank must hold the same finest level
    same_everywhere = True
    single_hash = my_hash
    if world > 1:
        hs = torch.tensor([_s64(my_hash)], dtype=torch.int64, device="cuda")
        all_h = [torch.empty_like(hs) for _ in range(world)]
        dist.all_gather(all_h, hs)
        same_everywhere = all(int(h.item()) == int(hs.item()) for h in all_h)
        # rank 0 alone rasterises the WHOLE cloud (shard by shard, to bound memory) with the single-GPU path
        if rank == 0:
            rp1 = hd.RasterPipeline(ctx, COARSE, LEVELS, single=True)
            ref = torch.empty(total, dtype=torch.float32, device="cuda")
            ctx.clear_section(ref, COARSE, LEVELS)
            for r in range(world):
                a, b = hd.shard_range(points, r, world)
                make_records(torch, a, b, out=rec[: b - a])
                rp1.scatter_only(rec, b - a, REC_LEN, FMT, xf, ref, first_index=a)
            torch.cuda.synchronize()
            single_hash = finest_hash(torch, ref[idx[0]:])
            del ref
        sh = torch.tensor([_s64(single_hash)], dtype=torch.int64, device="cuda")
        dist.broadcast(sh, 0)
        single_hash = int(sh.item()) & _M64
    del rec, pyr
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    algo = points * (REC_LEN + 4) / world  # per GPU: read the record, one 4-byte atomic max
    scatter_ms = phases.get("scatter", 0.0) + phases.get("bin", 0.0) + phases.get("exchange_apply", 0.0)
    return {
        "workload": f"{points} LAS format-{FMT} points ({REC_LEN} B records, uniformly scattered: no spatial order) -> {R0}^2 grid, "
                    f"sharded by contiguous point range over {world} GPU(s)",
        "phases_ms": phases, "end_to_end_ms": end_to_end_ms, "Mpoints_per_s": points / end_to_end_ms / 1e3,
        "scatter_roofline_frac": algo / (scatter_ms * 1e-3) / 1e9 / hbm_peak if scatter_ms > 0 else None,
        "exchange": rp.exchange_description(),
        "finest_hash": f"{my_hash:016x}", "single_gpu_hash": f"{single_hash:016x}",
        "hash_equals_single_gpu": bool(same_everywhere and my_hash == single_hash),
        "top_level_max": top_max, "point_generation_s": gen_s,
    }
