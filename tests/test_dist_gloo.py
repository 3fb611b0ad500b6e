"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes (hmrt/dist.py)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmp):
    sys.path.insert(0, str(REPO / "gpu-heightmap-raytracer_b200"))
    sys.path.insert(0, str(REPO / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hmrt import dist as hd

    import oraclelib as ol
    import rasterlib as rl

    # --- rasterisation: shard by point range, max all-reduce on the int32 view ---------------
    r0, levels = 128, 5
    hdr, rec = rl.synthetic_las(40_000, r0, seed=5)
    lo, hi = hd.shard_range(len(rec), rank, world)
    part, _ = rl.oracle_rasterise(hdr, rec[lo:hi], r0 >> (levels - 1), levels, with_colors=False)
    res, idx, total = ol.pyramid_layout(r0 >> (levels - 1), levels)
    finest = torch.from_numpy(part[idx[0]:].copy())
    hd.allreduce_max_heights(finest)
    full, _ = rl.oracle_rasterise(hdr, rec, r0 >> (levels - 1), levels, with_colors=False)
    assert (finest.numpy() == full[idx[0]:]).all(), "sharded scatter + max all-reduce != single pass"

    # --- colour keys: max over ranks keeps the last writer in file order ---------------------
    keys = torch.zeros(16, dtype=torch.int64)
    keys[3] = ((lo + 5 + 1) << 24) | (rank + 1)
    hd.allreduce_max_keys(keys)
    assert int(keys[3]) & 0xFFFFFF == world  # the highest rank owns the later file range

    # --- broadcast + row-tile assembly ---------------------------------------------------------
    pyr = torch.from_numpy(full.copy()) if rank == 0 else torch.zeros(total)
    hd.broadcast_pyramid(pyr, src=0)
    assert (pyr.numpy() == full).all()
    H, W = 83, 16  # ragged: 11 tiles, last one 3 rows
    frame = torch.arange(H * W * 3, dtype=torch.int64).remainder(251).to(torch.uint8).reshape(H, W, 3)
    rows = torch.cat([frame[t * 8:min(H, t * 8 + 8)] for t in hd.local_tiles(H, rank, world)])
    got = hd.gather_frame(rows, H, W)
    assert (got == frame).all()
    # --- host logic of the peer-memory rasterisation exchange (hmrt.dist.RasterPipeline) ----------
    blobs = hd.exchange_handles(bytes([rank + 1]) * 64)
    assert blobs == [bytes([r + 1]) * 64 for r in range(world)]
    assert hd.all_ranks_agree(True) is True
    assert hd.all_ranks_agree(rank == 0) is False      # one rank cannot take the peer path -> nobody does
    Path(tmp, f"ok{rank}").write_text("ok")
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_band_rows_partition_the_grid():
    """Owned bands of the peer exchange: whole tile rows, contiguous, in rank order, covering every row once."""
    sys.path.insert(0, str(REPO / "gpu-heightmap-raytracer_b200"))
    from hmrt import dist as hd

    for r0 in (2048, 4096, 16384, 32768, 2176):
        for w in (1, 2, 3, 4, 8, 16):
            rows = hd.band_rows(r0, w)
            assert rows[0] == 0 and rows[-1] == r0 and len(rows) == w + 1
            assert all(a <= b for a, b in zip(rows[:-1], rows[1:]))
            tile = max(b - a for a, b in zip(rows[:-1], rows[1:]))
            assert all(r % 128 == 0 for r in rows[:-1]) or r0 % 128
    assert hd.band_rows(16384, 8) == [2048 * i for i in range(9)]


def test_shard_range_covers_everything():
    sys.path.insert(0, str(REPO / "gpu-heightmap-raytracer_b200"))
    from hmrt import dist as hd

    for n in (0, 1, 7, 500_000_000):
        for w in (1, 2, 4, 8):
            spans = [hd.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
