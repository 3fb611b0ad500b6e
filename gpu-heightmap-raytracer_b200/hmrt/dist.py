"""Multi-GPU plumbing: one process per GPU over torch.distributed (NCCL on GPUs, gloo in CPU tests).

How the path shards:
  ray traversal   pixels are independent -> row tiles of HMRT_ROW_TILE rows, interleaved across
                  ranks (tile t belongs to rank t % world); the pyramid is replicated by ONE
                  broadcast; no collective on the per-frame data path.
  rasterisation   points sharded by contiguous file range; every rank scatters into a private
                  full-size finest level; one max all-reduce combines them (max is associative,
                  commutative and idempotent, so the result is bit-identical to one GPU); each
                  rank then builds the mip levels locally.
Heights are non-negative floats, so their int32 bit patterns order the same way: the all-reduce
runs on the int32 view (exact, NaN-free) exactly like the scatter kernel's atomic max.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from ._abi import HMRT_ROW_TILE


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, rank: int, world_size: int):
    """Contiguous [lo, hi) of n points for this rank (file order is kept inside a shard)."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def tiles_for_rank(rank: int, world_size: int):
    """(tile_first, tile_stride) for hmrt_trace_opts: interleaved row tiles."""
    return rank, world_size


def allreduce_max_heights(finest: torch.Tensor, group=None) -> torch.Tensor:
    """In-place max all-reduce of a finest level (float32 >= +0) on its int32 view."""
    if finest.dtype != torch.float32:
        raise TypeError("finest level must be float32")
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(finest.view(torch.int32), op=dist.ReduceOp.MAX, group=group)
    return finest


def allreduce_max_keys(keys: torch.Tensor, group=None) -> torch.Tensor:
    """Colour keys ((file index + 1) << 24 | rgb, int64 >= 0): max = last writer in file order."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(keys, op=dist.ReduceOp.MAX, group=group)
    return keys


def broadcast_pyramid(pyramid: torch.Tensor, src: int = 0, group=None) -> torch.Tensor:
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(pyramid, src=src, group=group)
    return pyramid


def local_tiles(H: int, rank: int, world_size: int):
    n_tiles = (H + HMRT_ROW_TILE - 1) // HMRT_ROW_TILE
    return list(range(rank, n_tiles, world_size))


def assemble_frame(parts, H: int, W: int) -> torch.Tensor:
    """Full [H, W, 3] frame from per-rank compact row blocks (parts[r] = rank r's [rows_r, W, 3])."""
    world_size = len(parts)
    out = torch.empty((H, W, 3), dtype=parts[0].dtype, device=parts[0].device)
    for r, p in enumerate(parts):
        row = 0
        for t in local_tiles(H, r, world_size):
            r0 = t * HMRT_ROW_TILE
            n = min(HMRT_ROW_TILE, H - r0)
            out[r0:r0 + n] = p[row:row + n]
            row += n
    return out


def gather_frame(local_rows: torch.Tensor, H: int, W: int, group=None) -> torch.Tensor:
    """All-gather the row blocks of one frame and interleave them (every rank gets the frame)."""
    rank, world_size = world()
    if world_size == 1:
        return assemble_frame([local_rows], H, W)
    n_tiles = (H + HMRT_ROW_TILE - 1) // HMRT_ROW_TILE
    max_rows = ((n_tiles + world_size - 1) // world_size) * HMRT_ROW_TILE
    padded = torch.zeros((max_rows, W, 3), dtype=local_rows.dtype, device=local_rows.device)
    padded[: local_rows.shape[0]] = local_rows
    parts = [torch.empty_like(padded) for _ in range(world_size)]
    dist.all_gather(parts, padded, group=group)
    return assemble_frame(parts, H, W)
