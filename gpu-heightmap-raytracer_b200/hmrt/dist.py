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

import ctypes as C

from . import _abi
from ._abi import HMRT_ROW_TILE, check


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


class SharedFrames:
    """[n_frames, H, W, 3] uint8 frames that live on rank `root` and are mapped by every other rank (cudaIpc peer memory).
    A row-tile-sharded render with hmrt_trace_opts.full_frame_output = 1 and `out = shared.tensor` has every rank store
    its tiles straight into the root's frames over NVLink, inside the traversal kernel: the frame assembly the reference
    does by rendering one whole frame per call (main.cpp:675-703) costs no extra pass and no collective.  Collective
    constructor (all ranks call it); after a device barrier the root holds complete frames."""

    def __init__(self, ctx, n_frames: int, H: int, W: int, root: int = 0, group=None):
        rank, world_size = world()
        self.root, self.shape = root, (n_frames, H, W, 3)
        nbytes = n_frames * H * W * 3
        handle = bytes(64)
        self.buf, self.tensor = None, None
        why = None
        if rank == root:
            try:
                self.buf, handle = ctx.ipc_alloc(nbytes)
            except _abi.HmrtError as e:
                why = str(e)
        ok = all_ranks_agree(why is None, group)
        handles = exchange_handles(handle, group)
        if ok and rank != root:
            try:
                self.buf = ctx.ipc_open(handles[root], nbytes)
            except _abi.HmrtError as e:
                why = str(e)
        ok = all_ranks_agree(ok and why is None, group)
        if not ok:  # the same outcome on every rank: nobody is left waiting in a collective
            if self.buf is not None:
                self.buf.close()
                self.buf = None
            raise RuntimeError(f"shared frames unavailable (cudaIpc): {why or 'refused on another rank'}")
        self.tensor = self.buf.tensor(self.shape)

    def close(self, group=None):
        rank, world_size = world()
        if self.buf is None:
            return
        self.tensor = None
        if world_size > 1:
            dist.barrier(group)          # nobody still writes into the root's memory
        if rank != self.root:
            self.buf.close()
        if world_size > 1:
            dist.barrier(group)          # every mapping is gone before the owner frees
        if rank == self.root:
            self.buf.close()
        self.buf = None


# ---------------------------------------------------------------------------------------------------------------------
# rasterisation pipelines (BASELINE config 4)

def band_rows(res0: int, world_size: int):
    """rows[r] .. rows[r + 1] = the finest-level rows rank r owns in the peer-memory exchange: whole rows of grid tiles
    (ceil(res0 / 2048) tiles per axis, at least 8, at most 16; 16 when the world size does not divide 8 --
    binned_tile_shift, csrc/rasterx.cu), dealt out contiguously and as evenly as possible -- the host-side mirror of
    hmrt_rx_bands (tests/test_dist_gloo.py compares the two)."""
    per_axis = min(16, max(8, (res0 + 2047) // 2048))
    if 8 % world_size != 0:
        per_axis = 16
    shift = 0
    while ((res0 + (1 << shift) - 1) >> shift) > per_axis:
        shift += 1
    tile = 1 << shift
    tiles = (res0 + tile - 1) >> shift
    base, rem = divmod(tiles, world_size)
    rows, t = [0], 0
    for r in range(world_size):
        t += base + (1 if r < rem else 0)
        rows.append(min(res0, t * tile))
    return rows


def exchange_handles(handle: bytes, group=None):
    """All-gather one fixed-size opaque blob per rank (the 64-byte cudaIpc handle of a rank's exchange region)."""
    rank, world_size = world()
    if world_size == 1:
        return [handle]
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.frombuffer(bytearray(handle), dtype=torch.uint8).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world_size)]
    dist.all_gather(parts, mine, group=group)
    return [bytes(p.cpu().numpy().tobytes()) for p in parts]


def all_ranks_agree(ok: bool, group=None) -> bool:
    """True when `ok` holds on EVERY rank (the exchange path is taken by all ranks or by none)."""
    rank, world_size = world()
    if world_size == 1:
        return bool(ok)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(int(t.item()))


class RasterPipeline:
    """clear -> scatter -> exchange -> mip build for one point cloud sharded by contiguous range over the ranks.

    mode "single"     one GPU: hmrt_clear_section, hmrt_scatter_las, hmrt_build_mips.
    mode "peer"       N GPUs, the owner-computes exchange over peer memory (hmrt_rx_*, csrc/rasterx.cu).
    mode "allreduce"  N GPUs, the north star's literal form: private full-size grids + NCCL max all-reduce of the finest
                      level + local mip build.  Taken when the peer path is unavailable on ANY rank (grid shape, cudaIpc
                      refused) or when it reported an overflow; which one ran is in `self.mode` / exchange_description().
    """

    def __init__(self, ctx, coarse_res: int, levels: int, single: bool = False, max_points_per_rank: int | None = None, force_mode: str | None = None):
        self.ctx, self.coarse, self.levels = ctx, coarse_res, levels
        self.rank, self.world = (0, 1) if single else world()
        self.res0 = coarse_res << (levels - 1)
        self.rx = None
        self.peer_failure = None
        self.mode = "single" if self.world == 1 else "allreduce"
        self._max_points = max_points_per_rank
        self._force = force_mode
        if self.world == 1 and force_mode == "peer":
            self.mode = "peer"
        self.PHASES = ("clear", "scatter", "mips")

    # -- peer path set-up (lazy: needs the per-rank point budget) ----------------------------------------------------
    def _setup_peer(self, n_local: int):
        lib = self.ctx.lib
        want = self._max_points
        if want is None:
            t = torch.tensor([n_local], dtype=torch.int64, device="cuda")
            if self.world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            want = int(t.item())
        h = C.c_void_p()
        rc = lib.hmrt_rx_create(self.ctx._h, self.coarse, self.levels, self.rank, self.world, want, C.byref(h))
        ok = rc == 0
        handle = bytes(64)
        if ok:
            buf = (C.c_uint8 * 64)()
            rc = lib.hmrt_rx_export(h, buf)
            ok = rc == 0
            handle = bytes(buf)
        if not ok:
            self.peer_failure = f"hmrt_rx_create/export: {_abi.error_string(rc)}"
        handles = exchange_handles(handle)
        if all_ranks_agree(ok):
            blob = b"".join(handles)
            rc = lib.hmrt_rx_connect(h, blob)
            ok = rc == 0
            if not ok:
                self.peer_failure = f"hmrt_rx_connect: {_abi.error_string(rc)}"
        else:
            ok = False
        ok = all_ranks_agree(ok)
        if ok:
            self.rx, self._rx_cap = h, want
            self.mode = "peer"
        else:
            if h:
                lib.hmrt_rx_destroy(h)
            self.mode = "allreduce" if self.world > 1 else "single"
            if self._force == "peer":
                raise RuntimeError(f"peer-memory exchange unavailable: {self.peer_failure}")
        return ok

    def close(self):
        if self.rx is not None:
            self.ctx.lib.hmrt_rx_destroy(self.rx)
            self.rx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def exchange_description(self):
        if self.mode == "peer":
            return ("owner-computes over peer memory (cudaIpc / NVLink P2P loads): each rank pulls the (cell, height) pairs of the tile rows it "
                    "owns from every rank's buckets, then one kernel all-gathers the finished bands and builds the mip levels; no NCCL on the data path")
        if self.mode == "allreduce":
            why = f" (peer path unavailable: {self.peer_failure})" if self.peer_failure else ""
            return "NCCL max all-reduce of the dense finest level on its int32 view + local mip build" + why
        return "none (one GPU)"

    # -- runs ------------------------------------------------------------------------------------------------------------
    def scatter_only(self, rec, n, rec_len, fmt, xf, pyr, first_index=0):
        self.ctx.scatter_las(rec, n, rec_len, fmt, xf, pyr, self.coarse, self.levels, first_index=first_index)

    def run(self, rec, n, rec_len, fmt, xf, pyr, first_index=0, timed=False):
        """One rasterisation of this rank's `n` records into `pyr` (every rank ends with the whole pyramid).
        Returns {phase: ms} when timed (synchronises), else None."""
        want_peer = (self.world > 1 and self._force != "allreduce") or self._force == "peer"
        if want_peer and self.rx is None and self.peer_failure is None:
            self._setup_peer(n)
        ev = []

        def mark():
            if timed:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                ev.append(e)

        lib, ctx = self.ctx.lib, self.ctx
        if self.mode == "peer":
            self.PHASES = ("clear", "bin", "exchange_apply", "gather_mips")
            ctx._bind_stream()
            mark()
            check(lib.hmrt_rx_begin(self.rx), "hmrt_rx_begin")
            mark()
            check(lib.hmrt_rx_bin(self.rx, C.c_void_p(rec.data_ptr()) if n > 0 else None, n, rec_len, fmt, C.byref(xf)), "hmrt_rx_bin")
            mark()
            check(lib.hmrt_rx_barrier(self.rx), "hmrt_rx_barrier")
            check(lib.hmrt_rx_apply(self.rx), "hmrt_rx_apply")
            check(lib.hmrt_rx_barrier(self.rx), "hmrt_rx_barrier")
            mark()
            check(lib.hmrt_rx_gather_mips(self.rx, C.c_void_p(pyr.data_ptr())), "hmrt_rx_gather_mips")
            mark()
            ov, er = C.c_uint32(), C.c_uint32()
            check(lib.hmrt_rx_status(self.rx, C.byref(ov), C.byref(er)), "hmrt_rx_status")
            bad = ov.value != 0 or er.value != 0
            if not all_ranks_agree(not bad):
                # a slice overflowed somewhere (heavily skewed cloud) or a barrier timed out: redo with the dense exchange
                self.peer_failure = f"overflow {ov.value}, barrier error {er.value} on rank {self.rank} (or on another rank)"
                self.close()
                self.mode = "allreduce" if self.world > 1 else "single"
                return self.run(rec, n, rec_len, fmt, xf, pyr, first_index, timed)
        else:
            self.PHASES = ("clear", "scatter", "exchange", "mips") if self.world > 1 else ("clear", "scatter", "mips")
            res, idx, total = _layout(self.coarse, self.levels)
            mark()
            ctx.clear_section(pyr, self.coarse, self.levels)
            mark()
            ctx.scatter_las(rec, n, rec_len, fmt, xf, pyr, self.coarse, self.levels, first_index=first_index)
            mark()
            if self.world > 1:
                allreduce_max_heights(pyr[idx[0]:])
                mark()
            ctx.build_mips(pyr, self.coarse, self.levels)
            mark()
        if not timed:
            return None
        torch.cuda.synchronize()
        return {k: ev[i].elapsed_time(ev[i + 1]) for i, k in enumerate(self.PHASES)}


def _layout(coarse_res, levels):
    from .context import pyramid_layout

    return pyramid_layout(coarse_res, levels)
