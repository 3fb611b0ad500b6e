#!/usr/bin/env python
"""bench.py -- Mrays/s of the heightfield ray traversal (BASELINE.json metric) on N B200s.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on; it fits one GPU):
16384^2 synthetic heightmap (8-level max-mipmap pyramid, 1.43 GB), 3840x2160 primary rays +
height-ramp shading, POSES camera poses per step.  A "step" is one pass of the traversal over one
batch of POSES frames; every step uses a different pose batch, and the pyramid is far larger
than L2, so successive steps do not replay a cached working set.

Both arms build the terrain with the SAME host code (torch CPU ops, fixed seed): identical bytes, checked by a
checksum both arms print.  The GPU arm uploads it and builds the mip levels with the product's kernel.

N > 1 (torchrun, one rank per GPU): rank 0 builds the pyramid and broadcasts it over NCCL/NVLink;
each frame is cut into 8-row tiles interleaved across ranks (no collective on the per-frame data
path).  Total work per step is fixed -> "scaling": "strong".

Timing: W warm-up steps, then the K steps are timed; when K steps last less than --min-seconds (1 s) the K-step loop
is repeated R times inside ONE timed region (same K pose batches) and ms_per_step = region / (K * R), so that the
region is long enough for clock sampling at every N.  Consecutive steps are independent frames batches: they are
issued alternately on two CUDA streams with their own framebuffers (a double-buffered renderer), so the drain of
one step's persistent kernel runs under the head of the next.

`--impl reference` times the reference's own CPU code (oracle/_ref: CudaKernel.cu host-compiled;
the oracle port if that library is absent) on the host cores over a bounded row sample of the
same workload.  Nothing under oracle/ is used by the measured GPU arm except as the separately
reported `cpu_baseline` (whose rows double as the `parity` check of the GPU frames).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO / "gpu-heightmap-raytracer_b200"))

R0, LEVELS = 16384, 8
COARSE = R0 >> (LEVELS - 1)
W, H = 3840, 2160
POSES = 16
FRAME_DIM = (32.0, 18.0, 20.0)  # main.cpp:57
WORKLOAD = f"{R0}^2 heightmap ({LEVELS}-level max-mip pyramid, 1.43 GB), {W}x{H} primary rays + height-ramp shading, {POSES} camera poses per step"
# the SAME dict in both arms (the driver compares them)
CONFIG = {
    "workload": WORKLOAD,
    "l2": "inputs larger than L2 (1.43 GB pyramid); a different pose batch every step",
    "levels": LEVELS,
    "frame_dimension": list(FRAME_DIM),
    "poses_per_step": POSES,
    "pose_family": "altitude 2000-4000 cells above the terrain maximum, pitch -0.10..-0.35, all headings, positions spread over the map",
    "terrain": "analytic hills + seeded noise generated on the HOST by the same code in both arms (identical bytes; see terrain_checksum)",
}
NCU_SUMMARY = REPO / "profiles" / "ncu_trace_r02.json"  # per-launch counters of the dominant kernel from the committed ncu capture
RASTER_POINTS = 500_000_000  # BASELINE.json configs[3]


# ------------------------------------------------------------------------------------------------
# workload definition (shared by both arms; pure host logic)

def pose_batch(step: int, max_height: float, family: str = "high"):
    """POSES deterministic camera poses for step `step`.
    "high" (the metric's family): altitude 2000-4000 cells, pitch -0.35..-0.1.
    "low"  (descent-dominated views): altitude 200-500 cells above the terrain maximum, pitch -0.3..-3 (down to nadir-like)."""
    poses = []
    for i in range(POSES):
        k = step * POSES + i
        u = (k * 0.6180339887498949) % 1.0
        v = (k * 0.7548776662466927) % 1.0
        heading = 2 * math.pi * ((k * 0.5698402909980532) % 1.0)
        a = (k * 0.3819660112501051) % 1.0
        b = (k * 0.2451223337533073) % 1.0
        if family == "high":
            pitch = -0.10 - 0.25 * a
            alt = max_height + 2000.0 + 2000.0 * b
        else:
            pitch = -0.3 * (10.0 ** a)  # -0.3 .. -3.0: from a shallow look down to ~72 degrees below the horizon
            alt = max_height + 200.0 + 300.0 * b
        pos = (R0 * (0.2 + 0.6 * u), alt, R0 * (0.2 + 0.6 * v))
        fwd = (math.cos(heading), pitch, math.sin(heading))
        poses.append((pos, fwd))
    return poses


def make_cameras(hmrt, step, max_height, family="high"):
    return [hmrt.camera(p, f, FRAME_DIM) for p, f in pose_batch(step, max_height, family)]


def build_terrain_host() -> np.ndarray:
    """The finest level [R0, R0] float32 on the HOST -- the one terrain generator of both arms (same torch CPU code, same
    seed, same box => same bytes).  ~5 s."""
    import torch  # CPU tensors only

    xs = torch.arange(R0, dtype=torch.float32)
    x, z = xs[None, :], xs[:, None]
    fin = torch.empty((R0, R0), dtype=torch.float32)
    torch.mul((torch.sin(xs * 0.00121) * 260)[None, :], torch.cos(xs * 0.00097)[:, None], out=fin)
    fin += 420
    tmp = x * 0.0047 + z * 0.0039
    fin += tmp.sin_().mul_(110)
    fin += (torch.sin(xs * 0.019) * 45)[None, :] * torch.sin(xs * 0.023)[:, None]
    torch.add(x * 0.11, z * 0.07, out=tmp)
    fin += tmp.sin_().mul_(12)
    g = torch.Generator().manual_seed(1234)
    tmp.uniform_(0.0, 3.0, generator=g)
    fin += tmp
    fin.clamp_(min=0)
    return fin.numpy()


def terrain_checksum(fin: np.ndarray) -> str:
    w = np.ascontiguousarray(fin).view(np.uint32).ravel()
    return f"{int(w.sum(dtype=np.uint64)):016x}-{int(np.bitwise_xor.reduce(w)):08x}"


# ------------------------------------------------------------------------------------------------
# clocks

CLOCK_QUERY = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
               "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


class ClockSampler:
    def __init__(self, gpu_index: int):
        self.path = tempfile.NamedTemporaryFile(prefix="hmrt_clocks_", suffix=".csv", delete=False).name
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={CLOCK_QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self, t_begin=None, t_end=None):
        """Summarise the samples whose timestamp falls inside [t_begin, t_end] (epoch seconds)."""
        import datetime

        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [t.strip() for t in line.split(",")]
            if len(f) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if t_begin is not None and not (t_begin - 0.05 <= ts <= t_end + 0.05):
                    continue
                sm_v, mx_v = float(f[2]), float(f[3])
            except ValueError:
                continue
            sm.append(sm_v)
            mx.append(mx_v)
            for name, val in zip(names, f[6:10]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own code on the host cores

def _load_cpu_impl():
    """('reference', fn) from oracle/_ref if present, else ('port', fn) from the plain-C oracle."""
    sys.path.insert(0, str(REPO / "tests"))
    import oraclelib as ol

    lib = ol.ref()
    if lib is not None:
        return "reference", lib.hmrt_ref_trace, ol
    return "port", ol.oracle().hmrt_oracle_trace, ol


def cpu_sample_rate(pyramid_host: np.ndarray, max_height: float, step: int, budget_s: float, n_threads: int, family="high",
                    compare_with=None):
    """Trace 8-row bands, spread over the frames of pose batch `step`, until `budget_s` of tracing time is spent.
    `compare_with` (optional, [POSES, H, W, 3] uint8 host array of the GPU's frames of the same pose batch): every traced
    band is compared with the same rows of it (outside the timed part).
    Returns (Mrays/s, rays traced, seconds, kind, parity dict or None)."""
    kind, fn, ol = _load_cpu_impl()
    cams = [ol.make_camera(p, f, FRAME_DIM) for p, f in pose_batch(step, max_height, family)]
    opts = ol.make_opts(max_height)
    rgb = np.zeros((H, W, 3), np.uint8)
    rows_per_call = max(8, 2 * n_threads)
    # bands interleaved over the frame height (sky and ground rows alike), cycling through the poses;
    # several passes with shifted offsets so the budget, not the band list, ends the sample
    bands = [(b * (H // 8) + off + shift) for shift in range(0, rows_per_call, 8) for off in range(0, H // 8, rows_per_call)
             for b in range(8)]
    bands = bands * len(cams)
    rays, spent = 0, 0.0
    par = {"rows": 0, "pixels": 0, "pixels_differ": 0, "frames_touched": 0, "checked_against": f"oracle/_ref ({kind})"} if compare_with is not None else None
    touched = set()
    for j, r0 in enumerate(bands):
        ci = (j // 8) % len(cams)
        r1 = min(H, r0 + rows_per_call)
        t0 = time.perf_counter()
        rc = fn(pyramid_host.ctypes.data, None, COARSE, LEVELS, W, H, C.byref(cams[ci]), C.byref(opts), n_threads, r0, r1,
                rgb.ctypes.data, None)
        spent += time.perf_counter() - t0
        if rc != 0:
            raise RuntimeError(f"cpu arm failed: {rc}")
        rays += (r1 - r0) * W
        if par is not None:
            par["rows"] += r1 - r0
            par["pixels"] += (r1 - r0) * W
            par["pixels_differ"] += int((rgb[r0:r1] != compare_with[ci, r0:r1]).any(axis=2).sum())
            touched.add(ci)
        if spent >= budget_s:
            break
    if par is not None:
        par["frames_touched"] = len(touched)
    return rays / spent / 1e6, rays, spent, kind, par


def host_pyramid_from_finest(fin: np.ndarray) -> np.ndarray:
    sys.path.insert(0, str(REPO / "tests"))
    import oraclelib as ol

    return ol.pyramid_from_finest(fin, LEVELS)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n_threads = os.cpu_count() or 1
    fin = build_terrain_host()
    checksum = terrain_checksum(fin)
    pyr = host_pyramid_from_finest(fin)
    mh = float(fin.max())
    per_step_budget = max(0.5, min(4.0, 150.0 / max(1, args.steps)))  # the whole arm stays within a few minutes for any --steps
    for s in range(args.warmup):
        cpu_sample_rate(pyr, mh, s, min(1.0, per_step_budget), n_threads)
    rays_total, t_total, kind = 0, 0.0, "port"
    for s in range(args.steps):
        _, rays, dt, kind, _ = cpu_sample_rate(pyr, mh, args.warmup + s, per_step_budget, n_threads)
        rays_total += rays
        t_total += dt
    value = rays_total / t_total / 1e6
    sample = f"each step: 8-row-band samples of the {POSES} 4K frames of that step's pose batch until {per_step_budget:.1f} s elapse ({rays_total // max(1, args.steps)} rays/step)"
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(1, args.steps), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": CONFIG, "terrain_checksum": checksum,
        "arm": "the reference's own CPU code (oracle/_ref) on all host cores",
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": n_threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm

def run_gpu_arm(args):
    # Libraries (NCCL's version banner, for one) write to stdout; the contract is ONE JSON line there.
    # Everything until the final print goes to stderr instead.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line = _run_gpu_arm(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


class StepRunner:
    """Issues trace steps alternately on two streams with their own framebuffers and times whole regions on the device."""

    def __init__(self, torch, dist, ctx, world, fbs):
        self.torch, self.dist, self.ctx, self.world, self.fbs = torch, dist, ctx, world, fbs
        self.streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        self.issued = 0

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()

    def issue(self, cams, opts, hits=None):
        k = self.issued & 1
        self.issued += 1
        with self.torch.cuda.stream(self.streams[k]):
            self.ctx.trace(W, H, cams, opts, out=self.fbs[k], hits=hits if hits is not None else False)

    def with_buffers(self, fbs):
        """The same streams, other output buffers (e.g. frames shared with rank 0)."""
        other = StepRunner.__new__(StepRunner)
        other.__dict__.update(self.__dict__)
        other.fbs = fbs
        return other

    def timed(self, cam_batches, opts, repeats=1):
        """Device time (ms) of len(cam_batches) * repeats steps, max over ranks, bracketed by barrier + synchronize."""
        torch = self.torch
        self.barrier()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        t_begin = time.time()
        start.record(cur)
        for s in self.streams:
            s.wait_event(start)
        for _ in range(repeats):
            for cams in cam_batches:
                self.issue(cams, opts)
        for s in self.streams:
            cur.wait_stream(s)
        stop.record(cur)
        self.barrier()
        t_end = time.time()
        ms = start.elapsed_time(stop)
        if self.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, t_begin, t_end

    def pick_repeats(self, cam_batches, opts, min_seconds):
        ms, _, _ = self.timed(cam_batches, opts, 1)
        return max(1, int(math.ceil(min_seconds * 1e3 / max(ms, 1e-3))))


def _allsum(torch, dist, world, v):
    if world == 1:
        return int(v)
    t = torch.tensor([int(v)], dtype=torch.int64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def _frames_hash(torch, fb, tile_first, tile_stride):
    """Order-independent 63-bit hash contribution of this rank's rows of one step's frames: sum over bytes of
    byte * odd weight(global position).  Summed over ranks it is the same number for every N iff the frames are."""
    from hmrt import HMRT_ROW_TILE

    n, rows, w, _ = fb.shape
    tiles = torch.arange(tile_first, tile_first + (rows + HMRT_ROW_TILE - 1) // HMRT_ROW_TILE * tile_stride, tile_stride, device=fb.device)
    grow = (tiles[:, None] * HMRT_ROW_TILE + torch.arange(HMRT_ROW_TILE, device=fb.device)[None, :]).reshape(-1)[:rows]  # global row of each local row
    total = torch.zeros((), dtype=torch.int64, device=fb.device)
    col = torch.arange(w * 3, device=fb.device, dtype=torch.int64)
    for f in range(n):
        pos = (f * H + grow.to(torch.int64))[:, None] * (w * 3) + col[None, :]
        total += (fb[f].reshape(rows, w * 3).to(torch.int64) * (pos * 2654435761 % 2147483647 * 2 + 1)).sum()
    return int(total.item()) & ((1 << 63) - 1)


def _tolerance_agreement(torch, hits_exact, rgb_exact, hits_tol, rgb_tol, cam_positions):
    """North-star acceptance bars of the tolerance mode against the exact walk (== the reference, see `parity`) on the same
    frames: hit cell, hit distance from the camera (on pixels that hit the same cell), colour."""
    fe, ft = hits_exact[..., 3], hits_tol[..., 3]
    hit_e, hit_t = (fe & 1) != 0, (ft & 1) != 0
    pe = hits_exact[..., :3].view(torch.float32)
    pt = hits_tol[..., :3].view(torch.float32)
    both = hit_e & hit_t
    same_cell = (torch.floor(pe[..., 0]) == torch.floor(pt[..., 0])) & (torch.floor(pe[..., 2]) == torch.floor(pt[..., 2])) & ((fe & 6) == (ft & 6))
    cell_ok = (both & same_cell) | (~hit_e & ~hit_t)
    n = hit_e.numel()

    def dist_to_camera(p, flags):
        x = torch.where((flags & 2) != 0, float(R0) - p[..., 0], p[..., 0]).double()
        z = torch.where((flags & 4) != 0, float(R0) - p[..., 2], p[..., 2]).double()
        c = cam_positions.double()[:, None, None, :]
        return torch.sqrt((x - c[..., 0]) ** 2 + (p[..., 1].double() - c[..., 1]) ** 2 + (z - c[..., 2]) ** 2)

    de, dt = dist_to_camera(pe, fe), dist_to_camera(pt, ft)
    ok = both & same_cell
    rel = torch.where(ok, (de - dt).abs() / de.clamp(min=1e-9), torch.zeros_like(de))
    col = (rgb_exact.to(torch.int16) - rgb_tol.to(torch.int16)).abs().amax(dim=-1)
    return {"pixels": n, "hit_cell_match_pct": 100.0 * int(cell_ok.sum().item()) / n,
            "hit_miss_flips": int((hit_e != hit_t).sum().item()),
            "max_relative_hit_distance_error_on_matching_cells": float(rel.max().item()),
            "hit_distance_within_1e-4_pct_of_matching_cells": 100.0 * (1.0 - int((rel > 1e-4).sum().item()) / max(1, int(ok.sum().item()))),
            "colour_within_1_of_255_pct": 100.0 * int((col <= 1).sum().item()) / n,
            "pixel_exact_pct": 100.0 * int((col == 0).sum().item()) / n,
            "bars": "north star: hit cell >= 99.9 %, hit distance within 1e-4 relative, colour within 1/255, >= 95 % pixel-exact"}


def _run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    import hmrt
    from hmrt import dist as hd

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the GPU arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world != args.gpus and rank == 0:
        print(f"bench.py: warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    sampler = ClockSampler(local) if rank == 0 else None  # started early (nvidia-smi is slow to start); filtered to the timed window
    ctx = hmrt.Context(local)
    # heightmap: generated on rank 0's host (same code as the reference arm), mip levels by the product's kernel, replicated
    # by one broadcast over NVLink
    res, idx, total = hmrt.pyramid_layout(COARSE, LEVELS)
    pyr = torch.empty(total, dtype=torch.float32, device="cuda")
    mh, checksum, fin_host = 0.0, None, None
    if rank == 0:
        fin_host = build_terrain_host()
        checksum = terrain_checksum(fin_host)
        mh = float(fin_host.max())
        pyr[idx[0]:].view(R0, R0).copy_(torch.from_numpy(fin_host))
        ctx.build_mips(pyr, COARSE, LEVELS)
        torch.cuda.synchronize()
        if world > 1 or args.no_cpu_baseline:
            fin_host = None
    bcast_ms = None
    if world > 1:
        mh_t = torch.tensor([mh], dtype=torch.float32, device="cuda")
        dist.broadcast(mh_t, 0)
        mh = float(mh_t.item())
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        hd.broadcast_pyramid(pyr, 0)
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
    ctx.set_heightmap(pyr, None, COARSE, LEVELS, mh)
    if args.l2_persist:
        lvl, ratio = args.l2_persist.split(":")
        ctx.set_l2_persist(int(lvl), float(ratio))

    tile_first, tile_stride = hd.tiles_for_rank(rank, world)
    opts = hmrt.trace_opts(mh, tile_first=tile_first, tile_stride=tile_stride)
    rows = hmrt.rows_local(H, tile_first, tile_stride)
    fbs = [torch.empty((POSES, rows, W, 3), dtype=torch.uint8, device="cuda") for _ in range(2)]
    K, Wu = args.steps, args.warmup
    cams_by_step = [hmrt.context._cam_array(make_cameras(hmrt, s, mh)) for s in range(Wu + K)]
    timed_batches = cams_by_step[Wu:]
    run = StepRunner(torch, dist, ctx, world, fbs)
    rays_per_step = POSES * W * H

    # ---- value: inputs resident, device-timed --------------------------------------------------
    for s in range(Wu):
        run.issue(cams_by_step[s], opts)
    repeats = run.pick_repeats(timed_batches, opts, args.min_seconds)
    if world > 1:
        r_t = torch.tensor([repeats], dtype=torch.int64, device="cuda")
        dist.all_reduce(r_t, op=dist.ReduceOp.MAX)
        repeats = int(r_t.item())
    launches0 = ctx.launch_count
    ms, t_begin, t_end = run.timed(timed_batches, opts, repeats)
    launches = _allsum(torch, dist, world, ctx.launch_count - launches0)
    clocks = sampler.stop(t_begin, t_end) if sampler else None
    steps_timed = K * repeats
    value = rays_per_step * steps_timed / (ms * 1e-3) / 1e6

    # ---- algorithmic bytes of the timed steps: 4 B per traversal iteration + 3 B RGB per ray ----
    hit_buf = torch.empty((POSES, rows, W, 4), dtype=torch.int32, device="cuda")
    ctx.trace_stats(reset=True)
    frames_hash = 0
    for i, cams in enumerate(timed_batches):
        ctx.trace(W, H, cams, opts, out=fbs[0], hits=hit_buf)
        if i == 0:
            frames_hash = _frames_hash(torch, fbs[0], tile_first, max(1, tile_stride))
    st = ctx.trace_stats(reset=True)
    iters = _allsum(torch, dist, world, st["iterations"])
    air_iters = _allsum(torch, dist, world, st["air_iterations"])
    frames_hash = _allsum(torch, dist, world, frames_hash) & ((1 << 63) - 1)
    algo_bytes_k = 4 * iters + 3 * rays_per_step * K  # of the K distinct steps
    peaks_path = REPO / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = algo_bytes_k * repeats / (ms * 1e-3) / 1e9 / world  # per GPU, like the per-GPU peak
    fetched_bytes_k = 4 * (iters - air_iters) + 3 * rays_per_step * K  # air-phase iterations issue no load at all
    ncu = json.loads(NCU_SUMMARY.read_text()) if NCU_SUMMARY.exists() else None
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    issue = None
    if ncu:
        # warp instructions per loop iteration of the captured launch x the iterations of the timed steps: the share of the
        # issue slots (148 SMs x 4 sub-partitions x f_SM) the kernel used
        inst = ncu["warp_instructions_per_launch"] / ncu["iterations_per_launch"] * iters * repeats / world
        issue = {"warp_instructions_est": inst, "issue_slot_frac": inst / (ms * 1e-3) / (148 * 4 * sm_mhz * 1e6),
                 "how": f"warp instructions per iteration from {ncu['source']} x iterations of the timed steps / (time x 148 SMs x 4 SMSPs x {sm_mhz:.0f} MHz)",
                 "ncu_pipes": ncu.get("pipes")}
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (ncu["dram_bytes_per_launch"] / world) if ncu else None, "traffic_source": ncu["source"] if ncu else None,
                "kernel": "trace_persistent_kernel<false, kWalkFastPow2>", "peak_source": peak_src,
                "iterations_per_ray": iters / (rays_per_step * K), "air_phase_share_of_iterations": air_iters / max(1, iters),
                "algorithmic_bytes_per_launch": algo_bytes_k / K / world,
                "fetched_bytes_per_launch": fetched_bytes_k / K / world,
                "frac_by_fetched_bytes": fetched_bytes_k * repeats / (ms * 1e-3) / 1e9 / world / peak,
                "binding": issue,
                "note": "algorithmic bytes (SURVEY 8(d)) = the reference algorithm's height fetches, 4 B x loop iterations + 3 B RGB per ray; the production "
                        "walk issues no load for air-phase iterations and most of the rest hits L1/L2, so the kernel is issue/FMA-bound, not bandwidth-bound "
                        "(see `binding` and DESIGN.md section 4.1)"}

    # ---- e2e: host buffers through the C ABI: camera H2D + framebuffer D2H in the timed region.  The way a double-buffered
    # renderer calls it: hmrt_trace_host_begin(step s + 1) before hmrt_trace_host_wait(step s), two pinned host frame buffers,
    # so the traversal of one step runs under the device->host copies of the previous one.  Every step's frames are complete in
    # host memory inside the timed region.  The one-call-at-a-time form (hmrt_trace_host) is timed next to it.
    host_fbs = [torch.empty((POSES, rows, W, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
    for s in range(min(2, Wu)):
        ctx.trace_host(W, H, cams_by_step[s], opts, host_fbs[s & 1])
    e2e_repeats = max(1, min(repeats, int(math.ceil(args.min_seconds / max(1e-3, K * (ms / steps_timed) * 1e-3 * 1.1)))))

    def e2e_loop(pipelined):
        run.barrier()
        t0 = time.perf_counter()
        n = 0
        for _ in range(e2e_repeats):
            for cams in timed_batches:
                if pipelined:
                    ctx.trace_host_begin(W, H, cams, opts, host_fbs[n & 1])
                    if n:
                        ctx.trace_host_wait()
                else:
                    ctx.trace_host(W, H, cams, opts, host_fbs[n & 1])
                n += 1
        if pipelined:
            ctx.trace_host_wait()
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([sec], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        return sec, n

    sync_s, n_calls = e2e_loop(False)
    e2e_s, n_calls = e2e_loop(True)
    # the frames of the last step, as delivered to host memory, equal a device-side render of the same cameras
    ctx.trace(W, H, timed_batches[-1], opts, out=fbs[0])
    torch.cuda.synchronize()
    host_ok = bool(torch.equal(host_fbs[(n_calls - 1) & 1].cuda(), fbs[0]))
    if world > 1:
        t = torch.tensor([1 if host_ok else 0], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        host_ok = bool(t.item())
    # Both call patterns deliver every step's frames to host memory inside the timed region; the line's e2e value is the faster
    # of the two on this box at this N (`pattern` says which) and both are printed: with one GPU the traversal and the copies
    # take about equally long and two calls in flight hide one under the other; with 8 GPUs the box's aggregate
    # device->host rate (~100 GB/s) is the bound either way and back-to-back copies from all ranks contend slightly more.
    rate = lambda sec: rays_per_step * K * e2e_repeats / sec / 1e6  # noqa: E731
    per_step = lambda sec: 1e3 * sec / (K * e2e_repeats)  # noqa: E731
    best_s, pattern = (e2e_s, "two_calls_in_flight") if e2e_s <= sync_s else (sync_s, "one_call_at_a_time")
    e2e = {"value": rate(best_s), "unit": "Mrays/s", "h2d_bytes_per_step": POSES * 36,
           "d2h_bytes_per_step": POSES * W * H * 3, "ms_per_step": per_step(best_s), "steps_timed": K * e2e_repeats,
           "pattern": pattern, "host_frames_equal_device": host_ok,
           "two_calls_in_flight": {"value": rate(e2e_s), "ms_per_step": per_step(e2e_s),
                                   "what": "hmrt_trace_host_begin(step s + 1) before hmrt_trace_host_wait(step s): a double-buffered renderer, two "
                                           "alternating pinned host buffers; the traversal of a step runs under the device->host copies of the previous one"},
           "one_call_at_a_time": {"value": rate(sync_s), "ms_per_step": per_step(sync_s),
                                  "what": "hmrt_trace_host (synchronous: returns with the step's frames in host memory), the same two host buffers"},
           "note": "per-step cameras from host memory (36 B each, sent with the launches) + whole-job RGB8 framebuffers D2H into pinned host memory; "
                   "inside a call one lean launch per frame group on alternating streams, each followed by its device->host copy on a copy stream; "
                   "every step's frames are complete in host memory inside the timed region; heightmap resident"}

    # ---- frame assembly on the multi-GPU path (the reference delivers ONE complete frame per call, main.cpp:675-703):
    # every rank stores its row tiles straight into rank 0's frames over NVLink from inside the traversal kernel
    # (hmrt_trace_opts.full_frame_output + a cudaIpc mapping of rank 0's buffer): the same timed loop, whole frames on rank 0
    gather = None
    if world > 1:
        shared = []
        try:
            shared = [hd.SharedFrames(ctx, POSES, H, W, root=0) for _ in range(2)]
        except RuntimeError as e:  # raised on every rank alike (hmrt.dist.SharedFrames)
            for sf in shared:
                sf.close()
            shared = []
            gather = {"unavailable": str(e)}
    if world > 1 and shared:
        opts_full = hmrt.trace_opts(mh, tile_first=tile_first, tile_stride=tile_stride, full_frame_output=True)
        run_full = run.with_buffers([sf.tensor for sf in shared])
        for cams in timed_batches[:2]:
            run_full.issue(cams, opts_full)
        ms_full, _, _ = run_full.timed(timed_batches, opts_full, repeats)
        # the assembled frames of the first timed step, hashed on rank 0 alone, must equal the N-independent hash of the tiles
        run_full.barrier()
        with torch.cuda.stream(run_full.streams[0]):
            ctx.trace(W, H, timed_batches[0], opts_full, out=shared[0].tensor)
        run_full.barrier()
        assembled_hash = _frames_hash(torch, shared[0].tensor, 0, 1) if rank == 0 else 0
        run_full.barrier()  # rank 0 has read its frames before anybody writes into them again
        # the same with copy engines instead of in-kernel stores: compact local output, then 2-D peer copies on the step's stream
        def issue_copy(cams):
            k = run.issued & 1
            run.issue(cams, opts)
            with torch.cuda.stream(run.streams[k]):
                ctx.copy_tiles_to_frames(run.fbs[k], shared[k].tensor, W, H, POSES, tile_first, tile_stride)
        for cams in timed_batches[:2]:
            issue_copy(cams)
        run.barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        c0.record(cur)
        for s_ in run.streams:
            s_.wait_event(c0)
        for _ in range(repeats):
            for cams in timed_batches:
                issue_copy(cams)
        for s_ in run.streams:
            cur.wait_stream(s_)
        c1.record(cur)
        run.barrier()
        ms_copy = torch.tensor([c0.elapsed_time(c1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(ms_copy, op=dist.ReduceOp.MAX)
        ms_copy = float(ms_copy.item())
        shared[0].tensor.zero_() if rank == 0 else None
        run.barrier()
        with torch.cuda.stream(run.streams[0]):
            ctx.trace(W, H, timed_batches[0], opts, out=run.fbs[0])
            ctx.copy_tiles_to_frames(run.fbs[0], shared[0].tensor, W, H, POSES, tile_first, tile_stride)
        run.barrier()
        copied_hash = _frames_hash(torch, shared[0].tensor, 0, 1) if rank == 0 else 0
        for sf in shared:
            sf.close()
        gather = {"value": rays_per_step * steps_timed / (ms_full * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": ms_full / steps_timed,
                  "bytes_into_rank0_per_step": POSES * W * H * 3 * (world - 1) // world,
                  "assembled_frames_hash": f"{assembled_hash:016x}", "assembled_equals_tiles": assembled_hash == frames_hash,
                  "copy_engine_variant": {"value": rays_per_step * steps_timed / (ms_copy * 1e-3) / 1e6, "ms_per_step": ms_copy / steps_timed,
                                          "assembled_equals_tiles": copied_hash == frames_hash,
                                          "what": "compact local output + hmrt_copy_tiles_to_frames (2-D peer copies on the step's stream) instead of in-kernel stores"},
                  "what": "whole frames assembled on rank 0 inside the timed loop: every rank's traversal kernel stores its 8-row tiles at their place in "
                          "rank 0's frame buffer (cudaIpc peer mapping, 128-bit stores over NVLink); no gather pass, no collective"}

    extra = {}
    # ---- second pose family (descent-dominated: low altitude, steep pitch) ------------------------
    if not args.no_extras:
        low = [hmrt.context._cam_array(make_cameras(hmrt, s, mh, "low")) for s in range(K)]
        for cams in low[:2]:
            run.issue(cams, opts)
        r_low = run.pick_repeats(low, opts, args.min_seconds / 2)
        if world > 1:
            r_t = torch.tensor([r_low], dtype=torch.int64, device="cuda")
            dist.all_reduce(r_t, op=dist.ReduceOp.MAX)
            r_low = int(r_t.item())
        ms_low, _, _ = run.timed(low, opts, r_low)
        ctx.trace_stats(reset=True)
        for cams in low:
            ctx.trace(W, H, cams, opts, out=fbs[0], hits=hit_buf)
        st_low = ctx.trace_stats(reset=True)
        it_low = _allsum(torch, dist, world, st_low["iterations"])
        air_low = _allsum(torch, dist, world, st_low["air_iterations"])
        extra["pose_family_low"] = {
            "what": "altitude 200-500 cells above the terrain maximum, pitch -0.3..-3 (shallow to ~72 degrees down), same map, same frame size",
            "value": rays_per_step * K * r_low / (ms_low * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": ms_low / (K * r_low),
            "iterations_per_ray": it_low / (rays_per_step * K), "air_phase_share_of_iterations": air_low / max(1, it_low),
            "roofline_frac_algorithmic": (4 * it_low + 3 * rays_per_step * K) * r_low / (ms_low * 1e-3) / 1e9 / world / peak}

    # ---- tolerance mode (hmrt_set_trace_variant(2)): speed + agreement with the exact walk on the first timed batch
    if not args.no_extras:
        hits_e = torch.empty_like(hit_buf)
        rgb_e = torch.empty_like(fbs[0])
        ctx.trace(W, H, timed_batches[0], opts, out=rgb_e, hits=hits_e)
        ctx.set_trace_variant(2)
        for cams in timed_batches[:2]:
            run.issue(cams, opts)
        ms_tol, _, _ = run.timed(timed_batches, opts, max(1, repeats // 2))
        ctx.trace(W, H, timed_batches[0], opts, out=fbs[0], hits=hit_buf)
        torch.cuda.synchronize()
        cam_pos = torch.tensor([list(p) for p, _ in pose_batch(Wu, mh)], dtype=torch.float32, device="cuda")
        agree = _tolerance_agreement(torch, hits_e, rgb_e, hit_buf, fbs[0], cam_pos)
        ctx.set_trace_variant(0)
        del hits_e, rgb_e
        extra["tolerance_mode"] = {
            "what": "hmrt_set_trace_variant(2): the air phase in one closed-form step, exact descent.  Opt-in experiment, NOT used by any other "
                    "number of this line: it misses the 99.9 % hit-cell bar on grazing views (see agreement), so the default stays the bit-exact walk",
            "value": rays_per_step * K * max(1, repeats // 2) / (ms_tol * 1e-3) / 1e6, "unit": "Mrays/s",
            "agreement_with_exact_walk_rank0_rows": agree}
    del hit_buf

    # ---- CPU baseline + parity (rank 0, N = 1 only): the reference's own code on the host cores ----------
    cpu, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_threads = os.cpu_count() or 1
        host_pyr = pyr.cpu().numpy()
        assert (host_pyr[idx[0]:].view(np.uint32) == fin_host.view(np.uint32).ravel()).all(), "device finest level differs from the host terrain"
        ctx.trace(W, H, timed_batches[0], opts, out=fbs[0])
        gpu_frames = fbs[0].cpu().numpy()
        rate, rays, dt, kind, parity = cpu_sample_rate(host_pyr, mh, Wu, 12.0, n_threads, compare_with=gpu_frames)
        cpu = {"value": rate, "unit": "Mrays/s", "cores": n_threads, "kind": kind,
               "sample": f"8-row bands spread over the {POSES} 4K frames of the first timed pose batch, {rays} rays in {dt:.1f} s"}
        if not args.no_extras:
            ctx.trace(W, H, low[0], opts, out=fbs[0])
            gpu_low = fbs[0].cpu().numpy()
            _, _, _, _, par_low = cpu_sample_rate(host_pyr, mh, 0, 3.0, n_threads, family="low", compare_with=gpu_low)
            extra["pose_family_low"]["parity"] = par_low
            del gpu_low
        del gpu_frames

    # ---- the reference's own CUDA kernel, recompiled for sm_100a, on the same GPU and the same frames (reported baseline)
    gpu_ref = None
    ref_so = REPO / "oracle" / "_ref" / "libhmrt_ref_gpu.so"
    if rank == 0 and world == 1 and not args.no_cpu_baseline and ref_so.exists():
        lib = C.CDLL(str(ref_so))
        lib.hmrt_refgpu_trace.restype = C.c_int
        lib.hmrt_refgpu_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_void_p,
                                          C.c_int, C.POINTER(C.c_float)]
        one = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
        total_ms, ms_f = 0.0, C.c_float()
        cams = timed_batches[0]
        torch.cuda.synchronize()
        for i in range(POSES):
            rc = lib.hmrt_refgpu_trace(pyr.data_ptr(), None, COARSE, LEVELS, W, H, C.byref(cams, i * C.sizeof(hmrt.Camera)), 0, C.c_float(mh),
                                       one.data_ptr(), 1, C.byref(ms_f))
            if rc != 0:
                raise RuntimeError(f"reference CUDA kernel failed: {rc}")
            total_ms += ms_f.value
        gpu_ref = {"value": POSES * W * H / (total_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": total_ms,
                   "what": "the reference's own cuda_rayTrace (CudaKernel.cu:195-222) recompiled for sm_100a, one launch per frame with a legal "
                           "block shape, same B200, same pose batch (oracle/refgpu_harness.cu)"}

    # ---- rasterisation (BASELINE configs[3]) at this N -------------------------------------------
    raster = None
    if not args.no_raster:
        del fbs, run
        torch.cuda.empty_cache()
        sys.path.insert(0, str(REPO / "benchmarks"))
        import raster_pipeline

        raster = raster_pipeline.measure(torch, dist, hmrt, ctx, rank, world, args.raster_points, peak)

    if rank == 0:
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": Wu,
            "ms_per_step": ms / steps_timed, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": CONFIG, "terrain_checksum": checksum,
            "l2_persist_experiment": args.l2_persist or None,
            "arm": f"row tiles of 8 rows interleaved over {world} GPU(s); pyramid replicated by one NCCL broadcast ({bcast_ms} ms); steps alternate between two streams",
            "timed_region": {"steps_timed": steps_timed, "repeats_of_the_k_steps": repeats, "seconds": ms * 1e-3},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
            "frames_hash_first_timed_step": f"{frames_hash:016x}", "frame_gather": gather,
            "gpu_reference_baseline": gpu_ref, "raster": raster, "extra": extra,
        }
    else:
        line = None
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--min-seconds", type=float, default=1.0, help="minimum length of the timed region (the K steps are repeated)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the second pose family and the tolerance-mode measurement")
    ap.add_argument("--no-raster", action="store_true", help="skip the rasterisation sub-record")
    ap.add_argument("--raster-points", type=int, default=RASTER_POINTS)
    ap.add_argument("--l2-persist", default="", help="experiment: LEVEL:HIT_RATIO persisting-L2 window over the pyramid levels >= LEVEL")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
