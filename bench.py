#!/usr/bin/env python
"""bench.py -- Mrays/s of the heightfield ray traversal (BASELINE.json metric) on N B200s.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on; it fits one GPU):
16384^2 synthetic heightmap (8-level max-mipmap pyramid, 1.43 GB), 3840x2160 primary rays +
height-ramp shading, POSES camera poses per step.  A "step" is one pass of the traversal over one
batch of POSES frames; every step uses a different pose batch, and the pyramid is far larger
than L2, so successive steps do not replay a cached working set.

N > 1 (torchrun, one rank per GPU): rank 0 builds the pyramid and broadcasts it over NCCL/NVLink;
each frame is cut into 8-row tiles interleaved across ranks (no collective on the per-frame data
path).  Total work per step is fixed -> "scaling": "strong".

`--impl reference` times the reference's own CPU code (oracle/_ref: CudaKernel.cu host-compiled;
the oracle port if that library is absent) on the host cores over a bounded row sample of the
same workload.  Nothing under oracle/ is used by the measured GPU arm except as the separately
reported `cpu_baseline`.
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
NCU_DRAM_BYTES_PER_LAUNCH = 2.225832e9 + 0.377172e9  # profiles/ncu_trace_r01_e.txt (1 GPU, 16 frames per launch): dram__bytes_read.sum + dram__bytes_write.sum
NCU_SOURCE = "profiles/ncu_trace_r01_e.txt"
# the same capture: what actually limits the kernel (it is pipe/issue-bound, DRAM is 4 % busy)
NCU_PIPES = {"issue_slots_busy_pct": 72.1, "fma_pipe_cycles_active_pct": 61.1, "alu_pipe_pct": 38.3, "dram_throughput_pct": 4.0,
             "active_lanes_per_instruction": 27.6, "warp_instructions_per_launch": 6.572e9}
WORKLOAD = f"{R0}^2 heightmap ({LEVELS}-level max-mip pyramid, 1.43 GB), {W}x{H} primary rays + height-ramp shading, {POSES} camera poses per step"


# ------------------------------------------------------------------------------------------------
# workload definition (shared by both arms; pure host logic)

def pose_batch(step: int, max_height: float):
    """POSES deterministic camera poses for step `step`: altitude 2000-4000 cells, pitch -0.35..-0.1,
    headings all around, positions spread over the map."""
    poses = []
    for i in range(POSES):
        k = step * POSES + i
        u = (k * 0.6180339887498949) % 1.0
        v = (k * 0.7548776662466927) % 1.0
        heading = 2 * math.pi * ((k * 0.5698402909980532) % 1.0)
        pitch = -0.10 - 0.25 * ((k * 0.3819660112501051) % 1.0)
        alt = max_height + 2000.0 + 2000.0 * ((k * 0.2451223337533073) % 1.0)
        pos = (R0 * (0.2 + 0.6 * u), alt, R0 * (0.2 + 0.6 * v))
        fwd = (math.cos(heading), pitch, math.sin(heading))
        poses.append((pos, fwd))
    return poses


def make_cameras(hmrt, step, max_height):
    return [hmrt.camera(p, f, FRAME_DIM) for p, f in pose_batch(step, max_height)]


def build_terrain(torch, ctx, hmrt):
    """Synthetic terrain on the device (torch used as a buffer filler), then the product's own mip kernel."""
    res, idx, total = hmrt.pyramid_layout(COARSE, LEVELS)
    pyr = torch.zeros(total, dtype=torch.float32, device="cuda")
    fin = pyr[idx[0]:].view(R0, R0)
    xs = torch.arange(R0, device="cuda", dtype=torch.float32)
    x, z = xs[None, :], xs[:, None]
    fin.copy_(420 + 260 * torch.sin(x * 0.00121) * torch.cos(z * 0.00097) + 110 * torch.sin(x * 0.0047 + z * 0.0039)
              + 45 * torch.sin(x * 0.019) * torch.sin(z * 0.023) + 12 * torch.sin(x * 0.11 + z * 0.07))
    g = torch.Generator(device="cuda").manual_seed(1234)
    fin.add_(torch.rand((R0, R0), device="cuda", generator=g) * 3.0).clamp_(min=0)
    ctx.build_mips(pyr, COARSE, LEVELS)
    torch.cuda.synchronize()
    return pyr, float(fin.max())


def terrain_cpu_rows(n_threads_hint=None):
    """Same terrain on the host for the CPU arm (numpy; only the reference/oracle arm uses it)."""
    import torch  # CPU tensors only

    xs = torch.arange(R0, dtype=torch.float32)
    x, z = xs[None, :], xs[:, None]
    fin = (420 + 260 * torch.sin(x * 0.00121) * torch.cos(z * 0.00097) + 110 * torch.sin(x * 0.0047 + z * 0.0039)
           + 45 * torch.sin(x * 0.019) * torch.sin(z * 0.023) + 12 * torch.sin(x * 0.11 + z * 0.07))
    # the CPU arm times the traversal; it does not need the GPU arm's exact noise bits
    g = torch.Generator().manual_seed(1234)
    fin.add_(torch.rand((R0, R0), generator=g) * 3.0).clamp_(min=0)
    return fin.numpy()


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


def cpu_sample_rate(pyramid_host: np.ndarray, max_height: float, step: int, budget_s: float, n_threads: int):
    """Trace 8-row bands, spread over the frames of pose batch `step`, until `budget_s` is spent.
    Returns (Mrays/s, rays traced, seconds, kind)."""
    kind, fn, ol = _load_cpu_impl()
    cams = [ol.make_camera(p, f, FRAME_DIM) for p, f in pose_batch(step, max_height)]
    opts = ol.make_opts(max_height)
    rgb = np.zeros((H, W, 3), np.uint8)
    rows_per_call = max(8, 2 * n_threads)
    # bands interleaved over the frame height (sky and ground rows alike), cycling through the poses;
    # several passes with shifted offsets so the budget, not the band list, ends the sample
    bands = [(b * (H // 8) + off + shift) for shift in range(0, rows_per_call, 8) for off in range(0, H // 8, rows_per_call)
             for b in range(8)]
    bands = bands * len(cams)
    rays, t0 = 0, time.perf_counter()
    for j, r0 in enumerate(bands):
        cam = cams[(j // 8) % len(cams)]
        r1 = min(H, r0 + rows_per_call)
        rc = fn(pyramid_host.ctypes.data, None, COARSE, LEVELS, W, H, C.byref(cam), C.byref(opts), n_threads, r0, r1,
                rgb.ctypes.data, None)
        if rc != 0:
            raise RuntimeError(f"cpu arm failed: {rc}")
        rays += (r1 - r0) * W
        if time.perf_counter() - t0 >= budget_s:
            break
    dt = time.perf_counter() - t0
    return rays / dt / 1e6, rays, dt, kind


def host_pyramid_from_finest(fin: np.ndarray) -> np.ndarray:
    sys.path.insert(0, str(REPO / "tests"))
    import oraclelib as ol

    return ol.pyramid_from_finest(fin, LEVELS)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n_threads = os.cpu_count() or 1
    fin = terrain_cpu_rows()
    pyr = host_pyramid_from_finest(fin)
    mh = float(fin.max())
    per_step_budget = max(0.5, min(4.0, 150.0 / max(1, args.steps)))  # the whole arm stays within a few minutes for any --steps
    for s in range(args.warmup):
        cpu_sample_rate(pyr, mh, s, min(1.0, per_step_budget), n_threads)
    rays_total, t_total, kind = 0, 0.0, "port"
    for s in range(args.steps):
        _, rays, dt, kind = cpu_sample_rate(pyr, mh, args.warmup + s, per_step_budget, n_threads)
        rays_total += rays
        t_total += dt
    value = rays_total / t_total / 1e6
    sample = f"each step: 8-row-band samples of the {POSES} 4K frames of that step's pose batch until {per_step_budget:.1f} s elapse ({rays_total // max(1, args.steps)} rays/step)"
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(1, args.steps), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "l2": "inputs larger than L2 (1.43 GB pyramid)", "arm": "reference CPU code on host cores"},
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
    # heightmap: built on rank 0 with the product's kernels, replicated by one broadcast over NVLink
    res, idx, total = hmrt.pyramid_layout(COARSE, LEVELS)
    if rank == 0:
        pyr, mh = build_terrain(torch, ctx, hmrt)
    else:
        pyr, mh = torch.empty(total, dtype=torch.float32, device="cuda"), 0.0
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

    tile_first, tile_stride = hd.tiles_for_rank(rank, world)
    opts = hmrt.trace_opts(mh, tile_first=tile_first, tile_stride=tile_stride)
    rows = hmrt.rows_local(H, tile_first, tile_stride)
    fb = torch.empty((POSES, rows, W, 3), dtype=torch.uint8, device="cuda")
    cams_by_step = [hmrt.context._cam_array(make_cameras(hmrt, s, mh)) for s in range(args.warmup + args.steps)]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- value: inputs resident, device-timed --------------------------------------------------
    for s in range(args.warmup):
        ctx.trace(W, H, cams_by_step[s], opts, out=fb)
    barrier()
    t_begin = time.time()
    launches0 = ctx.launch_count
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for s in range(args.steps):
        ctx.trace(W, H, cams_by_step[args.warmup + s], opts, out=fb)
    stop.record()
    barrier()
    launches = ctx.launch_count - launches0
    ms = start.elapsed_time(stop)
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end) if sampler else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    rays_per_step = POSES * W * H
    value = rays_per_step * args.steps / (ms * 1e-3) / 1e6

    # ---- algorithmic bytes of the timed steps: 4 B per traversal iteration + 3 B RGB per ray ----
    iters = 0
    hit_buf = torch.empty((POSES, rows, W, 4), dtype=torch.int32, device="cuda")
    for s in range(args.steps):
        ctx.trace(W, H, cams_by_step[args.warmup + s], opts, out=fb, hits=hit_buf)
        iters += int((hit_buf[..., 3].view(torch.int32) >> 8).to(torch.int64).sum().item())
    del hit_buf
    if world > 1:
        it = torch.tensor([iters], dtype=torch.int64, device="cuda")
        dist.all_reduce(it, op=dist.ReduceOp.SUM)
        iters = int(it.item())
    algo_bytes = 4 * iters + 3 * rays_per_step * args.steps
    peaks_path = REPO / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = algo_bytes / (ms * 1e-3) / 1e9 / world  # per GPU, like the per-GPU peak
    # DRAM traffic of one single-GPU launch (16 frames) from the committed ncu capture: dram__bytes_read.sum + dram__bytes_write.sum
    traffic = NCU_DRAM_BYTES_PER_LAUNCH / world
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": NCU_SOURCE, "kernel": "trace_persistent_kernel<false, kWalkFastPow2>", "peak_source": peak_src,
                "iterations_per_ray": iters / (rays_per_step * args.steps), "ncu_pipes": NCU_PIPES,
                "algorithmic_bytes_per_launch": algo_bytes / max(1, args.steps) / world,
                "note": "issue-bound, not bandwidth-bound: algorithmic bytes are the reference algorithm's height fetches (4 B x loop iterations + 3 B RGB per ray); "
                        "most of them hit L1/L2, so DRAM traffic is ~13x smaller (see DESIGN.md section 4.1)"}

    # ---- e2e: host buffers through the C ABI (hmrt_trace_host): camera H2D + framebuffer D2H in the timed region
    host_fb = torch.empty((POSES, rows, W, 3), dtype=torch.uint8).pin_memory()
    for s in range(min(2, args.warmup)):
        ctx.trace_host(W, H, cams_by_step[s], opts, host_fb)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        ctx.trace_host(W, H, cams_by_step[args.warmup + s], opts, host_fb)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": rays_per_step * args.steps / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": POSES * 36,
           "d2h_bytes_per_step": POSES * W * H * 3, "ms_per_step": 1e3 * e2e_s / args.steps,
           "note": "hmrt_trace_host: per-step cameras from host memory (36 B each, sent with the launch) + whole-job RGB8 framebuffers D2H into pinned host memory, copy of frame f overlapped with the traversal of frame f+1; heightmap resident"}

    # ---- CPU baseline (rank 0, N = 1 only): the reference's own code on the host cores ----------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_threads = os.cpu_count() or 1
        host_pyr = pyr.cpu().numpy()
        rate, rays, dt, kind = cpu_sample_rate(host_pyr, mh, args.warmup, 12.0, n_threads)
        cpu = {"value": rate, "unit": "Mrays/s", "cores": n_threads, "kind": kind,
               "sample": f"8-row bands spread over the {POSES} 4K frames of the first timed pose batch, {rays} rays in {dt:.1f} s"}

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
        cams = cams_by_step[args.warmup]
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

    if rank == 0:
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "l2": "inputs larger than L2 (1.43 GB pyramid); a different pose batch every step",
                       "parallelism": f"row tiles of 8 rows interleaved over {world} GPU(s); pyramid replicated by one NCCL broadcast",
                       "levels": LEVELS, "frame_dimension": FRAME_DIM, "pyramid_broadcast_ms": bcast_ms},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "gpu_reference_baseline": gpu_ref,
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
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
