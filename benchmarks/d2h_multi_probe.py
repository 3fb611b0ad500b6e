#!/usr/bin/env python
"""Aggregate device->host copy rate of all GPUs of a box (development probe, run under torchrun): every rank copies 400 MB
frames-sized buffers into pinned host memory at the same time -- (a) buffers allocated wherever the process happens to run,
(b) after binding the process to the CPUs `nvidia-smi topo -m` lists as local to its GPU (first-touch NUMA placement of the
pinned pages), (c) one rank at a time.  Answers: is the 8-GPU end-to-end rate bound by the links or by host-memory placement?"""
import json
import os
import re
import subprocess
import time

import torch
import torch.distributed as dist


def cpu_affinity_of_gpu(index):
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=60).stdout
    except Exception:
        return None, None
    for line in out.splitlines():
        cols = line.split("\t")
        if cols and cols[0].strip() == f"GPU{index}":
            for c in cols[1:]:
                c = c.strip()
                if re.fullmatch(r"[0-9,\-]+", c) and ("-" in c or "," in c):
                    cpus = set()
                    for part in c.split(","):
                        a, _, b = part.partition("-")
                        cpus.update(range(int(a), int(b or a) + 1))
                    return cpus, out
    return None, out


def rate(dev, host, reps, group_barrier):
    torch.cuda.synchronize()
    if group_barrier:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    return dev.numel() * reps / (time.perf_counter() - t0) / 1e9


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nbytes = 400 << 20
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    out = {"world": world, "cpus_allowed_before": len(os.sched_getaffinity(0))}
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    rate(dev, host, 2, world > 1)
    mine = {"default_together": rate(dev, host, 10, world > 1)}
    cpus, topo = cpu_affinity_of_gpu(local)
    if cpus:
        try:
            os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or os.sched_getaffinity(0))
        except OSError:
            pass
    del host
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.fill_(1)  # first touch from a local CPU
    rate(dev, host, 2, world > 1)
    mine["bound_together"] = rate(dev, host, 10, world > 1)
    mine["cpus_local"] = len(cpus) if cpus else None
    # one rank at a time
    alone = 0.0
    for r in range(world):
        if world > 1:
            dist.barrier()
        if r == rank:
            alone = rate(dev, host, 5, False)
    mine["bound_alone"] = alone
    vals = torch.tensor([mine["default_together"], mine["bound_together"], mine["bound_alone"]], dtype=torch.float64, device="cuda")
    allv = [torch.empty_like(vals) for _ in range(world)]
    if world > 1:
        dist.all_gather(allv, vals)
    else:
        allv = [vals]
    if rank == 0:
        out["per_rank_GBps"] = {k: [round(float(v[i]), 1) for v in allv] for i, k in enumerate(["default_together", "bound_together", "bound_alone"])}
        out["aggregate_GBps"] = {k: round(sum(v), 1) for k, v in out["per_rank_GBps"].items()}
        out["cpus_local_rank0"] = mine["cpus_local"]
        out["topo"] = topo
        try:
            out["numa"] = subprocess.run(["lscpu"], capture_output=True, text=True, timeout=30).stdout.split("NUMA", 1)[-1][:400]
        except Exception:
            pass
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
