#!/usr/bin/env python
"""Per-frame window preparation (preparePointBuffer + copyPointBuffer, main.cpp:459-625) at the reference's default size:
sections of 4096^2 cells (coarse 32, 8 levels: 89.5 MB of floats + 50.3 MB of colours per window).

  new        hmrt_compose_window: the window is gathered on the device out of four resident section pyramids
  reference  the four-memcpy-loops composition on one host core (the oracle's restatement of main.cpp:519-618) followed by
             the two cudaMemcpy uploads of main.cpp:622-624 from pageable memory, timed on the same box

One JSON line: milliseconds, GB/s of algorithmic bytes (read + write of every window byte) against the measured HBM peak.
"""
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "gpu-heightmap-raytracer_b200"))
sys.path.insert(0, str(REPO / "tests"))

import torch  # noqa: E402

import hmrt  # noqa: E402


def main():
    coarse, levels = 32, 8
    res, idx, total = hmrt.pyramid_layout(coarse, levels)
    r0 = res[0]
    ctx = hmrt.Context(0)
    variant = int(sys.argv[sys.argv.index("--variant") + 1]) if "--variant" in sys.argv else 0
    ctx.set_window_variant(variant)
    g = torch.Generator(device="cuda").manual_seed(3)
    secs = [[torch.rand(total, device="cuda", generator=g) * 90 for _ in range(2)] for _ in range(2)]
    cols = [[torch.randint(0, 256, (r0, r0, 3), dtype=torch.uint8, device="cuda", generator=g) for _ in range(2)] for _ in range(2)]
    win = torch.empty(total, dtype=torch.float32, device="cuda")
    wcol = torch.empty((r0, r0, 3), dtype=torch.uint8, device="cuda")
    cells = [(13, 22), (0, 0), (31, 5), (7, 31), (16, 16)]
    for cx, cy in cells:
        ctx.compose_window(secs, cols, coarse, levels, cx, cy, win, wcol)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for k in range(reps):
        cx, cy = cells[k % len(cells)]
        ctx.compose_window(secs, cols, coarse, levels, cx, cy, win, wcol)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bytes_rw = 2 * (total * 4 + r0 * r0 * 3)
    peaks = REPO / "MEASURED_PEAKS.json"
    peak = float(json.loads(peaks.read_text())["hbm_gbs"]) if peaks.exists() else 6650.0

    # the reference's flow on this box: host composition on one core + two pageable uploads
    import oraclelib as ol

    hs = [[s.cpu().numpy() for s in row] for row in secs]
    hc = [[c.cpu().numpy() for c in row] for row in cols]
    ol.oracle_compose_window(hs, hc, coarse, levels, 13, 22)
    t0 = time.perf_counter()
    hw, hwc = ol.oracle_compose_window(hs, hc, coarse, levels, 7, 31)
    t_compose = time.perf_counter() - t0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        win.copy_(torch.from_numpy(hw))       # cudaMemcpy H2D from pageable memory, main.cpp:623
        wcol.copy_(torch.from_numpy(hwc))     # :624
        torch.cuda.synchronize()
    t_upload = (time.perf_counter() - t0) / 3
    line = {
        "workload": f"window of a {r0}^2-cell section grid (coarse {coarse}, {levels} levels): {total * 4 / 1e6:.1f} MB pyramid + {r0 * r0 * 3 / 1e6:.1f} MB colours per frame",
        "formulation": "TMA bulk copies (cp.async.bulk), one launch" if variant == 0 else "per-thread 128-bit gather, two launches",
        "compose_window_ms": ms, "algorithmic_GBps": bytes_rw / (ms * 1e-3) / 1e9, "hbm_peak_GBps": peak,
        "roofline_frac": bytes_rw / (ms * 1e-3) / 1e9 / peak,
        "reference_flow_ms": {"host_compose_1_core": 1e3 * t_compose, "upload_pageable_h2d": 1e3 * t_upload, "total": 1e3 * (t_compose + t_upload)},
        "speedup_vs_reference_flow": 1e3 * (t_compose + t_upload) / ms,
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
