#!/usr/bin/env python
"""Device->host copy rate of one GPU for the bench's frame traffic (398 MB per step in 24.9 MB frames), over 1 / 2 / 4
copy streams and pinned host buffers (development probe: is hmrt_trace_host bound by one copy engine or by the link?)."""
import json
import time

import torch


def main():
    torch.cuda.set_device(0)
    n, frame = 16, 3840 * 2160 * 3
    dev = torch.empty((n, frame), dtype=torch.uint8, device="cuda")
    host = torch.empty((n, frame), dtype=torch.uint8).pin_memory()
    out = {}
    for streams in (1, 2, 4):
        ss = [torch.cuda.Stream() for _ in range(streams)]
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for r in range(5):
                for f in range(n):
                    with torch.cuda.stream(ss[f % streams]):
                        host[f].copy_(dev[f], non_blocking=True)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 5
        out[f"{streams}_streams_GBps"] = n * frame / dt / 1e9
    # one big copy
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for r in range(5):
        host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    out["one_copy_GBps"] = n * frame / ((time.perf_counter() - t0) / 5) / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
