"""GPU-side helpers for the parity tests: everything goes through the C ABI (hmrt.Context)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

import oraclelib as ol


def upload_scene(ctx, sc):
    pyr = torch.from_numpy(sc["pyramid"]).cuda()
    cmap = torch.from_numpy(sc["color_map"]).cuda() if sc.get("color_map") is not None else None
    ctx.set_heightmap(pyr, cmap, sc["coarse"], sc["levels"], sc["max_height"])
    return pyr, cmap


def gpu_trace(ctx, W, H, cams, opts, hits=True):
    rgb, h = ctx.trace(W, H, cams, opts, hits=hits)
    torch.cuda.synchronize()
    rgb = rgb.cpu().numpy()
    if h is not None:
        h = h.cpu().numpy().view(np.uint32)
        h = np.ascontiguousarray(h).view(ol.hit_dtype).reshape(h.shape[:-1])
    return rgb, h
