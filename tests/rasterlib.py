"""Synthetic point clouds for the rasterisation tests (host logic only)."""
from __future__ import annotations

import ctypes as C

import numpy as np

import oraclelib as ol
from hmrt import las


def synthetic_las(n, r0, cell=2.0, point_format=2, seed=0, record_len=None, frac_outside=0.05, frac_class7=0.05,
                  scale=(0.01, 0.01, 0.01), offset=(635000.0, 848000.0, 400.0)):
    """n LAS records over a r0 x r0 grid of `cell`-sized cells: terrain-like Z, some points outside the
    section, some classified 7 (noise) -- the two rejection rules of main.cpp:209."""
    rng = np.random.default_rng(seed)
    ext = r0 * cell
    x = rng.random(n) * ext * (1 + 2 * frac_outside) - ext * frac_outside
    y = rng.random(n) * ext * (1 + 2 * frac_outside) - ext * frac_outside
    z = 20 + 10 * np.sin(x * 0.01) * np.cos(y * 0.013) + rng.random(n) * 3
    X = np.round(x / scale[0]).astype(np.int32)
    Y = np.round(y / scale[1]).astype(np.int32)
    Z = np.round(z / scale[2]).astype(np.int32)
    cls = rng.integers(0, 32, n).astype(np.uint8)
    cls[rng.random(n) < frac_class7] = 7
    cls |= (rng.integers(0, 8, n).astype(np.uint8) << 5)  # upper 3 bits are flags, not class
    rgb = rng.integers(0, 65536, (n, 3)).astype(np.uint16)
    rec = las.encode_points(X, Y, Z, point_format, cls, rgb, record_len)
    # header min = offset + min(raw)*scale like a LAS writer would store it
    mn = (offset[0], offset[1], offset[2] + float(Z.min()) * scale[2])
    mx = (offset[0] + ext, offset[1] + ext, offset[2] + float(Z.max()) * scale[2])
    hdr = las.LasHeader(point_format, rec.shape[1], n, scale, offset, mn, mx)
    return hdr, rec


def oracle_rasterise(hdr, rec, coarse, levels, cell=(2.0, 2.0, 2.0), origin=(0.0, 0.0), with_colors=True):
    res, idx, total = ol.pyramid_layout(coarse, levels)
    pyr = np.zeros(total, np.float32)
    cmap = np.zeros((res[0], res[0], 3), np.uint8) if with_colors else None
    xf = hdr.transform(cell, origin)
    rec = np.ascontiguousarray(rec)
    rc = ol.oracle().hmrt_oracle_rasterise_las(rec.ctypes.data, rec.shape[0], rec.shape[1], hdr.point_format, C.byref(xf),
                                               pyr.ctypes.data, coarse, levels, cmap.ctypes.data if with_colors else None)
    assert rc == 0
    return pyr, cmap
