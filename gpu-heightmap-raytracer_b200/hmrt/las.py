"""Minimal LAS 1.2 codec (public header block + point data record formats 0-3).

Host-side data format either side of the rasterisation path: the reference reads LAS through
libLAS 1.8.0 (main.cpp:124-168, :174-224), which is not available as a library here; this module
reads/writes the subset the reference touches -- header min/max/scale/offset, and per point
X/Y/Z (int32), classification (low 5 bits of byte 15) and RGB (uint16 x 3 in formats 2 and 3).
Records stay raw bytes: decoding happens in the CUDA scatter kernel (csrc/raster.cu).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np

from ._abi import LasTransform

HEADER_SIZE = 227
RECORD_MIN_LEN = {0: 20, 1: 28, 2: 26, 3: 34}
RGB_OFFSET = {0: None, 1: None, 2: 20, 3: 28}


@dataclass
class LasHeader:
    point_format: int
    record_len: int
    n_points: int
    scale: tuple
    offset: tuple
    min: tuple
    max: tuple
    offset_to_points: int = HEADER_SIZE

    def transform(self, cell_size=(2.0, 2.0, 2.0), origin=(0.0, 0.0)) -> LasTransform:
        """hmrt_las_transform for this file; cell_size 2.0 is the reference's (main.cpp:154-155)."""
        xf = LasTransform()
        xf.scale[:] = self.scale
        xf.offset[:] = self.offset
        xf.min[:] = self.min
        xf.cell_size[:] = [float(np.float32(c)) for c in cell_size]
        xf.origin[:] = [float(np.float32(o)) for o in origin]
        return xf


def pack_header(h: LasHeader) -> bytes:
    b = bytearray(HEADER_SIZE)
    b[0:4] = b"LASF"
    b[24] = 1
    b[25] = 2
    b[26:26 + 4] = b"hmrt"
    b[58:58 + 4] = b"hmrt"
    struct.pack_into("<H", b, 94, HEADER_SIZE)
    struct.pack_into("<I", b, 96, h.offset_to_points)
    struct.pack_into("<I", b, 100, 0)
    b[104] = h.point_format
    struct.pack_into("<H", b, 105, h.record_len)
    struct.pack_into("<I", b, 107, h.n_points & 0xFFFFFFFF)
    struct.pack_into("<3d", b, 131, *h.scale)
    struct.pack_into("<3d", b, 155, *h.offset)
    struct.pack_into("<6d", b, 179, h.max[0], h.min[0], h.max[1], h.min[1], h.max[2], h.min[2])
    return bytes(b)


def parse_header(buf: bytes) -> LasHeader:
    if len(buf) < HEADER_SIZE or buf[0:4] != b"LASF":
        raise ValueError("not a LAS file")
    fmt = buf[104] & 0x3F
    if buf[104] & 0xC0:
        raise ValueError("compressed (LAZ) point data is not supported")
    if fmt not in RECORD_MIN_LEN:
        raise ValueError(f"unsupported point data format {fmt}")
    (rec_len,) = struct.unpack_from("<H", buf, 105)
    if rec_len < RECORD_MIN_LEN[fmt]:
        raise ValueError("record length shorter than the point format")
    (n,) = struct.unpack_from("<I", buf, 107)
    (off,) = struct.unpack_from("<I", buf, 96)
    scale = struct.unpack_from("<3d", buf, 131)
    offset = struct.unpack_from("<3d", buf, 155)
    mx, mnx, my, mny, mz, mnz = struct.unpack_from("<6d", buf, 179)
    return LasHeader(fmt, rec_len, n, scale, offset, (mnx, mny, mnz), (mx, my, mz), off)


def encode_points(X, Y, Z, point_format=2, classification=None, rgb=None, record_len=None) -> np.ndarray:
    """Raw LAS records [n, record_len] uint8 from int32 X/Y/Z (+ classification, uint16 rgb [n,3])."""
    n = len(X)
    record_len = record_len or RECORD_MIN_LEN[point_format]
    rec = np.zeros((n, record_len), dtype=np.uint8)
    rec[:, 0:4] = np.asarray(X, dtype="<i4").view(np.uint8).reshape(n, 4)
    rec[:, 4:8] = np.asarray(Y, dtype="<i4").view(np.uint8).reshape(n, 4)
    rec[:, 8:12] = np.asarray(Z, dtype="<i4").view(np.uint8).reshape(n, 4)
    if classification is not None:
        rec[:, 15] = np.asarray(classification, dtype=np.uint8)
    ro = RGB_OFFSET[point_format]
    if rgb is not None and ro is not None:
        rec[:, ro:ro + 6] = np.ascontiguousarray(rgb, dtype="<u2").view(np.uint8).reshape(n, 6)
    return rec


def write_las(path, header: LasHeader, records: np.ndarray) -> None:
    with open(path, "wb") as f:
        f.write(pack_header(header))
        f.write(np.ascontiguousarray(records, dtype=np.uint8).tobytes())


def read_las(path):
    """(LasHeader, records uint8 [n, record_len]) -- raw records, ready for Context.scatter_las."""
    with open(path, "rb") as f:
        head = f.read(HEADER_SIZE)
        h = parse_header(head)
        f.seek(h.offset_to_points)
        raw = np.frombuffer(f.read(h.n_points * h.record_len), dtype=np.uint8)
    if raw.size != h.n_points * h.record_len:
        raise ValueError("truncated LAS point data")
    return h, raw.reshape(h.n_points, h.record_len)


def read_pdg_text(path) -> np.ndarray:
    """PointdataGenerator output (PointdataGenerator/main.cpp:186-205): 'x y z' per line -> float32 [n,3]."""
    return np.loadtxt(path, dtype=np.float32).reshape(-1, 3)
