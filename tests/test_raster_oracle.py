"""CPU checks of the rasterisation oracle (restatement of main.cpp:193-234), the LAS codec and the
seeded PointdataGenerator restatement.  The reference rasteriser itself cannot be built here
(libLAS binary, <Windows.h>), so these pin the restatement through its defining properties."""
import ctypes as C
import struct

import numpy as np
import pytest

import oraclelib as ol
import rasterlib as rl
from hmrt import las


def test_pyramid_layout_matches_reference_tables():
    res = (C.c_int * 8)()
    idx = (C.c_int64 * 8)()
    total = C.c_int64()
    assert ol.oracle().hmrt_oracle_pyramid_layout(32, 8, res, idx, C.byref(total)) == 0
    # main.cpp:71,83: point_buffer_resolution 32, LOD_levels 8 -> 4096^2 finest, stride_x = 21845
    assert list(res) == [4096 >> i for i in range(8)]
    assert total.value == 32 * 32 * 21845
    assert list(idx) == ol.pyramid_layout(32, 8)[1]


def test_early_break_propagation_equals_max_pyramid():
    """main.cpp:227-233 keeps parent >= child, so its result is 'finest = max, level i+1 = 2x2 max'."""
    hdr, rec = rl.synthetic_las(200_000, 256, seed=1)
    pyr, _ = rl.oracle_rasterise(hdr, rec, 256 >> 5, 6)
    res, idx, total = ol.pyramid_layout(256 >> 5, 6)
    finest = pyr[idx[0]:idx[0] + 256 * 256].reshape(256, 256)
    assert (ol.pyramid_from_finest(finest, 6) == pyr).all()
    chk = pyr.copy()
    assert ol.oracle().hmrt_oracle_build_mips(chk.ctypes.data, 256 >> 5, 6) == 0
    assert (chk == pyr).all()


def test_heights_are_point_order_independent_colours_are_last_writer():
    hdr, rec = rl.synthetic_las(50_000, 128, seed=2)
    pyr_a, cm_a = rl.oracle_rasterise(hdr, rec, 128 >> 4, 5)
    perm = np.random.default_rng(0).permutation(len(rec))
    pyr_b, cm_b = rl.oracle_rasterise(hdr, rec[perm], 128 >> 4, 5)
    assert (pyr_a == pyr_b).all()
    assert (cm_a != cm_b).any()  # colour is order dependent by construction (main.cpp:223-224)


def test_rejection_rules_and_floor():
    """Outside-section and class-7 points are skipped (main.cpp:209); empty cells stay +0 (main.cpp:259)."""
    scale, offset = (0.5, 0.5, 0.5), (0.0, 0.0, 0.0)
    # cell size 2, 4x4 grid: points at x in {1, 3, 9 (outside), -1 (outside)}
    X = np.array([2, 6, 18, -2, 2], np.int32)
    Y = np.array([2, 2, 2, 2, 6], np.int32)
    Z = np.array([10, 20, 30, 40, 50], np.int32)
    cls = np.array([0, 7, 0, 0, 2 | 0xE0], np.uint8)
    rec = las.encode_points(X, Y, Z, 0, cls)
    hdr = las.LasHeader(0, 20, 5, scale, offset, (0.0, 0.0, 0.0), (8.0, 8.0, 25.0))
    pyr, cmap = rl.oracle_rasterise(hdr, rec, 1, 3)
    res, idx, _ = ol.pyramid_layout(1, 3)
    fin = pyr[idx[0]:].reshape(4, 4)
    expect = np.zeros((4, 4), np.float32)
    expect[0, 0] = 2.5   # point 0: z = 10*0.5 = 5 -> /2
    expect[1, 0] = 12.5  # point 4: class 2 (upper bits are flags): kept
    assert (fin == expect).all()
    assert pyr[idx[2]] == 12.5 and (np.signbit(pyr) == False).all()  # noqa: E712


def test_color16_conversion_rule():
    """CudaSpace::Color(unsigned short...) = floor(c / 65535.f * 255.f) (CudaKernel.cuh:41-46)."""
    vals = np.array([0, 1, 256, 257, 32767, 32768, 65534, 65535], np.uint16)
    n = len(vals)
    rgb = np.stack([vals, vals, vals], axis=1)
    X = np.arange(n, dtype=np.int32) * 2 + 1
    rec = las.encode_points(X, np.ones(n, np.int32), np.ones(n, np.int32), 2, None, rgb)
    hdr = las.LasHeader(2, 26, n, (1.0, 1.0, 1.0), (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), (16.0, 16.0, 1.0))
    _, cmap = rl.oracle_rasterise(hdr, rec, 8, 1)
    want = np.floor(vals.astype(np.float32) / np.float32(65535) * np.float32(255)).astype(np.uint8)
    assert (cmap[0, :n, 0] == want).all() and want[-1] == 255 and want[0] == 0


def test_las_codec_known_answer(tmp_path):
    """Hand-assembled LAS 1.2 header + one format-2 record, byte by byte."""
    head = bytearray(227)
    head[0:4] = b"LASF"
    head[24:26] = bytes([1, 2])
    struct.pack_into("<H", head, 94, 227)
    struct.pack_into("<I", head, 96, 227)
    head[104] = 2
    struct.pack_into("<H", head, 105, 26)
    struct.pack_into("<I", head, 107, 1)
    struct.pack_into("<3d", head, 131, 0.01, 0.01, 0.001)
    struct.pack_into("<3d", head, 155, 1000.0, 2000.0, 50.0)
    struct.pack_into("<6d", head, 179, 1100.0, 1000.0, 2100.0, 2000.0, 60.0, 50.0)
    rec = bytearray(26)
    struct.pack_into("<iii", rec, 0, 12345, -678, 9000)
    rec[15] = 0x27  # class 7 with a flag bit
    struct.pack_into("<HHH", rec, 20, 65535, 32768, 1)
    path = tmp_path / "one.las"
    path.write_bytes(bytes(head) + bytes(rec))
    h, r = las.read_las(path)
    assert (h.point_format, h.record_len, h.n_points) == (2, 26, 1)
    assert h.scale == (0.01, 0.01, 0.001) and h.offset == (1000.0, 2000.0, 50.0)
    assert h.min == (1000.0, 2000.0, 50.0) and h.max == (1100.0, 2100.0, 60.0)
    assert bytes(r[0]) == bytes(rec)
    # round trip through the writer
    las.write_las(tmp_path / "two.las", h, r)
    h2, r2 = las.read_las(tmp_path / "two.las")
    assert h2 == h and (r2 == r).all()
    with pytest.raises(ValueError):
        las.parse_header(b"NOPE" + bytes(223))


def test_pdg_generator_is_seeded_and_shaped():
    """PointdataGenerator/main.cpp:72-184: (n+1)^2 points at integer (i, j); z*10 except last row/col."""
    n = 64
    a = np.zeros(((n + 1) ** 2, 3), np.float32)
    b = np.zeros_like(a)
    c = np.zeros_like(a)
    assert ol.oracle().hmrt_oracle_pdg_generate(n, 42, a.ctypes.data) == 0
    assert ol.oracle().hmrt_oracle_pdg_generate(n, 42, b.ctypes.data) == 0
    assert ol.oracle().hmrt_oracle_pdg_generate(n, 43, c.ctypes.data) == 0
    assert (a == b).all() and (a != c).any()
    g = a.reshape(n + 1, n + 1, 3)
    assert (g[..., 0] == np.arange(n + 1, dtype=np.float32)[:, None]).all()
    assert (g[..., 1] == np.arange(n + 1, dtype=np.float32)[None, :]).all()
    assert g[:-1, :-1, 2].mean() > 5 * g[-1, :, 2].mean()  # scaleData skips the last row/column (PDG:177-179)
    assert ol.oracle().hmrt_oracle_pdg_generate(63, 1, a.ctypes.data) != 0
