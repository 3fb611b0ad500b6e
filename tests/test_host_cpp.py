"""C++ host side (gpu-heightmap-raytracer_b200/host): the drop-in CudaSpace header compiles against the
reference's own call sites, and the host helpers (LAS header reader, seeded PointdataGenerator, camera
rules) behave like the reference's host code.  CPU only; the CLI end-to-end test is in test_gpu_host_cli.py."""
import ctypes as C
import math
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oraclelib as ol
import rasterlib as rl
from hmrt import las

REPO = Path(__file__).resolve().parent.parent
HOST = REPO / "gpu-heightmap-raytracer_b200" / "host"
REF_INC = Path("/root/reference/GPUHeightmapRaytracer/inc")


@pytest.fixture(scope="module")
def hostlib():
    subprocess.run(["make", "-s", "-C", str(HOST), "libhmrt_host.so"], check=True)
    lib = C.CDLL(str(HOST / "libhmrt_host.so"))
    lib.hmrt_host_pdg_generate.argtypes = [C.c_int, C.c_uint64, C.c_void_p]
    lib.hmrt_host_las_info.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p]
    lib.hmrt_host_las_scene.argtypes = [C.c_char_p, C.c_void_p]
    lib.hmrt_host_camera_step.argtypes = [C.c_void_p] + [C.c_float] * 6 + [C.c_void_p, C.c_float]
    lib.hmrt_host_camera_step.restype = None
    return lib


@pytest.mark.skipif(not REF_INC.exists(), reason="reference include tree (GLM) not present")
def test_dropin_header_compiles_against_reference_call_sites():
    """main.cpp:1014, :686, :1024 with the reference's glm types, unchanged, against host/CudaKernel.cuh."""
    subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-w", f"-I{REF_INC}", f"-I{HOST}",
                    str(REPO / "tests" / "hostsim" / "shim_callsites.cpp")], check=True)


def test_host_pdg_generator_equals_oracle(hostlib):
    n = 128
    a = np.zeros(((n + 1) ** 2, 3), np.float32)
    b = np.zeros_like(a)
    assert hostlib.hmrt_host_pdg_generate(n, 99, a.ctypes.data) == 0
    assert ol.oracle().hmrt_oracle_pdg_generate(n, 99, b.ctypes.data) == 0
    assert (a.view(np.uint32) == b.view(np.uint32)).all()
    assert hostlib.hmrt_host_pdg_generate(100, 1, a.ctypes.data) != 0


def test_host_las_reader_and_scene(hostlib, tmp_path):
    hdr, rec = rl.synthetic_las(1000, 64, seed=3)
    path = tmp_path / "t.las"
    las.write_las(path, hdr, rec)
    info = np.zeros(12, np.float64)
    meta = np.zeros(4, np.uint64)
    assert hostlib.hmrt_host_las_info(str(path).encode(), info.ctypes.data, meta.ctypes.data) == 0
    assert tuple(info[0:3]) == hdr.scale and tuple(info[3:6]) == hdr.offset
    assert tuple(info[6:9]) == hdr.min and tuple(info[9:12]) == hdr.max
    assert list(meta) == [hdr.point_format, hdr.record_len, hdr.n_points, 227]
    scene = np.zeros(9, np.float32)
    assert hostlib.hmrt_host_las_scene(str(path).encode(), scene.ctypes.data) == 0
    # readLASHeader, main.cpp:153-164: cell 2.0, boundaries = extent / cell, camera on the first point
    assert list(scene[0:3]) == [2.0, 2.0, 2.0]
    assert scene[3] == np.float32((hdr.max[0] - hdr.min[0]) / 2.0)
    X0 = int(rec[0, 0:4].view("<i4")[0])
    assert scene[5] == np.float32(((X0 * hdr.scale[0] + hdr.offset[0]) - hdr.min[0]) / 2.0)
    assert scene[8] == np.float32(np.float32(hdr.max[2] - hdr.min[2]) / np.float32(2.0))
    assert hostlib.hmrt_host_las_info(b"/nonexistent.las", info.ctypes.data, meta.ctypes.data) != 0


def test_camera_rules(hostlib):
    """moveCamera clamps (main.cpp:758-771); rotateCamera keeps |forward| = 1 (main.cpp:779-780)."""
    st = np.array([10, 5, 10, 0, -0.6689647, 0.7432941], np.float32)
    bounds = np.array([100, 100], np.float32)
    hostlib.hmrt_host_camera_step(st.ctypes.data, 250.0, 0.0, 0.0, 1.0, 0.0, 0.1, bounds.ctypes.data, 50.0)
    assert st[0] == 10 and abs(st[2] - 35.0) < 1e-4 and st[1] == 5       # forward motion is horizontal
    assert abs(np.linalg.norm(st[3:6]) - 1) < 1e-6
    assert abs(math.atan2(st[3], st[5]) - 0.1) < 1e-5                       # yaw of 1 rad/s * 0.1 s about +y
    hostlib.hmrt_host_camera_step(st.ctypes.data, 25000.0, 0.0, 5000.0, 0.0, 0.0, 5.0, bounds.ctypes.data, 50.0)
    assert st[2] < 100 and st[0] < 100 and st[1] == 200.0                   # clamped: boundaries, 4 * max_height
