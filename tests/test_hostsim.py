"""The product's per-ray arithmetic (csrc/ray_core.cuh) compiled for the host must be
bit-identical to the oracle: checks the kernel FORMULATION (power-of-two constants, closed-form
level offsets, cell-index reuse) without a GPU.  The GPU tests then check the compiled kernel."""
import numpy as np
import pytest

import oraclelib as ol


@pytest.mark.parametrize("name", list(ol.SCENES))
@pytest.mark.parametrize("mode", ["ramp", "shadow", "colormap+shadow"])
def test_core_equals_oracle(name, mode):
    sc = ol.scene(name, seed=3)
    W, H = 128, 96
    for cam in ol.cameras_for(sc, 6):
        opts = ol.make_opts(sc["max_height"], use_color_map="colormap" in mode, shadows="shadow" in mode)
        a = ol.cpu_trace(ol.oracle().hmrt_oracle_trace, sc["pyramid"], sc["color_map"], sc["coarse"], sc["levels"], W, H, cam, opts)
        b = ol.cpu_trace(ol.hostsim().hostsim_trace, sc["pyramid"], sc["color_map"], sc["coarse"], sc["levels"], W, H, cam, opts)
        ol.assert_same_trace(a, b, f"{name}/{mode}")


def test_edge_cases():
    """Hand-made cases: flat map, single column, ray along +x, camera inside a column, start outside."""
    r0, levels = 64, 4
    flat = np.zeros((r0, r0), np.float32)
    col = flat.copy()
    col[20, 40] = 30.0
    for fin in (flat, col, np.full((r0, r0), 5.0, np.float32)):
        pyr = ol.pyramid_from_finest(fin, levels)
        for pos, fwd in [((32.5, 10.0, 10.5), (0.0, -0.2, 1.0)),      # towards the column
                         ((1.5, 3.0, 20.5), (1.0, -0.05, 0.0)),       # exactly along +x (dir.z == 0)
                         ((40.5, 2.0, 20.5), (0.3, -0.1, 0.4)),       # camera inside the column's cell
                         ((32.0, 50.0, 32.0), (0.0, -1.0, 1e-3)),     # nearly straight down, on a cell edge
                         ((-30.0, 10.0, 32.0), (-1.0, -0.1, 0.0)),    # image plane starts outside the grid
                         ((32.5, 10.0, 32.5), (0.2, 0.9, 0.1))]:      # looking up: background
            cam = ol.make_camera(pos, fwd)
            opts = ol.make_opts(30.0, shadows=True)
            a = ol.cpu_trace(ol.oracle().hmrt_oracle_trace, pyr, None, r0 >> (levels - 1), levels, 64, 48, cam, opts)
            b = ol.cpu_trace(ol.hostsim().hostsim_trace, pyr, None, r0 >> (levels - 1), levels, 64, 48, cam, opts)
            ol.assert_same_trace(a, b, f"edge {pos} {fwd}")


def test_empty_heightmap_is_all_background_when_looking_up():
    r0, levels = 64, 3
    pyr = ol.pyramid_from_finest(np.zeros((r0, r0), np.float32), levels)
    cam = ol.make_camera((32.5, 10.0, 32.5), (0.0, 0.95, 0.3))
    rgb, hits = ol.cpu_trace(ol.hostsim().hostsim_trace, pyr, None, r0 >> (levels - 1), levels, 32, 24, cam, ol.make_opts(10.0))
    assert (rgb == 200).all() and not (hits["flags"] & 1).any()


def test_arithmetic_identities_of_the_production_walk():
    """Random-sample CPU cross-check of the FFMA.RM floor, the exact boundary FMA and the 3-operation division
    (csrc/ray_fast.cuh).  The division is proven exhaustively on the GPU (tests/cuda/divcheck.cu, profiles/divcheck_r01.txt)."""
    import ctypes as C
    import subprocess

    src = ol.REPO / "tests" / "hostsim" / "tricks_check.c"
    so = ol.REPO / "tests" / "_build" / "libtricks_check.so"
    so.parent.mkdir(parents=True, exist_ok=True)
    subprocess.run(["gcc", "-std=c11", "-O1", "-fPIC", "-frounding-math", "-ffp-contract=off", "-shared", "-o", str(so), str(src), "-lm"], check=True)
    lib = C.CDLL(str(so))
    lib.tricks_check.restype = C.c_long
    lib.tricks_check.argtypes = [C.c_uint64, C.c_long]
    assert lib.tricks_check(12345, 20_000_000) == 0
