"""The parity tests once more against libhmrt_checked.so: the same sources compiled with -DHMRT_CHECKED, where every computed
index of the rasterisation, window and traversal kernels is tested on the device before it is used (HMRT_DCHECK,
csrc/hmrt_internal.cuh; a failure prints the expression and traps, which fails the test that launched the kernel).
compute-sanitizer is not available on the GPU pool; this is the bounds evidence for the hand-written index arithmetic."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
CHECKED = REPO / "gpu-heightmap-raytracer_b200" / "csrc" / "libhmrt_checked.so"

pytestmark = pytest.mark.gpu


def test_parity_suites_pass_with_device_side_bounds_checks():
    if not CHECKED.exists():
        pytest.fail(f"{CHECKED} is missing: run __graft_entry__.build() (make -C gpu-heightmap-raytracer_b200/csrc libhmrt_checked.so)")
    env = dict(os.environ, HMRT_LIB=str(CHECKED))
    suites = ["tests/test_gpu_rx.py", "tests/test_gpu_raster.py", "tests/test_gpu_window.py", "tests/test_gpu_edge.py", "tests/test_gpu_trace.py"]
    # the whole-frame 4K comparisons and the exhaustive arithmetic enumerations do not touch other index arithmetic
    cmd = [sys.executable, "-m", "pytest", *suites, "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider", "-k", "not full_size and not exhaustive"]
    r = subprocess.run(cmd, cwd=REPO, env=env, capture_output=True, text=True, timeout=1500)
    tail = (r.stdout[-3000:] + "\n" + r.stderr[-2000:])
    assert r.returncode == 0, tail
    assert "HMRT_DCHECK failed" not in r.stdout and "HMRT_DCHECK failed" not in r.stderr, tail
    assert " passed" in r.stdout, tail


def test_a_failing_check_is_reported():
    """The mechanism itself: the checked library says so, a check that fails on purpose prints its expression and surfaces as a CUDA
    error; the product library carries no checks."""
    code = (
        "import ctypes as C, sys\n"
        "sys.path.insert(0, 'gpu-heightmap-raytracer_b200')\n"
        "import hmrt\n"
        "ctx = hmrt.Context(0)\n"
        "print('checked', ctx.lib.hmrt_debug_checked_build())\n"
        "ctx.lib.hmrt_debug_dcheck_selftest.argtypes = [C.c_void_p]\n"
        "print('selftest', ctx.lib.hmrt_debug_dcheck_selftest(ctx._h))\n"
        "sys.stdout.flush()\n"
        "import os; os._exit(0)\n"
    )
    for lib, checked in ((CHECKED, True), (CHECKED.with_name("libhmrt.so"), False)):
        r = subprocess.run([sys.executable, "-c", code], cwd=REPO, env=dict(os.environ, HMRT_LIB=str(lib)), capture_output=True, text=True, timeout=300)
        out = r.stdout + r.stderr
        assert f"checked {1 if checked else 0}" in out, out
        if checked:
            assert "HMRT_DCHECK failed: v == 0" in out and "selftest 0" not in out, out
        else:
            assert "selftest 0" in out and "HMRT_DCHECK failed" not in out, out
