"""C-ABI collective steps (hmrt_allreduce_max_heights / hmrt_broadcast_heightmap, csrc/comm.cu) through a C++ host harness
(tests/cuda/comm_check.cpp): one context + one NCCL communicator per visible GPU, results bit-exact against host-computed
maxima.  Runs on however many GPUs the box has (a communicator of size 1 on a single-GPU box)."""
import shutil
import subprocess
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = Path(__file__).resolve().parent
REPO = HERE.parent


def test_c_abi_collectives_exact():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    gxx = shutil.which("g++")
    exe = HERE / "_build" / "comm_check"
    csrc = REPO / "gpu-heightmap-raytracer_b200" / "csrc"
    if gxx and Path("/usr/include/nccl.h").exists():
        (HERE / "_build").mkdir(exist_ok=True)
        subprocess.run([gxx, "-std=c++17", "-O1", str(HERE / "cuda" / "comm_check.cpp"), f"-I{REPO / 'include'}", "-I/usr/local/cuda/include",
                        f"-L{csrc}", "-lhmrt", "-L/usr/local/cuda/lib64", "-lcudart", "-lnccl", f"-Wl,-rpath,{csrc}",
                        "-Wl,-rpath,/usr/local/cuda/lib64", "-o", str(exe)], check=True)
    if not exe.exists():
        pytest.skip("no NCCL development files to build the harness")
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "comm_check: ok" in out.stdout
