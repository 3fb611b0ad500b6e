"""The drop-in header EXECUTED: tests/cuda/shim_run.cpp replays the reference's call order (main.cpp:1012-1014, :623-624,
:686, :1022-1024) with the reference's GLM types against host/CudaKernel.cuh on a GPU, and the frames it returns are
compared with the reference's own code (oracle/_ref) -- including a frame after the caller has REFILLED the borrowed
heightmap buffer, as the reference does every frame."""
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

import oraclelib as ol

pytestmark = pytest.mark.gpu
HERE = Path(__file__).resolve().parent
EXE = HERE / "_build" / "shim_run"


def test_dropin_header_runs_the_reference_call_order(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not EXE.exists():
        pytest.skip("tests/_build/shim_run not built (needs the reference's GLM at build time: __graft_entry__.build())")
    if ol.ref() is None:
        pytest.skip("oracle/_ref/libhmrt_ref.so not built")
    sc = ol.scene("r1024_l8", seed=4)
    W, H = 640, 480
    cams = ol.cameras_for(sc, 3)
    for use_color in (0, 1):
        scene = tmp_path / f"scene{use_color}.bin"
        with open(scene, "wb") as f:
            np.array([sc["coarse"], sc["levels"], W, H, len(cams), use_color], np.int32).tofile(f)
            np.array([sc["max_height"]], np.float32).tofile(f)
            for c in cams:
                np.array(list(c.frame_dim) + list(c.forward) + list(c.position), np.float32).tofile(f)
            sc["pyramid"].tofile(f)
            sc["color_map"].tofile(f)
        out = tmp_path / f"frames{use_color}.bin"
        r = subprocess.run([str(EXE), str(scene), str(out)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "shim_run: ok" in r.stdout, r.stdout + r.stderr
        assert "CUDA error" not in r.stderr
        frames = np.fromfile(out, np.uint8).reshape(len(cams), H, W, 3)
        opts = ol.make_opts(sc["max_height"], use_color_map=bool(use_color))
        halved = (sc["pyramid"] * np.float32(0.5)).astype(np.float32)
        for i, cam in enumerate(cams):
            pyr = sc["pyramid"] if i == 0 else halved   # shim_run scales the host copy after the first frame
            want, _ = ol.cpu_trace(ol.ref().hmrt_ref_trace, pyr, sc["color_map"], sc["coarse"], sc["levels"], W, H, cam, opts, want_hits=False)
            assert (frames[i] == want).all(), f"frame {i} (use_color={use_color}) differs from the reference"
