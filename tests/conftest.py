import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "tests"))
sys.path.insert(0, str(REPO / "gpu-heightmap-raytracer_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_ctx():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import hmrt

    ctx = hmrt.Context(0)
    yield ctx
    ctx.close()
