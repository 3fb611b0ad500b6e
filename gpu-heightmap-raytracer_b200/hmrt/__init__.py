"""hmrt -- Python binding of the B200-native heightfield ray traversal / rasterisation library.

Thin ctypes layer over include/hmrt.h (csrc/libhmrt.so).  torch is used for device memory,
streams and torch.distributed only; every computation on the path is a hand-written CUDA kernel
inside libhmrt.so, and importing/using this package without that library (or without a GPU)
raises instead of falling back to anything.
"""
from ._abi import (  # noqa: F401
    Camera,
    Color,
    Hit,
    HmrtError,
    LasTransform,
    TraceOpts,
    WindowPlacement,
    WindowSections,
    HMRT_ROW_TILE,
    HIT_HIT,
    HIT_MIRROR_X,
    HIT_MIRROR_Z,
    HIT_SHADOWED,
    HIT_STEPS_SHIFT,
    load,
)
from .context import Context, camera, pyramid_layout, rows_local, trace_opts, window_place  # noqa: F401
from . import las, dist  # noqa: F401
