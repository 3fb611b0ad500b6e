"""Context: one device, one stream -- the Python face of hmrt_ctx (include/hmrt.h).

Method names follow the C ABI, which in turn mirrors the reference's host interface
(CudaSpace::initializeDeviceVariables / rayTrace / freeDeviceVariables, CudaKernel.cuh:49-51,
and the loadLASToSection loop, main.cpp:174-244).  Tensors are torch CUDA tensors used purely as
device buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Sequence

import numpy as np

from . import _abi
from ._abi import Camera, LasTransform, TraceOpts, check


def pyramid_layout(coarse_res: int, levels: int):
    """(res[levels], idx[levels], total floats) -- main.cpp:995-1003 via hmrt_pyramid_layout."""
    lib = _abi.load()
    res = (C.c_int * levels)()
    idx = (C.c_int64 * levels)()
    total = C.c_int64()
    check(lib.hmrt_pyramid_layout(coarse_res, levels, res, idx, C.byref(total)), "hmrt_pyramid_layout")
    return list(res), list(idx), total.value


def rows_local(H: int, tile_first: int = 0, tile_stride: int = 1) -> int:
    r = _abi.load().hmrt_rows_local(H, tile_first, tile_stride)
    if r < 0:
        raise _abi.HmrtError(r, "hmrt_rows_local")
    return r


def camera(position, forward, frame_dim=(32.0, 18.0, 20.0), normalize=True) -> Camera:
    """Camera as the reference holds it: frame_dimension (main.cpp:57), unit camera_forward
    (main.cpp:56), grid_camera_position in finest-cell units."""
    f = np.asarray(forward, dtype=np.float64)
    if normalize:
        f = f / np.linalg.norm(f)
    f = f.astype(np.float32)
    cam = Camera()
    cam.frame_dim[:] = [float(np.float32(v)) for v in frame_dim]
    cam.forward[:] = [float(v) for v in f]
    cam.position[:] = [float(np.float32(v)) for v in position]
    return cam


def trace_opts(max_height: float, use_color_map: bool = False, shadows: bool = False,
               light_dir=(0.3, 0.8, 0.52), shadow_bias: float = 0.0, tile_first: int = 0,
               tile_stride: int = 1, full_frame_output: bool = False) -> TraceOpts:
    o = TraceOpts()
    _abi.load().hmrt_trace_opts_default(C.byref(o), float(np.float32(max_height)))
    o.use_color_map = int(bool(use_color_map))
    o.shadows = int(bool(shadows))
    l = np.asarray(light_dir, dtype=np.float64)
    l = (l / np.linalg.norm(l)).astype(np.float32)
    o.light_dir[:] = [float(v) for v in l]
    o.shadow_bias = float(shadow_bias)
    o.tile_first = int(tile_first)
    o.tile_stride = int(tile_stride)
    o.full_frame_output = int(bool(full_frame_output))
    return o


def _cam_array(cameras) -> "C.Array[Camera]":
    if isinstance(cameras, Camera):
        cameras = [cameras]
    if isinstance(cameras, C.Array):
        return cameras
    cams = list(cameras)
    arr = (Camera * len(cams))()
    for i, c in enumerate(cams):
        C.memmove(C.byref(arr, i * C.sizeof(Camera)), C.byref(c), C.sizeof(Camera))
    return arr


class IpcBuffer:
    """A raw device allocation (own or peer) viewed as a torch tensor through __cuda_array_interface__."""

    def __init__(self, ctx, ptr: int, nbytes: int, owner: bool):
        self.ctx, self.ptr, self.nbytes, self.owner = ctx, ptr, nbytes, owner
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}

    def tensor(self, shape):
        import torch

        t = torch.as_tensor(self, device=torch.device("cuda", self.ctx.device))
        return t.view(*shape)

    def close(self):
        if self.ptr:
            fn = self.ctx.lib.hmrt_ipc_free if self.owner else self.ctx.lib.hmrt_ipc_close
            fn(self.ctx._h, C.c_void_p(self.ptr))
            self.ptr = 0


class Context:
    def __init__(self, device: int | None = None):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("hmrt.Context needs a CUDA device: this path has no CPU fallback")
        self._torch = torch
        self.lib = _abi.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        h = C.c_void_p()
        check(self.lib.hmrt_create(self.device, C.byref(h)), "hmrt_create")
        self._h = h
        self._keep = None  # tensors borrowed by set_heightmap
        self.grid = None   # (coarse_res, levels)

    # -- plumbing --------------------------------------------------------------------------
    def _bind_stream(self):
        s = self._torch.cuda.current_stream(self.device).cuda_stream
        check(self.lib.hmrt_set_stream(self._h, C.c_void_p(s)), "hmrt_set_stream")

    def _dev(self, t, dtype, what):
        torch = self._torch
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.device.index == self.device):
            raise TypeError(f"{what}: expected a CUDA tensor on device {self.device}")
        if t.dtype != dtype or not t.is_contiguous():
            raise TypeError(f"{what}: expected a contiguous {dtype} tensor, got {t.dtype}")
        return C.c_void_p(t.data_ptr())

    def close(self):
        if getattr(self, "_h", None):
            self.lib.hmrt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        check(self.lib.hmrt_synchronize(self._h), "hmrt_synchronize")

    def set_trace_variant(self, variant: int):
        """0 = production kernel, 1 = operation-by-operation walk (diagnostic, must agree bit for bit),
        2 = tolerance mode (air phase in one closed-form step; meets the BASELINE acceptance bars, not bit-identical)."""
        check(self.lib.hmrt_set_trace_variant(self._h, int(variant)), "hmrt_set_trace_variant")

    def set_host_variant(self, variant: int):
        """trace_host: 0 = auto (default), 1 = one launch per frame group, 2 = streamed (per-segment copies); identical results."""
        check(self.lib.hmrt_set_host_variant(self._h, int(variant)), "hmrt_set_host_variant")

    def set_window_variant(self, variant: int):
        """compose_window: 0 = TMA bulk-copy gather (default), 1 = per-thread 128-bit gather; identical results."""
        check(self.lib.hmrt_set_window_variant(self._h, int(variant)), "hmrt_set_window_variant")

    def set_l2_persist(self, first_level: int, hit_ratio: float = 1.0):
        """Experiment: persisting-L2 access policy window over the pyramid levels >= first_level (0 / negative = off)."""
        check(self.lib.hmrt_set_l2_persist(self._h, int(first_level), float(hit_ratio)), "hmrt_set_l2_persist")

    def trace_stats(self, reset: bool = True):
        """{rays, iterations, air_iterations} accumulated by the instrumented kernels (trace(..., hits=...)) since the last reset."""
        out = (C.c_uint64 * 4)()
        self._bind_stream()
        check(self.lib.hmrt_trace_stats(self._h, out, int(bool(reset))), "hmrt_trace_stats")
        return {"rays": int(out[0]), "iterations": int(out[1]), "air_iterations": int(out[2])}

    @property
    def launch_count(self) -> int:
        return int(self.lib.hmrt_launch_count(self._h))

    # -- ray traversal ---------------------------------------------------------------------
    def set_heightmap(self, pyramid, color_map, coarse_res: int, levels: int, max_height: float):
        """== CudaSpace::initializeDeviceVariables (CudaKernel.cuh:50)."""
        torch = self._torch
        _, _, total = pyramid_layout(coarse_res, levels)
        p = self._dev(pyramid, torch.float32, "pyramid")
        if pyramid.numel() < total:
            raise ValueError(f"pyramid has {pyramid.numel()} floats, layout needs {total}")
        cm = None
        if color_map is not None:
            cm = self._dev(color_map, torch.uint8, "color_map")
            r0 = coarse_res << (levels - 1)
            if color_map.numel() < r0 * r0 * 3:
                raise ValueError("color_map too small")
        check(self.lib.hmrt_set_heightmap(self._h, p, cm, coarse_res, levels, float(max_height)), "hmrt_set_heightmap")
        self._keep = (pyramid, color_map)
        self.grid = (coarse_res, levels)

    def clear_heightmap(self):
        """== CudaSpace::freeDeviceVariables (CudaKernel.cuh:51)."""
        check(self.lib.hmrt_clear_heightmap(self._h), "hmrt_clear_heightmap")
        self._keep = None
        self.grid = None

    def trace(self, W: int, H: int, cameras, opts: TraceOpts, out=None, hits=False):
        """== CudaSpace::rayTrace (CudaKernel.cuh:49) for one or many cameras in one launch.
        Returns (rgb uint8 [frames, rows_local, W, 3], hits int32 [frames, rows_local, W, 4] | None);
        asynchronous on torch's current stream."""
        torch = self._torch
        cams = _cam_array(cameras)
        n = len(cams)
        rows = H if opts.full_frame_output else rows_local(H, opts.tile_first, opts.tile_stride)
        dev = torch.device("cuda", self.device)
        if out is None:
            out = torch.empty((n, rows, W, 3), dtype=torch.uint8, device=dev)
        elif out.numel() < n * rows * W * 3:
            raise ValueError("out too small")
        hit_t = None
        if hits is True:
            hit_t = torch.zeros((n, rows, W, 4), dtype=torch.int32, device=dev)
        elif isinstance(hits, torch.Tensor):
            if hits.numel() < n * rows * W * 4:
                raise ValueError("hits too small")
            hit_t = hits
        self._bind_stream()
        check(self.lib.hmrt_trace(self._h, W, H, cams, n, C.byref(opts), self._dev(out, torch.uint8, "out"),
                                  self._dev(hit_t, torch.int32, "hits") if hit_t is not None else None), "hmrt_trace")
        return out, hit_t

    @staticmethod
    def _host_ptr(out_host, need: int) -> int:
        if hasattr(out_host, "data_ptr"):
            if out_host.is_cuda or out_host.numel() < need or not out_host.is_contiguous():
                raise ValueError("out_host must be a contiguous CPU uint8 tensor of sufficient size")
            return out_host.data_ptr()
        if out_host.nbytes < need:
            raise ValueError("out_host too small")
        return out_host.ctypes.data

    def trace_host(self, W: int, H: int, cameras, opts: TraceOpts, out_host):
        """Host-buffer variant (hmrt_trace_host): out_host is a (pinned) CPU uint8 tensor/array."""
        cams = _cam_array(cameras)
        n = len(cams)
        ptr = self._host_ptr(out_host, n * rows_local(H, opts.tile_first, opts.tile_stride) * W * 3)
        self._bind_stream()
        check(self.lib.hmrt_trace_host(self._h, W, H, cams, n, C.byref(opts), C.c_void_p(ptr)), "hmrt_trace_host")
        return out_host

    def trace_host_begin(self, W: int, H: int, cameras, opts: TraceOpts, out_host):
        """hmrt_trace_host_begin: enqueue a host-output call and return; out_host (pinned) must stay valid and untouched until the
        matching trace_host_wait().  Up to two calls may be in flight (double-buffered renderer)."""
        cams = _cam_array(cameras)
        n = len(cams)
        ptr = self._host_ptr(out_host, n * rows_local(H, opts.tile_first, opts.tile_stride) * W * 3)
        self._bind_stream()
        check(self.lib.hmrt_trace_host_begin(self._h, W, H, cams, n, C.byref(opts), C.c_void_p(ptr)), "hmrt_trace_host_begin")

    def trace_host_wait(self):
        """hmrt_trace_host_wait: the oldest call begun and not yet waited for has delivered all its frames."""
        check(self.lib.hmrt_trace_host_wait(self._h), "hmrt_trace_host_wait")

    def copy_tiles_to_frames(self, tiles, frames, W: int, H: int, n_frames: int, tile_first: int, tile_stride: int):
        """Compact row-tile output of a sharded trace -> its place in whole frames (possibly peer memory), 2-D device copies."""
        torch = self._torch
        self._bind_stream()
        check(self.lib.hmrt_copy_tiles_to_frames(self._h, self._dev(tiles, torch.uint8, "tiles"), C.c_void_p(frames.data_ptr()), W, H, n_frames,
                                                 tile_first, tile_stride), "hmrt_copy_tiles_to_frames")

    # -- buffers other ranks can map (cudaIpc) -------------------------------------------------
    def ipc_alloc(self, nbytes: int):
        """(IpcBuffer, 64-byte handle): a device buffer of this context that other processes on the box can open."""
        ptr, handle = C.c_void_p(), (C.c_uint8 * 64)()
        check(self.lib.hmrt_ipc_alloc(self._h, nbytes, C.byref(ptr), handle), "hmrt_ipc_alloc")
        return IpcBuffer(self, ptr.value, nbytes, owner=True), bytes(handle)

    def ipc_open(self, handle: bytes, nbytes: int):
        """Map a buffer another process allocated with ipc_alloc (peer memory over NVLink)."""
        ptr = C.c_void_p()
        check(self.lib.hmrt_ipc_open(self._h, handle, C.byref(ptr)), "hmrt_ipc_open")
        return IpcBuffer(self, ptr.value, nbytes, owner=False)

    # -- rasterisation ---------------------------------------------------------------------
    def clear_section(self, pyramid, coarse_res: int, levels: int, color_keys=None, color_map=None):
        torch = self._torch
        self._bind_stream()
        check(self.lib.hmrt_clear_section(
            self._h, self._dev(pyramid, torch.float32, "pyramid"), coarse_res, levels,
            self._dev(color_keys, torch.int64, "color_keys") if color_keys is not None else None,
            self._dev(color_map, torch.uint8, "color_map") if color_map is not None else None), "hmrt_clear_section")

    def scatter_las(self, records, n: int, record_len: int, point_format: int, xf: LasTransform, pyramid,
                    coarse_res: int, levels: int, first_index: int = 0, color_keys=None):
        torch = self._torch
        self._bind_stream()
        if n > 0 and records.numel() < n * record_len:
            raise ValueError("records tensor too small")
        check(self.lib.hmrt_scatter_las(
            self._h, self._dev(records, torch.uint8, "records") if n > 0 else None, n, record_len, point_format,
            C.byref(xf), first_index, self._dev(pyramid, torch.float32, "pyramid"), coarse_res, levels,
            self._dev(color_keys, torch.int64, "color_keys") if color_keys is not None else None), "hmrt_scatter_las")

    def set_scatter_mode(self, mode: int):
        """0 = auto (locality probe), 1 = direct atomics, 2 = tile-binned; results are identical."""
        check(self.lib.hmrt_set_scatter_mode(self._h, int(mode)), "hmrt_set_scatter_mode")

    def scatter_xyz(self, xyz, n: int, xf: LasTransform, pyramid, coarse_res: int, levels: int):
        torch = self._torch
        self._bind_stream()
        if n > 0 and xyz.numel() < 3 * n:
            raise ValueError("xyz tensor too small")
        check(self.lib.hmrt_scatter_xyz(self._h, self._dev(xyz, torch.float32, "xyz") if n > 0 else None, n, C.byref(xf),
                                        self._dev(pyramid, torch.float32, "pyramid"), coarse_res, levels), "hmrt_scatter_xyz")

    def build_mips(self, pyramid, coarse_res: int, levels: int):
        self._bind_stream()
        check(self.lib.hmrt_build_mips(self._h, self._dev(pyramid, self._torch.float32, "pyramid"), coarse_res, levels),
              "hmrt_build_mips")

    def resolve_colors(self, color_keys, color_map, n_cells: int):
        torch = self._torch
        self._bind_stream()
        check(self.lib.hmrt_resolve_colors(self._h, self._dev(color_keys, torch.int64, "color_keys"),
                                           self._dev(color_map, torch.uint8, "color_map"), n_cells), "hmrt_resolve_colors")

    # -- camera window over resident sections ------------------------------------------------
    def compose_window(self, pyramids, color_maps, coarse_res: int, levels: int, cell_x: int, cell_y: int, out_pyramid,
                       out_color_map=None):
        """== the copy loops of preparePointBuffer + copyPointBuffer (main.cpp:516-625) on the device.
        pyramids / color_maps: [[left-bottom, left-top], [right-bottom, right-top]] CUDA tensors (may alias)."""
        torch = self._torch
        secs = _abi.WindowSections()
        for a in range(2):
            for b in range(2):
                secs.d_pyramid[a][b] = self._dev(pyramids[a][b], torch.float32, "section pyramid").value
                if out_color_map is not None:
                    secs.d_color_map[a][b] = self._dev(color_maps[a][b], torch.uint8, "section colour map").value
        self._bind_stream()
        check(self.lib.hmrt_compose_window(
            self._h, C.byref(secs), coarse_res, levels, int(cell_x), int(cell_y), self._dev(out_pyramid, torch.float32, "window pyramid"),
            self._dev(out_color_map, torch.uint8, "window colour map") if out_color_map is not None else None), "hmrt_compose_window")


def window_place(camera_position, section_origins, grid: int, coarse_res: int, levels: int) -> "_abi.WindowPlacement":
    """preparePointBuffer's host arithmetic (main.cpp:461-516); section_origins[i][j] = (x, y) like point_sections_origins."""
    import numpy as np

    cam = (C.c_float * 3)(*[float(v) for v in camera_position])
    org = np.ascontiguousarray(np.asarray(section_origins, dtype=np.float32).reshape(grid, grid, 2))
    out = _abi.WindowPlacement()
    check(_abi.load().hmrt_window_place(cam, org.ctypes.data_as(C.POINTER(C.c_float)), grid, coarse_res, levels, C.byref(out)),
          "hmrt_window_place")
    return out
