/*
 * api.cu -- context management and the small host-side pieces of the C ABI (include/hmrt.h).
 * Reference counterparts: initializeDeviceVariables / freeDeviceVariables
 * (GPUHeightmapRaytracer/src/CudaKernel.cu:313-326, :245-286) and the pyramid tables
 * (main.cpp:995-1003).
 */
#include <new>
#include <stdio.h>
#include <string.h>

#include "hmrt_internal.cuh"

namespace hmrt {

/* main.cpp:995-1003 == CudaKernel.cu:250-258 */
int pyramid_layout(int coarse_res, int levels, int* res, int64_t* idx, int64_t* total) {
  if (coarse_res < 1 || levels < 1 || levels > HMRT_MAX_LEVELS) return HMRT_E_ARG;
  int r[HMRT_MAX_LEVELS];
  int64_t ix[HMRT_MAX_LEVELS];
  r[levels - 1] = coarse_res;
  ix[levels - 1] = 0;
  for (int i = levels - 2; i >= 0; i--) {
    ix[i] = ix[i + 1] + (int64_t)r[i + 1] * r[i + 1];
    if (r[i + 1] > (1 << 29)) return HMRT_E_SHAPE;
    r[i] = r[i + 1] * 2;
  }
  const int64_t n = ix[0] + (int64_t)r[0] * r[0];
  /* cell indices are 32-bit in the kernels, coordinates must be exact in fp32 */
  if (n >= ((int64_t)1 << 32) || r[0] > (1 << 23)) return HMRT_E_SHAPE;
  for (int i = 0; i < levels; i++) {
    if (res) res[i] = r[i];
    if (idx) idx[i] = ix[i];
  }
  if (total) *total = n;
  return 0;
}

}  // namespace hmrt

extern "C" {

int hmrt_version(void) { return HMRT_VERSION; }

const char* hmrt_error_string(int code) {
  switch (code) {
    case 0: return "success";
    case HMRT_E_ARG: return "invalid argument";
    case HMRT_E_STATE: return "invalid call order (no heightmap set)";
    case HMRT_E_SHAPE: return "unsupported grid shape";
    case HMRT_E_NOMEM: return "host allocation failed";
    case HMRT_E_NCCL: return "NCCL unavailable or a NCCL call failed";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown error";
}

int hmrt_create(int device, hmrt_ctx** out) {
  if (!out) return HMRT_E_ARG;
  *out = nullptr;
  int n = 0;
  HMRT_CUDA(cudaGetDeviceCount(&n));
  if (device < 0 || device >= n) return HMRT_E_ARG;
  hmrt::DeviceGuard guard(device);
  HMRT_CUDA(cudaFree(0)); /* create the primary context now so later errors are attributable */
  hmrt_ctx* c = new (std::nothrow) hmrt_ctx();
  if (!c) return HMRT_E_NOMEM;
  memset(c, 0, sizeof(*c));
  c->device = device;
  c->probe_verdict = -1;
  c->l2_first_level = -1;
  cudaError_t e = cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (e != cudaSuccess) {
    delete c;
    return (int)e;
  }
  *out = c;
  return 0;
}

int hmrt_destroy(hmrt_ctx* ctx) {
  if (!ctx) return HMRT_E_ARG;
  hmrt::DeviceGuard guard(ctx->device);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (ctx->d_frames) cudaFree(ctx->d_frames);
  if (ctx->d_fb) cudaFree(ctx->d_fb);
  if (ctx->d_scratch) cudaFree(ctx->d_scratch);
  if (ctx->d_stats) cudaFree(ctx->d_stats);
  if (ctx->d_seg_done) cudaFree(ctx->d_seg_done);
  if (ctx->d_ws) cudaFree(ctx->d_ws);
  if (ctx->d_probe) cudaFree(ctx->d_probe);
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    for (int i = 0; i < 2; ++i) {
      cudaStreamSynchronize(ctx->frame_stream[i]);
      cudaEventDestroy(ctx->frame_event[i]);
      cudaStreamDestroy(ctx->frame_stream[i]);
    }
    cudaEventDestroy(ctx->prep_event);
    for (int i = 0; i < 2; ++i) cudaEventDestroy(ctx->host_done[i]);
    cudaStreamDestroy(ctx->copy_stream);
  }
  delete ctx;
  return (int)e;
}

int hmrt_set_stream(hmrt_ctx* ctx, void* cuda_stream) {
  if (!ctx) return HMRT_E_ARG;
  ctx->stream = (cudaStream_t)cuda_stream;
  return 0;
}

void* hmrt_get_stream(const hmrt_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int hmrt_synchronize(hmrt_ctx* ctx) {
  if (!ctx) return HMRT_E_ARG;
  hmrt::DeviceGuard guard(ctx->device);
  HMRT_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int hmrt_pyramid_layout(int coarse_res, int levels, int* res, int64_t* idx, int64_t* total) {
  return hmrt::pyramid_layout(coarse_res, levels, res, idx, total);
}

int hmrt_set_heightmap(hmrt_ctx* ctx, const float* d_pyramid, const hmrt_color* d_color_map, int coarse_res,
                       int levels, float max_height) {
  if (!ctx || !d_pyramid) return HMRT_E_ARG;
  int res[HMRT_MAX_LEVELS];
  int rc = hmrt::pyramid_layout(coarse_res, levels, res, nullptr, nullptr);
  if (rc) return rc;
  hmrt::Grid& g = ctx->grid;
  g.pyramid = d_pyramid;
  g.color_map = reinterpret_cast<const uint8_t*>(d_color_map);
  g.coarse_res = coarse_res;
  g.coarse_sq = (uint32_t)coarse_res * (uint32_t)coarse_res;
  g.levels = levels;
  g.res0 = res[0];
  /* point_buffer_resolution->x * pow(2.f, LOD_levels - 1) (CudaKernel.cu:134): exact, < 2^24 */
  g.extent = (float)coarse_res * (float)(1 << (levels - 1));
  ctx->init_max_height = max_height;
  ctx->have_grid = true;
  return 0;
}

int hmrt_clear_heightmap(hmrt_ctx* ctx) {
  if (!ctx) return HMRT_E_ARG;
  ctx->have_grid = false;
  memset(&ctx->grid, 0, sizeof(ctx->grid));
  return 0;
}

void hmrt_trace_opts_default(hmrt_trace_opts* o, float max_height) {
  if (!o) return;
  memset(o, 0, sizeof(*o));
  o->max_height = max_height;
  o->light_dir[0] = 0.0f;
  o->light_dir[1] = 1.0f;
  o->light_dir[2] = 0.0f;
  o->shadow_bias = 0.0625f;
  o->tile_first = 0;
  o->tile_stride = 1;
}

int hmrt_rows_local(int H, int tile_first, int tile_stride) {
  if (H < 1 || tile_first < 0) return HMRT_E_ARG;
  return hmrt::rows_local(H, tile_first, tile_stride);
}

int hmrt_set_trace_variant(hmrt_ctx* ctx, int variant) {
  if (!ctx || variant < 0 || variant > 2) return HMRT_E_ARG;
  ctx->trace_variant = variant;
  return 0;
}

int64_t hmrt_launch_count(const hmrt_ctx* ctx) { return ctx ? ctx->launches : -1; }

int hmrt_set_host_variant(hmrt_ctx* ctx, int variant) {
  if (!ctx || variant < 0 || variant > 2) return HMRT_E_ARG;
  ctx->host_variant = variant;
  return 0;
}

/* development knobs (not in include/hmrt.h): row segments per frame of the streamed hmrt_trace_host schedule; frame groups per
 * middle launch of the per-group schedule.  0 = built-in. */
int hmrt_debug_host_segments(hmrt_ctx* ctx, int segments) {
  if (!ctx || segments < 0 || segments > 4) return HMRT_E_ARG;
  ctx->host_segments = segments;
  return 0;
}
int hmrt_debug_host_mid_groups(hmrt_ctx* ctx, int groups) {
  if (!ctx || groups < 0 || groups > 64) return HMRT_E_ARG;
  ctx->host_mid_groups = groups;
  return 0;
}

/* Checked build (make libhmrt_checked.so, HMRT_DCHECK in hmrt_internal.cuh): 1 when this library carries the device-side bounds
 * checks; the self-test launches a kernel whose check fails on purpose and returns the CUDA error the trap produces (the
 * context is unusable afterwards: call it from a throw-away process).  Not part of include/hmrt.h. */
__global__ void dcheck_selftest_kernel(int v) { HMRT_DCHECK(v == 0); }
int hmrt_debug_checked_build(void) {
#if defined(HMRT_CHECKED)
  return 1;
#else
  return 0;
#endif
}
int hmrt_debug_dcheck_selftest(hmrt_ctx* ctx) {
  if (!ctx) return HMRT_E_ARG;
  hmrt::DeviceGuard guard(ctx->device);
  dcheck_selftest_kernel<<<1, 1, 0, ctx->stream>>>(1);
  return (int)cudaStreamSynchronize(ctx->stream);
}

int hmrt_set_window_variant(hmrt_ctx* ctx, int variant) {
  if (!ctx || variant < 0 || variant > 1) return HMRT_E_ARG;
  ctx->window_variant = variant;
  return 0;
}

int hmrt_set_l2_persist(hmrt_ctx* ctx, int first_level, float hit_ratio) {
  if (!ctx || first_level > HMRT_MAX_LEVELS || !(hit_ratio >= 0.0f && hit_ratio <= 1.0f)) return HMRT_E_ARG;
  hmrt::DeviceGuard guard(ctx->device);
  if (first_level < 1) {
    ctx->l2_first_level = -1;
    HMRT_CUDA(cudaCtxResetPersistingL2Cache());
    return 0;
  }
  int max_persist = 0;
  HMRT_CUDA(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device));
  if (max_persist <= 0) return HMRT_E_STATE;
  HMRT_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist));
  ctx->l2_first_level = first_level;
  ctx->l2_hit_ratio = hit_ratio;
  return 0;
}

int hmrt_copy_tiles_to_frames(hmrt_ctx* ctx, const uint8_t* d_tiles, uint8_t* d_frames, int W, int H, int n_frames, int tile_first,
                              int tile_stride) {
  if (!ctx || !d_tiles || !d_frames || W < 1 || H < 1 || n_frames < 1 || tile_first < 0) return HMRT_E_ARG;
  const int stride = tile_stride > 0 ? tile_stride : 1;
  const int n_tiles = (H + HMRT_ROW_TILE - 1) / HMRT_ROW_TILE;
  if (tile_first >= n_tiles) return 0;
  hmrt::DeviceGuard guard(ctx->device);
  const int local_tiles = (n_tiles - tile_first + stride - 1) / stride;
  const size_t row_bytes = (size_t)W * 3, tile_bytes = row_bytes * HMRT_ROW_TILE;
  const size_t rows_local = (size_t)hmrt::rows_local(H, tile_first, stride);
  /* does this shard own the frame's ragged last tile? */
  const bool owns_last = (n_tiles - 1 - tile_first) % stride == 0 && H % HMRT_ROW_TILE != 0;
  const int whole = owns_last ? local_tiles - 1 : local_tiles;
  for (int f = 0; f < n_frames; ++f) {
    const uint8_t* src = d_tiles + (size_t)f * rows_local * row_bytes;
    uint8_t* dst = d_frames + (size_t)f * (size_t)H * row_bytes + (size_t)tile_first * tile_bytes;
    /* local tile j -> frame tile tile_first + j * stride: a 2-D copy, one "row" per tile (copy engine, no SM time) */
    if (whole > 0)
      HMRT_CUDA(cudaMemcpy2DAsync(dst, tile_bytes * (size_t)stride, src, tile_bytes, tile_bytes, (size_t)whole, cudaMemcpyDeviceToDevice, ctx->stream));
    if (owns_last) {
      const size_t last_rows = (size_t)(H % HMRT_ROW_TILE);
      HMRT_CUDA(cudaMemcpyAsync(dst + (size_t)whole * tile_bytes * (size_t)stride, src + (size_t)whole * tile_bytes, last_rows * row_bytes,
                                cudaMemcpyDeviceToDevice, ctx->stream));
    }
  }
  return 0;
}

int hmrt_ipc_alloc(hmrt_ctx* ctx, size_t bytes, void** d_ptr, void* handle64) {
  if (!ctx || !d_ptr || !handle64 || bytes == 0) return HMRT_E_ARG;
  hmrt::DeviceGuard guard(ctx->device);
  *d_ptr = nullptr;
  void* p = nullptr;
  HMRT_CUDA(cudaMalloc(&p, bytes)); /* cudaMalloc, not a pooled allocation: only whole allocations can be exported */
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return (int)e;
  }
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  *d_ptr = p;
  return 0;
}

int hmrt_ipc_free(hmrt_ctx* ctx, void* d_ptr) {
  if (!ctx || !d_ptr) return HMRT_E_ARG;
  hmrt::DeviceGuard guard(ctx->device);
  HMRT_CUDA(cudaFree(d_ptr));
  return 0;
}

int hmrt_ipc_open(hmrt_ctx* ctx, const void* handle64, void** d_ptr) {
  if (!ctx || !handle64 || !d_ptr) return HMRT_E_ARG;
  hmrt::DeviceGuard guard(ctx->device);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  *d_ptr = nullptr;
  HMRT_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int hmrt_ipc_close(hmrt_ctx* ctx, void* d_ptr) {
  if (!ctx || !d_ptr) return HMRT_E_ARG;
  hmrt::DeviceGuard guard(ctx->device);
  HMRT_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return 0;
}

}  // extern "C"
