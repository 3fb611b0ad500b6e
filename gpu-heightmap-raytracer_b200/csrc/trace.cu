/*
 * trace.cu -- ray traversal kernels and their C-ABI launchers.
 *
 * Replaces CudaSpace::rayTrace (GPUHeightmapRaytracer/src/CudaKernel.cu:291-308), the
 * <<<1,1>>> parameter kernel cuda_setParameters (:227-240, now evaluated on the host into a
 * 64-byte FrameConsts) and cuda_rayTrace (:195-222).
 *
 * Launch geometry (vs the reference's column-shaped block (1, H/2), :303-305, which is illegal
 * at H = 2160 and stores 3 bytes per thread with a stride of W*3):
 *   CTA  = 256 threads = 32 x 8 pixel tile; each warp owns an 8 x 4 pixel sub-tile so that
 *          the rays of a warp are neighbours in both image directions (coherent walks),
 *   grid = (ceil(W/32), local row tiles, frames): many frames (a fly-through, or a batch of
 *          camera poses) run in ONE launch,
 *   the 8 x 96-byte RGB rows of a tile are staged in shared memory and written with 128-bit
 *   stores (48 x 16 B per CTA) when W % 16 == 0, byte stores otherwise.
 */
#include <string.h>

#include "hmrt_internal.cuh"
#include "ray_fast.cuh"

namespace hmrt {

constexpr int kTileW = 32;
constexpr int kTileH = HMRT_ROW_TILE; /* 8 */
constexpr int kThreads = kTileW * kTileH;

struct TraceParams {
  Grid grid;
  Shading shading;
  FrameConsts frame0;        /* used when frames == nullptr (single-frame launch) */
  const FrameConsts* frames; /* device array, one per blockIdx.z */
  uint8_t* rgb;
  hmrt_hit* hits;
  int W, H;
  int rows_local;
  int tile_first, tile_stride;
  int vec_store; /* W % 16 == 0 and rgb 16-byte aligned */
  const uint32_t* hmax_key; /* order-preserving key of max(top level), written by top_level_max_kernel */
};

/* float <-> unsigned key with the same ordering (so an unsigned atomicMax is a float max) */
__device__ __forceinline__ uint32_t float_to_key(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

/*
 * Maximum of the coarsest pyramid level, recomputed before every trace launch because the
 * heightmap buffer is borrowed and may have been refilled by the caller (the reference re-uploads
 * it every frame, main.cpp:623).  16 K floats for a 16384^2 map: a few microseconds.
 * A NaN cell yields a NaN maximum, which simply disables the air phase.
 */
__global__ void __launch_bounds__(256) top_level_max_kernel(const float* __restrict__ top, uint32_t n, uint32_t* __restrict__ out) {
  uint32_t k = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) k = max(k, float_to_key(__ldg(top + i)));
  k = __reduce_max_sync(0xffffffffu, k);
  if ((threadIdx.x & 31) == 0 && k) atomicMax(out, k);
}

/* WALK: kWalkFast* = production walk (ray_fast.cuh); kWalkReference = operation-by-operation walk
 * (ray_core.cuh), kept as the in-library parity reference (hmrt_set_trace_variant). */
enum Walk { kWalkReference = 0, kWalkFast = 1, kWalkFastPow2 = 2 };

template <bool HITS, int WALK>
__global__ void __launch_bounds__(kThreads) trace_tiles_kernel(const __grid_constant__ TraceParams p) {
  __shared__ __align__(16) uint8_t stage[kTileH][kTileW * 3];
  __shared__ LevelEntry level_tab[HMRT_MAX_LEVELS];
  uint32_t tab = 0; /* shared-window address of the level table, pinned in a register */
  float hmax = 0.0f;
  if (WALK != kWalkReference) {
    hmax = key_to_float(__ldg(p.hmax_key));
    fill_level_table(level_tab, p.grid);
    __syncthreads();
    tab = (uint32_t)__cvta_generic_to_shared(level_tab);
    asm volatile("" : "+r"(tab));
  }

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  /* warp (wx, wy) in a 4 x 2 arrangement; lane (lx, ly) in an 8 x 4 arrangement */
  const int tx = (warp & 3) * 8 + (lane & 7);
  const int ty = (warp >> 2) * 4 + (lane >> 3);
  const int px = blockIdx.x * kTileW + tx;
  const int tile = p.tile_first + blockIdx.y * p.tile_stride;
  const int py = tile * kTileH + ty;                 /* row in the frame */
  const int row_local = blockIdx.y * kTileH + ty;    /* row in this call's output */
  const size_t frame_px = (size_t)p.rows_local * (size_t)p.W;
  const size_t frame_base = (size_t)blockIdx.z * frame_px;

  const bool inside = px < p.W && py < p.H;
  if (inside) {
    FrameConsts f;
    if (p.frames) {
      const float4* src = reinterpret_cast<const float4*>(p.frames + blockIdx.z);
      float4* dst = reinterpret_cast<float4*>(&f);
#pragma unroll
      for (int i = 0; i < 4; ++i) dst[i] = __ldg(src + i);
    } else {
      f = p.frame0;
    }
    RayResult r;
    if (WALK == kWalkFastPow2)
      r = trace_pixel_fast<true>(p.grid, p.shading, tab, hmax, f, p.W, p.H, px, py);
    else if (WALK == kWalkFast)
      r = trace_pixel_fast<false>(p.grid, p.shading, tab, hmax, f, p.W, p.H, px, py);
    else
      r = trace_pixel(p.grid, p.shading, f, p.W, p.H, px, py);
    stage[ty][tx * 3 + 0] = r.r;
    stage[ty][tx * 3 + 1] = r.g;
    stage[ty][tx * 3 + 2] = r.b;
    if (HITS) {
      hmrt_hit h;
      h.x = r.pos.x;
      h.y = r.pos.y;
      h.z = r.pos.z;
      h.flags = r.flags;
      reinterpret_cast<float4*>(p.hits)[frame_base + (size_t)row_local * p.W + px] =
          make_float4(h.x, h.y, h.z, __uint_as_float(h.flags));
    }
  }
  __syncthreads();

  uint8_t* out = p.rgb + frame_base * 3;
  const int x0 = blockIdx.x * kTileW;
  const bool full = (x0 + kTileW <= p.W) && (tile * kTileH + kTileH <= p.H);
  if (p.vec_store && full) {
    /* 8 rows x 6 chunks of 16 B; (x0*3) % 16 == 0 because x0 is a multiple of 32 */
    if (threadIdx.x < kTileH * 6) {
      const int r = threadIdx.x / 6, c = threadIdx.x % 6;
      const uint4 v = *reinterpret_cast<const uint4*>(&stage[r][c * 16]);
      uint8_t* dst = out + ((size_t)(blockIdx.y * kTileH + r) * p.W + x0) * 3 + c * 16;
      *reinterpret_cast<uint4*>(dst) = v;
    }
  } else {
    const int valid_w = min(kTileW, p.W - x0);
    const int valid_h = min(kTileH, p.H - tile * kTileH);
    for (int i = threadIdx.x; i < kTileH * kTileW * 3; i += kThreads) {
      const int r = i / (kTileW * 3), b = i % (kTileW * 3);
      if (r < valid_h && b < valid_w * 3)
        out[((size_t)(blockIdx.y * kTileH + r) * p.W + x0) * 3 + b] = stage[r][b];
    }
  }
}

static int launch_trace(hmrt_ctx* ctx, int W, int H, const hmrt_camera* cams, int n_frames,
                        const hmrt_trace_opts* opts, uint8_t* d_rgb, hmrt_hit* d_hits) {
  if (!ctx || !cams || !opts) return HMRT_E_ARG;
  if (!ctx->have_grid) return HMRT_E_STATE;
  if (W < 2 || H < 2 || n_frames < 1 || n_frames > 65535) return HMRT_E_ARG;
  if (opts->use_color_map && !ctx->grid.color_map) return HMRT_E_ARG;
  if (opts->tile_first < 0 || opts->tile_stride < 0) return HMRT_E_ARG;
  const int stride = opts->tile_stride > 0 ? opts->tile_stride : 1;
  const int n_tiles = (H + kTileH - 1) / kTileH;
  if (opts->tile_first >= n_tiles) return 0; /* nothing to render on this rank */
  if (!d_rgb) return HMRT_E_ARG;
  const int local_tiles = (n_tiles - opts->tile_first + stride - 1) / stride;
  if (local_tiles > 65535) return HMRT_E_SHAPE;

  DeviceGuard guard(ctx->device);
  TraceParams p;
  memset(&p, 0, sizeof(p));
  p.grid = ctx->grid;
  p.shading.max_height = opts->max_height;
  p.shading.use_color_map = opts->use_color_map ? 1 : 0;
  p.shading.shadows = opts->shadows ? 1 : 0;
  for (int i = 0; i < 3; ++i) p.shading.light[i] = opts->light_dir[i];
  p.shading.bias = opts->shadow_bias > 0.f ? opts->shadow_bias : 0.0625f;
  p.rgb = d_rgb;
  p.hits = d_hits;
  p.W = W;
  p.H = H;
  p.rows_local = rows_local(H, opts->tile_first, stride);
  p.tile_first = opts->tile_first;
  p.tile_stride = stride;
  p.vec_store = (W % 16 == 0) && ((reinterpret_cast<uintptr_t>(d_rgb) & 15) == 0);

  if (n_frames == 1) {
    make_frame_consts(cams[0], p.frame0);
    p.frames = nullptr;
  } else {
    if (ctx->frames_cap < n_frames) {
      if (ctx->d_frames) HMRT_CUDA(cudaFree(ctx->d_frames));
      ctx->d_frames = nullptr;
      ctx->frames_cap = 0;
      HMRT_CUDA(cudaMalloc(&ctx->d_frames, sizeof(FrameConsts) * (size_t)n_frames));
      ctx->frames_cap = n_frames;
    }
    /* pageable source: the runtime stages it before returning, so the stack buffer may die */
    FrameConsts local[64];
    for (int base = 0; base < n_frames; base += 64) {
      const int n = n_frames - base < 64 ? n_frames - base : 64;
      for (int i = 0; i < n; ++i) make_frame_consts(cams[base + i], local[i]);
      HMRT_CUDA(cudaMemcpyAsync(ctx->d_frames + base, local, sizeof(FrameConsts) * (size_t)n,
                                cudaMemcpyHostToDevice, ctx->stream));
    }
    p.frames = ctx->d_frames;
  }

  const dim3 grid((W + kTileW - 1) / kTileW, local_tiles, n_frames);
  const bool pow2 = (ctx->grid.coarse_res & (ctx->grid.coarse_res - 1)) == 0;
  const int walk = ctx->trace_variant != 0 ? kWalkReference : (pow2 ? kWalkFastPow2 : kWalkFast);
  if (walk != kWalkReference) {
    if (!ctx->d_hmax) HMRT_CUDA(cudaMalloc(&ctx->d_hmax, sizeof(uint32_t)));
    HMRT_CUDA(cudaMemsetAsync(ctx->d_hmax, 0, sizeof(uint32_t), ctx->stream));
    const uint32_t n_top = ctx->grid.coarse_sq; /* the coarsest level sits at float offset 0 */
    const unsigned blocks = (unsigned)((n_top + 4095u) / 4096u);
    top_level_max_kernel<<<blocks > 1184u ? 1184u : blocks, 256, 0, ctx->stream>>>(ctx->grid.pyramid, n_top, ctx->d_hmax);
    HMRT_LAUNCHED(ctx);
    p.hmax_key = ctx->d_hmax;
  }
#define HMRT_LAUNCH_TRACE(HITS_, WALK_) trace_tiles_kernel<HITS_, WALK_><<<grid, kThreads, 0, ctx->stream>>>(p)
  if (d_hits) {
    if (walk == kWalkFastPow2) HMRT_LAUNCH_TRACE(true, kWalkFastPow2);
    else if (walk == kWalkFast) HMRT_LAUNCH_TRACE(true, kWalkFast);
    else HMRT_LAUNCH_TRACE(true, kWalkReference);
  } else {
    if (walk == kWalkFastPow2) HMRT_LAUNCH_TRACE(false, kWalkFastPow2);
    else if (walk == kWalkFast) HMRT_LAUNCH_TRACE(false, kWalkFast);
    else HMRT_LAUNCH_TRACE(false, kWalkReference);
  }
#undef HMRT_LAUNCH_TRACE
  HMRT_LAUNCHED(ctx);
  return 0;
}

}  // namespace hmrt

extern "C" {

int hmrt_trace(hmrt_ctx* ctx, int W, int H, const hmrt_camera* h_cameras, int n_frames,
               const hmrt_trace_opts* opts, uint8_t* d_rgb, hmrt_hit* d_hits) {
  return hmrt::launch_trace(ctx, W, H, h_cameras, n_frames, opts, d_rgb, d_hits);
}

int hmrt_trace_host(hmrt_ctx* ctx, int W, int H, const hmrt_camera* h_cameras, int n_frames,
                    const hmrt_trace_opts* opts, uint8_t* h_rgb) {
  if (!ctx || !opts || !h_rgb || W < 2 || H < 2 || n_frames < 1) return HMRT_E_ARG;
  hmrt::DeviceGuard guard(ctx->device);
  const int stride = opts->tile_stride > 0 ? opts->tile_stride : 1;
  const size_t bytes = (size_t)hmrt::rows_local(H, opts->tile_first, stride) * (size_t)W * 3 * (size_t)n_frames;
  if (bytes == 0) return 0;
  if (ctx->fb_cap < bytes) {
    if (ctx->d_fb) HMRT_CUDA(cudaFree(ctx->d_fb));
    ctx->d_fb = nullptr;
    ctx->fb_cap = 0;
    HMRT_CUDA(cudaMalloc(&ctx->d_fb, bytes));
    ctx->fb_cap = bytes;
  }
  int rc = hmrt::launch_trace(ctx, W, H, h_cameras, n_frames, opts, ctx->d_fb, nullptr);
  if (rc) return rc;
  HMRT_CUDA(cudaMemcpyAsync(h_rgb, ctx->d_fb, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  HMRT_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // extern "C"
