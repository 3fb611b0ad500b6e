/*
 * trace.cu -- ray traversal kernels and their C-ABI launchers.
 *
 * Replaces CudaSpace::rayTrace (GPUHeightmapRaytracer/src/CudaKernel.cu:291-308), the
 * <<<1,1>>> parameter kernel cuda_setParameters (:227-240, now evaluated on the host into a
 * 64-byte FrameConsts) and cuda_rayTrace (:195-222).
 *
 * Launch geometry (vs the reference's column-shaped block (1, H/2), :303-305, which is illegal
 * at H = 2160 and stores 3 bytes per thread with a stride of W*3): a persistent grid of
 * (SMs x resident CTAs) CTAs of 8 warps.  Every WARP repeatedly claims the next 32 x 4 pixel chunk
 * of any frame of the launch from a global counter and walks it as four 8 x 4 pixel tiles, so the
 * 32 rays in flight are neighbours in both image directions (coherent walks).  The chunk's
 * 4 x 96 bytes of RGB8 are staged in the warp's private slice of shared memory and written with 24
 * coalesced 128-bit stores (byte stores when W % 16 != 0 or on ragged edges).  Many frames
 * (a fly-through, a batch of camera poses) run in ONE launch.
 */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "hmrt_internal.cuh"
#include "ray_fast.cuh"

namespace hmrt {

constexpr int kThreads = 256; /* 8 warps per CTA */
constexpr int kWarps = kThreads / 32;
constexpr int kChunkW = 32, kChunkH = 4; /* a warp's unit of work: four 8 x 4 pixel tiles side by side */
constexpr int kStageRow = kChunkW * 3;   /* 96 bytes of RGB8 per chunk row */
static_assert(HMRT_ROW_TILE == 2 * kChunkH, "a row tile is two chunk rows");

constexpr int kInlineFrames = 48;        /* frames whose constants fit in the kernel parameters */
/* Calls rotate through this many independent sets of scratch slots / frame-constant slots, so that up to kCallSets trace
 * calls issued on DIFFERENT streams (hmrt_set_stream between calls) may be in flight at once: the drain of one call's
 * persistent kernel then runs under the head of the next one (bench.py alternates two streams). */
constexpr int kCallSets = 4;
constexpr uint32_t kPrepTopMax = 65536;  /* top levels up to this many cells are reduced by the one-CTA prep kernel */

/* device-side scratch refreshed by the launcher (16 bytes) */
struct TraceScratch {
  uint32_t hmax_key;             /* order-preserving key of max(top level), see top_level_max_kernel */
  uint32_t next_tail;            /* work counter of the tile-granular tail */
  uint32_t next_chunk;           /* work counter of the persistent kernel (whole chunks) */
  uint32_t pad;
};

constexpr int kMaxSegments = 4; /* row segments per frame whose completion can be signalled separately (TraceParams::segments of them) */
struct SegDone {
  uint32_t units; /* 8 x 4 tiles of the segment stored so far */
  uint32_t flag;  /* 1 once all of them are (polled by cuStreamWaitValue32 on the copy stream) */
};

struct TraceParams {
  Grid grid;
  Shading shading;
  const FrameConsts* frames; /* device array, one per frame; nullptr: the frames travel in frame_inline */
  uint8_t* rgb;
  hmrt_hit* hits;
  int W, H;
  int rows_local;   /* rows per frame of the OUTPUT: the rows this call renders (compact), or H (full_layout) */
  int full_layout;
  int tile_first, tile_stride;
  int vec_store; /* W % 16 == 0 and rgb 16-byte aligned */
  PixelGrid pixel_grid; /* (W - 1), (H - 1), their reciprocals; whether the frames allow primary_ray_fast */
  const uint32_t* hmax_key;          /* &scratch[0].hmax_key */
  uint32_t* next_chunk;              /* this launch's work counter (whole chunks) */
  uint32_t* next_tail;               /* ... and the counter of the tile-granular tail */
  uint32_t bulk_chunks;              /* chunks [0, bulk_chunks) are claimed whole, the rest tile by tile */
  uint32_t tail_tiles;               /* 4 * (total_chunks - bulk_chunks) */
  uint32_t chunks_x;               /* ceil(W / 32) */
  uint32_t chunks_per_frame;       /* local row tiles * 2 * chunks_x */
  uint32_t total_chunks;           /* frames * chunks_per_frame (< 2^31: larger calls are split by the launcher) */
  unsigned long long* stats;       /* instrumented kernels only: {rays, loop iterations, air-phase iterations} accumulated */
  /* hmrt_trace_host: completion tracking per (frame, row segment) so that the device->host copy of a segment can start the
   * moment its last tile is stored, while the same launch is still tracing the rest (null: no tracking) */
  SegDone* seg_done;               /* [frames * segments] */
  uint32_t segments;               /* row segments per frame (1 .. kMaxSegments) */
  uint32_t strips_per_frame;       /* chunks_per_frame / chunks_x */
  /* up to kInlineFrames per-frame constants ride in the kernel parameters: no host->device copy in front
   * of the launch (a 16-frame call spent ~27 us of device timeline on that copy) */
  alignas(16) FrameConsts frame_inline[kInlineFrames];
};
static_assert(sizeof(TraceParams) <= 4096, "kernel parameter space");

/* float <-> unsigned key with the same ordering (so an unsigned atomicMax is a float max) */
__device__ __forceinline__ uint32_t float_to_key(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

/*
 * Maximum of the coarsest pyramid level, recomputed before every trace call because the
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

/*
 * One-CTA preparation kernel in front of every trace call: zeroes the work counters of the call's launches and,
 * for top levels of up to kPrepTopMax cells (256^2: every BASELINE configuration), reduces the top-level maximum
 * itself -- one small launch instead of a memset, a reduction launch and (for multi-frame calls) a parameter copy.
 */
__global__ void __launch_bounds__(1024) trace_prep_kernel(const float* __restrict__ top, uint32_t n_top, TraceScratch* scratch, int slots,
                                                          int reduce_here) {
  __shared__ uint32_t warp_max[32];
  for (int s = threadIdx.x; s < slots; s += blockDim.x) {
    scratch[s].next_chunk = 0u;
    scratch[s].next_tail = 0u;
    if (s > 0 || !reduce_here) scratch[s].hmax_key = 0u; /* slot 0: written below, or by top_level_max_kernel */
  }
  if (!reduce_here) return;
  uint32_t k = 0;
  if ((reinterpret_cast<uintptr_t>(top) & 15) == 0) { /* 16-byte loads: 4 per thread for a 128^2 top level */
    const uint32_t n4 = n_top >> 2;
    for (uint32_t i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(top) + i);
      k = max(max(k, float_to_key(v.x)), max(float_to_key(v.y), max(float_to_key(v.z), float_to_key(v.w))));
    }
    for (uint32_t i = (n4 << 2) + threadIdx.x; i < n_top; i += blockDim.x) k = max(k, float_to_key(__ldg(top + i)));
  } else {
    for (uint32_t i = threadIdx.x; i < n_top; i += blockDim.x) k = max(k, float_to_key(__ldg(top + i)));
  }
  k = __reduce_max_sync(0xffffffffu, k);
  if ((threadIdx.x & 31) == 0) warp_max[threadIdx.x >> 5] = k;
  __syncthreads();
  if (threadIdx.x < 32) {
    k = __reduce_max_sync(0xffffffffu, warp_max[threadIdx.x]);
    if (threadIdx.x == 0) scratch[0].hmax_key = k;
  }
}

/* development knobs (benchmarks/shard_probe.py): -1 / 0 = the built-in choice */
static int g_knob_tailed = -1, g_knob_tail_per_warp = 0;

/* WALK: kWalkFast* = production walk (ray_fast.cuh); kWalkReference = operation-by-operation walk
 * (ray_core.cuh), kept as the in-library parity reference (hmrt_set_trace_variant). */
enum Walk { kWalkReference = 0, kWalkFast = 1, kWalkFastPow2 = 2, kWalkJumpPow2 = 3 /* tolerance mode, see ray_fast.cuh */ };

/* One unit of work of a warp: the tiles [k_first, k_last) of chunk `chunk` (all four in the bulk phase -> 128-bit
 * stores; a single one in the tail phase -> 8-byte stores). */
template <bool HITS, int WALK, bool TAIL, bool NOTIFY>
__device__ __forceinline__ void trace_unit(const TraceParams& p, uint32_t tab, float hmax, FrameConsts* frame_slot,
                                           uint8_t (*stage)[kStageRow], uint32_t chunk, int k_first, int k_last) {
  const int lane = threadIdx.x & 31;
  const int lx = lane & 7, ly = lane >> 3; /* 8 x 4 lanes */
  /* chunk -> (frame, local row tile j, 4-row half, 32-pixel column group); consecutive chunks are
   * neighbours along x, so concurrently running warps cover a compact band of the frame */
  const uint32_t frame = chunk / p.chunks_per_frame;
  const uint32_t in_frame = chunk - frame * p.chunks_per_frame;
  /* rows are visited from the top of the frame down: the far / near-horizon rows hold the longest walks, so they
   * are claimed first and the launch drains on the short, uniform walks of the bottom rows (longest-processing-
   * time-first; measured: the drain of a 1/8-frame-share launch cost ~5 % with bottom-up order) */
  const uint32_t strip_up = in_frame / p.chunks_x, cx = in_frame - strip_up * p.chunks_x;
  const uint32_t strip = p.chunks_per_frame / p.chunks_x - 1u - strip_up;
  const int j = (int)(strip >> 1), half = (int)(strip & 1u);
  const int tile = p.tile_first + j * p.tile_stride;         /* 8-row tile in the frame */
  const int row0 = tile * HMRT_ROW_TILE + half * kChunkH;    /* first row of the chunk in the frame */
  /* ... in this call's output: compact (local tile j at rows [8 j, 8 j + 8)), or -- full_layout -- at its place in the frame */
  const int row0_local = p.full_layout ? row0 : j * HMRT_ROW_TILE + half * kChunkH;
  const int py = row0 + ly;
  const int x0 = (int)cx * kChunkW;
  const size_t frame_base = (size_t)frame * ((size_t)p.rows_local * (size_t)p.W);

  /* frame constants -> the warp's shared slot (read back at the start of every ray; keeping the 16
   * floats in registers across the walk costs a whole CTA of occupancy) */
  if (lane < 4) {
    const float4* src = reinterpret_cast<const float4*>(p.frames ? p.frames + frame : &p.frame_inline[frame]);
    reinterpret_cast<float4*>(frame_slot)[lane] = src[lane];
  }
  __syncwarp();
  const FrameConsts& f = *frame_slot;

  uint32_t st_rays = 0, st_steps = 0, st_air = 0;
#pragma unroll 1
  for (int k = k_first; k < k_last; ++k) { /* the 8 x 4 tiles of the unit */
    const int tx = k * 8 + lx;
    const int px = x0 + tx;
    if (px < p.W && py < p.H) {
      RayResult r;
      if (WALK == kWalkJumpPow2)
        r = trace_pixel_fast<true, true>(p.grid, p.shading, tab, hmax, f, p.pixel_grid, p.W, p.H, px, py);
      else if (WALK == kWalkFastPow2)
        r = trace_pixel_fast<true, false>(p.grid, p.shading, tab, hmax, f, p.pixel_grid, p.W, p.H, px, py);
      else if (WALK == kWalkFast)
        r = trace_pixel_fast<false, false>(p.grid, p.shading, tab, hmax, f, p.pixel_grid, p.W, p.H, px, py);
      else
        r = trace_pixel(p.grid, p.shading, f, p.W, p.H, px, py);
      stage[ly][tx * 3 + 0] = r.r;
      stage[ly][tx * 3 + 1] = r.g;
      stage[ly][tx * 3 + 2] = r.b;
      if (HITS) {
        reinterpret_cast<float4*>(p.hits)[frame_base + (size_t)(row0_local + ly) * p.W + px] =
            make_float4(r.pos.x, r.pos.y, r.pos.z, __uint_as_float(r.flags));
        st_rays += 1u, st_steps += r.flags >> HMRT_HIT_STEPS_SHIFT, st_air += r.air;
      }
    }
  }
  __syncwarp();
  if (HITS && p.stats) { /* a unit is at most 128 rays of < 2^24 iterations each: the 32-bit warp sums cannot overflow in practice */
    st_rays = __reduce_add_sync(0xffffffffu, st_rays);
    st_steps = __reduce_add_sync(0xffffffffu, st_steps);
    st_air = __reduce_add_sync(0xffffffffu, st_air);
    if (lane == 0) {
      atomicAdd(p.stats + 0, (unsigned long long)st_rays);
      atomicAdd(p.stats + 1, (unsigned long long)st_steps);
      atomicAdd(p.stats + 2, (unsigned long long)st_air);
    }
  }

  uint8_t* out = p.rgb + frame_base * 3;
  const int xa = x0 + k_first * 8, xb = min(x0 + k_last * 8, p.W); /* pixel columns produced by this unit */
  const bool full = (x0 + k_last * 8 <= p.W) && (row0 + kChunkH <= p.H);
  if (!TAIL && p.vec_store && full) {
    /* 4 rows x 6 pieces of 16 B; (x0 * 3) % 16 == 0 because x0 is a multiple of 32 */
    if (lane < kChunkH * 6) {
      const int r = lane / 6, c = lane % 6;
      const uint4 v = *reinterpret_cast<const uint4*>(&stage[r][c * 16]);
      *reinterpret_cast<uint4*>(out + ((size_t)(row0_local + r) * p.W + x0) * 3 + c * 16) = v;
    }
  } else if (TAIL && p.vec_store && full) {
    /* one 8 x 4 tile: 4 rows x 3 pieces of 8 B (24-byte rows are 8-byte aligned when W % 16 == 0) */
    if (lane < kChunkH * 3) {
      const int r = lane / 3, c = lane % 3;
      const uint2 v = *reinterpret_cast<const uint2*>(&stage[r][k_first * 24 + c * 8]);
      *reinterpret_cast<uint2*>(out + ((size_t)(row0_local + r) * p.W + xa) * 3 + c * 8) = v;
    }
  } else {
    const int valid_h = min(kChunkH, p.H - row0);
    const int b0 = (xa - x0) * 3, b1 = (xb - x0) * 3;
    for (int i = lane; i < kChunkH * kStageRow; i += 32) {
      const int r = i / kStageRow, b = i % kStageRow;
      if (r < valid_h && b >= b0 && b < b1) out[((size_t)(row0_local + r) * p.W + x0) * 3 + b] = stage[r][b];
    }
  }
  __syncwarp(); /* the stage slice is reused by the next unit */
  if (NOTIFY) {
    if (lane == 0) {
      /* everything is recomputed from the chunk number: nothing extra stays live across the walk */
      const uint32_t frame_n = chunk / p.chunks_per_frame;
      const uint32_t strip_n = p.strips_per_frame - 1u - (chunk - frame_n * p.chunks_per_frame) / p.chunks_x;
      const uint32_t seg = strip_n * p.segments / p.strips_per_frame;
      const uint32_t s_lo = (seg * p.strips_per_frame + p.segments - 1) / p.segments, s_hi = ((seg + 1) * p.strips_per_frame + p.segments - 1) / p.segments;
      const uint32_t want = (s_hi - s_lo) * p.chunks_x * 4u, mine = (uint32_t)(k_last - k_first);
      SegDone* sd = p.seg_done + frame_n * p.segments + seg;
      __threadfence(); /* the tile's stores (ordered before this lane by the __syncwarp) before the count */
      if (atomicAdd(&sd->units, mine) + mine == want) {
        __threadfence_system();
        *reinterpret_cast<volatile uint32_t*>(&sd->flag) = 1u;
      }
    }
  }
}

/*
 * Persistent traversal kernel (see the file header).  There is no CTA-wide barrier after the
 * prologue: a warp that finishes a cheap (sky) chunk immediately claims the next one instead of
 * idling at a tile barrier (ncu r01: 17 % of the resident warps were parked at the barrier of the
 * former one-tile-per-CTA kernel).  Bulk phase: whole 32 x 4 chunks.  Tail phase: the last
 * "one chunk per resident warp" is handed out tile by tile (8 x 4), so all warps finish within one
 * tile's time of each other -- that matters for single-frame launches (interactive use, the per-frame
 * launches of hmrt_trace_host, 1/8 frames on 8 GPUs).  TAILED = false compiles the tail out: for launches
 * with hundreds of chunks per warp the tail is irrelevant and the leaner code is ~1.5 % faster (measured).
 * NOTIFY (hmrt_trace_host): every unit reports to its (frame, row segment) counter, see SegDone.
 */
#ifndef HMRT_MIN_CTAS
#define HMRT_MIN_CTAS (TAILED ? 5 : 0)
#endif
template <bool HITS, int WALK, bool TAILED, bool NOTIFY = false>
__global__ void __launch_bounds__(kThreads, HMRT_MIN_CTAS) trace_persistent_kernel(const __grid_constant__ TraceParams p) {
  __shared__ __align__(16) uint8_t stage[kWarps][kChunkH][kStageRow];
  __shared__ LevelEntry level_tab[HMRT_MAX_LEVELS];
  __shared__ __align__(16) FrameConsts frame_s[kWarps]; /* the frame constants of each warp's current unit */
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

  for (;;) { /* bulk */
    uint32_t c = 0;
    if (lane == 0) c = atomicAdd(p.next_chunk, 1u);
    c = __shfl_sync(0xffffffffu, c, 0);
    if (c >= p.bulk_chunks) break;
    trace_unit<HITS, WALK, false, NOTIFY>(p, tab, hmax, &frame_s[warp], stage[warp], c, 0, 4);
  }
  if (TAILED) {
    for (;;) { /* tail */
      uint32_t t = 0;
      if (lane == 0) t = atomicAdd(p.next_tail, 1u);
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t >= p.tail_tiles) break;
      const int k = (int)(t & 3u);
      trace_unit<HITS, WALK, true, NOTIFY>(p, tab, hmax, &frame_s[warp], stage[warp], p.bulk_chunks + (t >> 2), k, k + 1);
    }
  }
}

static const void* pick_kernel(bool hits, int walk, bool tailed, bool notify) {
#define HMRT_PICK(HITS_, WALK_)                                                                         \
  return tailed ? reinterpret_cast<const void*>(&trace_persistent_kernel<HITS_, WALK_, true>)           \
                : reinterpret_cast<const void*>(&trace_persistent_kernel<HITS_, WALK_, false>)
#define HMRT_PICK_NOTIFY(WALK_)                                                                         \
  return tailed ? reinterpret_cast<const void*>(&trace_persistent_kernel<false, WALK_, true, true>)     \
                : reinterpret_cast<const void*>(&trace_persistent_kernel<false, WALK_, false, true>)
  if (notify) { /* hmrt_trace_host only: no instrumented variant */
    if (walk == kWalkJumpPow2) HMRT_PICK_NOTIFY(kWalkJumpPow2);
    if (walk == kWalkFastPow2) HMRT_PICK_NOTIFY(kWalkFastPow2);
    if (walk == kWalkFast) HMRT_PICK_NOTIFY(kWalkFast);
    HMRT_PICK_NOTIFY(kWalkReference);
  }
  if (hits) {
    if (walk == kWalkJumpPow2) HMRT_PICK(true, kWalkJumpPow2);
    if (walk == kWalkFastPow2) HMRT_PICK(true, kWalkFastPow2);
    if (walk == kWalkFast) HMRT_PICK(true, kWalkFast);
    HMRT_PICK(true, kWalkReference);
  }
  if (walk == kWalkJumpPow2) HMRT_PICK(false, kWalkJumpPow2);
  if (walk == kWalkFastPow2) HMRT_PICK(false, kWalkFastPow2);
  if (walk == kWalkFast) HMRT_PICK(false, kWalkFast);
  HMRT_PICK(false, kWalkReference);
#undef HMRT_PICK
#undef HMRT_PICK_NOTIFY
}

/* Scratch: kCallSets sets of `scratch_cap` slots; a call takes the next set (round robin), its slot 0 also holds the
 * top-level maximum and every launch of the call uses its own counter slot. */
static int ensure_scratch(hmrt_ctx* ctx, int slots) {
  if (ctx->scratch_cap >= slots) return 0;
  if (ctx->d_scratch) HMRT_CUDA(cudaFree(ctx->d_scratch)); /* cudaFree waits for in-flight work */
  ctx->d_scratch = nullptr;
  ctx->scratch_cap = 0;
  HMRT_CUDA(cudaMalloc(&ctx->d_scratch, sizeof(TraceScratch) * (size_t)slots * kCallSets));
  ctx->scratch_cap = slots;
  return 0;
}

/* Zero `slots` counters of the next scratch set and refresh the top-level maximum, on the context's stream.
 * *base_out = index of the set's first slot. */
static int prepare_trace(hmrt_ctx* ctx, int slots, int* base_out) {
  int rc = ensure_scratch(ctx, slots);
  if (rc) return rc;
  if (!ctx->d_stats) {
    HMRT_CUDA(cudaMalloc(&ctx->d_stats, 4 * sizeof(unsigned long long)));
    HMRT_CUDA(cudaMemsetAsync(ctx->d_stats, 0, 4 * sizeof(unsigned long long), ctx->stream));
  }
  ctx->call_set = (ctx->call_set + 1) % kCallSets;
  const int base = ctx->call_set * ctx->scratch_cap;
  *base_out = base;
  TraceScratch* scratch = static_cast<TraceScratch*>(ctx->d_scratch) + base;
  const uint32_t n_top = ctx->grid.coarse_sq; /* the coarsest level sits at float offset 0 */
  const bool need_hmax = ctx->trace_variant != 1;
  const int reduce_here = need_hmax && n_top <= kPrepTopMax;
  trace_prep_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->grid.pyramid, n_top, scratch, slots, reduce_here);
  HMRT_LAUNCHED(ctx);
  if (need_hmax && !reduce_here) { /* huge top level (e.g. a single-level "flat DDA" map): multi-CTA reduction */
    const unsigned blocks = (unsigned)((n_top + 1023u) / 1024u);
    top_level_max_kernel<<<blocks > 1184u ? 1184u : blocks, 256, 0, ctx->stream>>>(ctx->grid.pyramid, n_top, &scratch->hmax_key);
    HMRT_LAUNCHED(ctx);
  }
  return 0;
}

static int check_trace_args(const hmrt_ctx* ctx, int W, int H, const hmrt_camera* cams, int n_frames,
                            const hmrt_trace_opts* opts) {
  if (!ctx || !cams || !opts) return HMRT_E_ARG;
  if (!ctx->have_grid) return HMRT_E_STATE;
  if (W < 2 || H < 2 || n_frames < 1 || n_frames > 65535) return HMRT_E_ARG;
  if (opts->use_color_map && !ctx->grid.color_map) return HMRT_E_ARG;
  if (opts->tile_first < 0 || opts->tile_stride < 0) return HMRT_E_ARG;
  return 0;
}

/* Per-frame constants of launches with more than kInlineFrames frames: kCallSets sets of `frames_cap` entries. */
static int ensure_frames(hmrt_ctx* ctx, int n) {
  if (ctx->frames_cap >= n) return 0;
  if (ctx->d_frames) HMRT_CUDA(cudaFree(ctx->d_frames)); /* cudaFree waits for in-flight work */
  ctx->d_frames = nullptr;
  ctx->frames_cap = 0;
  HMRT_CUDA(cudaMalloc(&ctx->d_frames, sizeof(FrameConsts) * (size_t)n * kCallSets));
  ctx->frames_cap = n;
  return 0;
}

/* One persistent launch over n_frames cameras on `stream`, using scratch slot `slot` of the call's set (first slot
 * `base`) as its counter (prepare_trace must have run on the context's stream and be ordered before `stream`). */
/* `frames_at`: index into ctx->d_frames where this launch keeps its per-frame constants (launches of
 * one call that run on different streams must not share them). */
/* `tile_cnt` > 0 (single-frame launches only): render just the local tiles [tile_lo, tile_lo + tile_cnt) of the frame; d_rgb then
 * points at the first row of local tile `tile_lo`. */
static int launch_trace(hmrt_ctx* ctx, cudaStream_t stream, int base, int slot, int frames_at, int W, int H, const hmrt_camera* cams,
                        int n_frames, const hmrt_trace_opts* opts, uint8_t* d_rgb, hmrt_hit* d_hits, int tile_lo = 0, int tile_cnt = 0,
                        SegDone* seg_done = nullptr, bool overlapped = false, int segments = 1) {
  const int stride = opts->tile_stride > 0 ? opts->tile_stride : 1;
  const int n_tiles = (H + HMRT_ROW_TILE - 1) / HMRT_ROW_TILE;
  if (opts->tile_first >= n_tiles) return 0; /* nothing to render on this rank */
  if (!d_rgb) return HMRT_E_ARG;
  int local_tiles = (n_tiles - opts->tile_first + stride - 1) / stride;
  int tile_first = opts->tile_first;
  if (tile_cnt > 0) {
    if (n_frames != 1 || tile_lo < 0 || tile_lo >= local_tiles) return HMRT_E_ARG;
    tile_first += tile_lo * stride;
    local_tiles = tile_cnt < local_tiles - tile_lo ? tile_cnt : local_tiles - tile_lo;
  }

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
  p.full_layout = opts->full_frame_output ? 1 : 0;
  p.rows_local = p.full_layout ? H : rows_local(H, opts->tile_first, stride);
  p.tile_first = tile_first;
  p.tile_stride = stride;
  p.vec_store = (W % 16 == 0) && ((reinterpret_cast<uintptr_t>(d_rgb) & 15) == 0);
  p.chunks_x = (uint32_t)((W + kChunkW - 1) / kChunkW);
  p.chunks_per_frame = (uint32_t)local_tiles * 2u * p.chunks_x;
  p.total_chunks = p.chunks_per_frame * (uint32_t)n_frames; /* < 2^31, see frames_per_launch() */
  TraceScratch* scratch = static_cast<TraceScratch*>(ctx->d_scratch);
  p.hmax_key = &scratch[base].hmax_key;
  p.next_chunk = &scratch[base + slot].next_chunk;
  p.next_tail = &scratch[base + slot].next_tail;
  p.stats = ctx->d_stats;

  p.pixel_grid.wm1 = (float)(W - 1);
  p.pixel_grid.hm1 = (float)(H - 1);
  p.pixel_grid.rw1 = 1.0f / p.pixel_grid.wm1; /* IEEE single division: the correctly rounded reciprocal div_by needs */
  p.pixel_grid.rh1 = 1.0f / p.pixel_grid.hm1;
  p.pixel_grid.fast = 1;
  for (int i = 0; i < n_frames; ++i)
    for (int a = 0; a < 2; ++a) {
      const float m = fabsf(cams[i].frame_dim[a]);
      if (!(m >= 9.094947017729282e-13f && m <= 1.099511627776e12f)) p.pixel_grid.fast = 0; /* 2^-40 .. 2^40; NaN fails too */
    }

  if (n_frames <= kInlineFrames) {
    for (int i = 0; i < n_frames; ++i) make_frame_consts(cams[i], p.frame_inline[i]);
    p.frames = nullptr;
  } else {
    int rc = ensure_frames(ctx, frames_at + n_frames);
    if (rc) return rc;
    frames_at += ctx->call_set * ctx->frames_cap;
    /* pageable source: the runtime stages it before returning, so the stack buffer may die */
    FrameConsts local[64];
    for (int base = 0; base < n_frames; base += 64) {
      const int n = n_frames - base < 64 ? n_frames - base : 64;
      for (int i = 0; i < n; ++i) make_frame_consts(cams[base + i], local[i]);
      HMRT_CUDA(cudaMemcpyAsync(ctx->d_frames + frames_at + base, local, sizeof(FrameConsts) * (size_t)n, cudaMemcpyHostToDevice,
                                stream));
    }
    p.frames = ctx->d_frames + frames_at;
  }

  const bool pow2 = (ctx->grid.coarse_res & (ctx->grid.coarse_res - 1)) == 0;
  const int walk = ctx->trace_variant == 1 ? kWalkReference
                   : !pow2                 ? kWalkFast
                   : ctx->trace_variant == 2 ? kWalkJumpPow2
                                             : kWalkFastPow2;
  /* persistent grid: SMs x resident CTAs of the chosen instantiation, never more warps than chunks */
  /* tile-granular tail only where it pays: fewer than 64 chunks per resident warp (see the kernel comment) */
  /* ... and only for launches that run alone: when the caller overlaps successive calls (alternating streams, see hmrt_trace)
   * the next call's head fills the drain, and the lean kernel wins (a 1/8-frame-share step of the benchmark: 1.001 ms
   * without the tail, 1.049 ms with it, 0.984 ms = one eighth of the whole-frame step; benchmarks/shard_probe.py) */
  bool tailed = !overlapped && (unsigned long long)p.total_chunks < 64ull * (unsigned long long)ctx->sm_count * 5ull * kWarps;
  if (g_knob_tailed >= 0) tailed = g_knob_tailed != 0;
  const bool notify = seg_done != nullptr && !d_hits;
  const void* fn = pick_kernel(d_hits != nullptr, walk, tailed, notify);
  const int kslot = (notify ? 16 : 0) + (tailed ? 8 : 0) + (d_hits ? 4 : 0) + walk;
  p.seg_done = notify ? seg_done : nullptr;
  p.segments = (uint32_t)segments;
  p.strips_per_frame = p.chunks_per_frame / p.chunks_x;
  if (ctx->ctas_per_sm[kslot] == 0) {
    int per_sm = 0;
    HMRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kThreads, 0));
    ctx->ctas_per_sm[kslot] = per_sm < 1 ? 1 : per_sm;
  }
  const unsigned long long want = ((unsigned long long)p.total_chunks + kWarps - 1) / kWarps;
  const unsigned long long cap = (unsigned long long)ctx->sm_count * (unsigned long long)ctx->ctas_per_sm[kslot];
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  /* the tail must be long enough to absorb the longest whole chunk claimed just before it (four near-horizon tiles
   * can take ~10x the mean): 8 chunks per resident warp measured best (1: -4 %, everything tile by tile: -5 %) */
  const unsigned long long tail_per_warp = g_knob_tail_per_warp > 0 ? (unsigned long long)g_knob_tail_per_warp : 8ull;
  unsigned long long tail_chunks = (unsigned long long)grid * kWarps * tail_per_warp;
  if (tail_chunks > p.total_chunks) tail_chunks = p.total_chunks;
  if (!tailed) tail_chunks = 0;
  p.bulk_chunks = (uint32_t)(p.total_chunks - tail_chunks);
  p.tail_tiles = (uint32_t)(tail_chunks * 4ull);
  void* args[] = {&p};
  if (ctx->l2_first_level >= 1 && ctx->l2_first_level < ctx->grid.levels) {
    /* experiment (hmrt_set_l2_persist): the levels >= l2_first_level -- the front of the pyramid, coarsest first -- are
     * fetched with the persisting property, everything else of the window with the streaming one */
    int max_window = 0;
    HMRT_CUDA(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device));
    size_t bytes = (size_t)level_offset(ctx->grid, ctx->l2_first_level - 1) * sizeof(float);
    if (bytes > (size_t)max_window) bytes = (size_t)max_window;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeAccessPolicyWindow;
    attr.val.accessPolicyWindow.base_ptr = const_cast<float*>(ctx->grid.pyramid);
    attr.val.accessPolicyWindow.num_bytes = bytes;
    attr.val.accessPolicyWindow.hitRatio = ctx->l2_hit_ratio;
    attr.val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.stream = stream;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    HMRT_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
  } else {
    HMRT_CUDA(cudaLaunchKernel(fn, dim3(grid), dim3(kThreads), args, 0, stream));
  }
  HMRT_LAUNCHED(ctx);
  return 0;
}

}  // namespace hmrt

extern "C" {

/* frames one launch may carry so that its chunk indices stay below 2^31 */
static int frames_per_launch(int W, int H) {
  const unsigned long long per_frame = (unsigned long long)((H + HMRT_ROW_TILE - 1) / HMRT_ROW_TILE) * 2ull * (unsigned long long)((W + 31) / 32);
  const unsigned long long n = per_frame ? (1ull << 31) / per_frame : 1ull;
  return n < 1 ? 1 : (n > 65535 ? 65535 : (int)n);
}

int hmrt_trace(hmrt_ctx* ctx, int W, int H, const hmrt_camera* h_cameras, int n_frames,
               const hmrt_trace_opts* opts, uint8_t* d_rgb, hmrt_hit* d_hits) {
  int rc = hmrt::check_trace_args(ctx, W, H, h_cameras, n_frames, opts);
  if (rc) return rc;
  hmrt::DeviceGuard guard(ctx->device);
  const int per = frames_per_launch(W, H);
  const int n_launches = (n_frames + per - 1) / per;
  /* a caller that switches streams between consecutive calls (HMRT_MAX_CALLS_IN_FLIGHT) is overlapping them */
  const bool overlapped = ctx->have_last_trace_stream && ctx->last_trace_stream != ctx->stream;
  ctx->last_trace_stream = ctx->stream;
  ctx->have_last_trace_stream = true;
  int base = 0;
  rc = hmrt::prepare_trace(ctx, n_launches, &base);
  if (rc) return rc;
  const int stride = opts->tile_stride > 0 ? opts->tile_stride : 1;
  const size_t frame_px = (size_t)(opts->full_frame_output ? H : hmrt::rows_local(H, opts->tile_first, stride)) * (size_t)W;
  for (int l = 0; l < n_launches; ++l) {
    const int f0 = l * per, nf = n_frames - f0 < per ? n_frames - f0 : per;
    rc = hmrt::launch_trace(ctx, ctx->stream, base, l, 0, W, H, h_cameras + f0, nf, opts, d_rgb ? d_rgb + (size_t)f0 * frame_px * 3 : nullptr,
                            d_hits ? d_hits + (size_t)f0 * frame_px : nullptr, 0, 0, nullptr, overlapped);
    if (rc) return rc;
  }
  return 0;
}

/* cuStreamWaitValue32 (stream memory operation of the driver API), bound at run time: the runtime API has no equivalent.
 * "Wait until (int32)(*addr - value) >= 0" (flags = CU_STREAM_WAIT_VALUE_GEQ = 0). */
typedef int (*stream_wait_value32_fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
static stream_wait_value32_fn stream_wait_value32() {
  static stream_wait_value32_fn fn = [] {
    void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return (stream_wait_value32_fn) nullptr;
    void* f = dlsym(h, "cuStreamWaitValue32_v2");
    if (!f) f = dlsym(h, "cuStreamWaitValue32");
    return reinterpret_cast<stream_wait_value32_fn>(f);
  }();
  return fn;
}

/* behind the traversal launch, on the same stream: every segment counts as complete (a safety net: a copy waiting on a
 * flag can never outlive the launch that should have set it) */
__global__ void seg_force_kernel(hmrt::SegDone* sd, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t*>(&sd[i].flag) = 1u;
  }
}

/*
 * hmrt_trace_host, streaming form: ONE persistent launch over all frames (the same launch shape as hmrt_trace: no per-frame
 * launch, no per-frame tail) whose units report to per-(frame, row segment) counters; the copy stream waits on each
 * segment's flag with a stream memory operation and copies the segment's rows the moment its last tile is stored.  The
 * device->host copies therefore trail the traversal by a quarter of a frame instead of a whole launch.
 */
static int trace_host_streamed(hmrt_ctx* ctx, int W, int H, const hmrt_camera* h_cameras, int n_frames, const hmrt_trace_opts* opts,
                               uint8_t* h_rgb, size_t frame_bytes, stream_wait_value32_fn wait_value, uint8_t* d_fb) {
  const int stride = opts->tile_stride > 0 ? opts->tile_stride : 1;
  const int n_tiles = (H + HMRT_ROW_TILE - 1) / HMRT_ROW_TILE;
  const int local_tiles = opts->tile_first < n_tiles ? (n_tiles - opts->tile_first + stride - 1) / stride : 0;
  if (local_tiles == 0) return 0;
  const uint32_t strips = (uint32_t)local_tiles * 2u;
  const size_t row_bytes = (size_t)W * 3, rows_total = frame_bytes / row_bytes;
  /* a single frame is latency-bound: release it in quarters; a batch is bound by the device->host link, where every stream
   * memory operation in front of a copy costs ~25 us of copy-stream time: one flag per frame */
  const int segs = ctx->host_segments > 0 ? ctx->host_segments : (n_frames == 1 ? hmrt::kMaxSegments : 1);
  const int n_seg = n_frames * segs;
  if (ctx->seg_cap < n_seg) {
    if (ctx->d_seg_done) HMRT_CUDA(cudaFree(ctx->d_seg_done));
    ctx->d_seg_done = nullptr;
    ctx->seg_cap = 0;
    HMRT_CUDA(cudaMalloc(&ctx->d_seg_done, sizeof(hmrt::SegDone) * (size_t)n_seg));
    ctx->seg_cap = n_seg;
  }
  hmrt::SegDone* sd = static_cast<hmrt::SegDone*>(ctx->d_seg_done);
  const int per = frames_per_launch(W, H);
  const int n_launches = (n_frames + per - 1) / per;
  int base = 0;
  int rc = hmrt::prepare_trace(ctx, n_launches, &base);
  if (rc) return rc;
  HMRT_CUDA(cudaMemsetAsync(sd, 0, sizeof(hmrt::SegDone) * (size_t)n_seg, ctx->stream));
  HMRT_CUDA(cudaEventRecord(ctx->prep_event, ctx->stream));
  HMRT_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->prep_event, 0)); /* the flags are zero before anybody polls them */
  for (int l = 0; l < n_launches; ++l) {
    const int f0 = l * per, nf = n_frames - f0 < per ? n_frames - f0 : per;
    rc = hmrt::launch_trace(ctx, ctx->stream, base, l, 0, W, H, h_cameras + f0, nf, opts, d_fb + (size_t)f0 * frame_bytes, nullptr, 0, 0,
                            sd + (size_t)f0 * segs, false, segs);
    if (rc) break;
  }
  /* the safety net goes in even when a launch failed: the waits below must always be released */
  seg_force_kernel<<<(n_seg + 255) / 256, 256, 0, ctx->stream>>>(sd, n_seg);
  if (rc == 0) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    ctx->launches++;
  }
  if (rc) return rc;
  for (int f = 0; f < n_frames; ++f)
    for (int seg = segs - 1; seg >= 0; --seg) { /* rows are traced from the top of the frame down */
      const uint32_t s_lo = ((uint32_t)seg * strips + (uint32_t)segs - 1) / (uint32_t)segs,
                     s_hi = (((uint32_t)seg + 1u) * strips + (uint32_t)segs - 1) / (uint32_t)segs;
      if (s_hi <= s_lo) continue;
      const size_t r0 = (size_t)s_lo * 4, r1 = (size_t)s_hi * 4 < rows_total ? (size_t)s_hi * 4 : rows_total;
      if (r1 <= r0) continue;
      const unsigned long long flag = reinterpret_cast<unsigned long long>(&sd[(size_t)f * segs + seg].flag);
      const int drc = wait_value(ctx->copy_stream, flag, 1u, 0u);
      if (drc != 0) return 999; /* cudaErrorUnknown: the driver refused the stream memory operation */
      const size_t off = (size_t)f * frame_bytes + r0 * row_bytes;
      HMRT_CUDA(cudaMemcpyAsync(h_rgb + off, d_fb + off, (r1 - r0) * row_bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
    }
  return 0;
}

/* The launch / copy schedule of hmrt_trace_host; any error leaves through the caller, which drains the internal streams
 * first so that no copy into h_rgb and no kernel writing d_fb is still in flight when the caller sees the error. */
static int trace_host_enqueue(hmrt_ctx* ctx, int W, int H, const hmrt_camera* h_cameras, int n_frames, const hmrt_trace_opts* opts,
                              uint8_t* h_rgb, size_t frame_bytes, uint8_t* d_fb) {
  const int stride = opts->tile_stride > 0 ? opts->tile_stride : 1;
  /*
   * One launch per frame group, alternating between two internal streams so that the tail of launch l
   * overlaps the head of launch l+1, each followed by its device->host copy on a third stream: the
   * copy of frame f overlaps the traversal of the following frames (the reference serialises
   * upload, trace and GL read-back every frame, main.cpp:947-966).  The top-level maximum is
   * refreshed once per call on the context's stream; everything is ordered after prior work on it.
   */
  /* frames per launch: enough rays (>= ~4 M) to amortise launch + tail, e.g. 1 whole 4K frame on one
   * GPU, 4 frames when a rank only owns 1/8 of every frame */
  const size_t rays_per_frame = frame_bytes / 3;
  const size_t group_rays = 4000000;
  int group = (int)((group_rays + rays_per_frame - 1) / rays_per_frame);
  if (group < 1) group = 1;
  if (group > n_frames) group = n_frames;
  /* Launch sizes: one group per launch.  (Development knob: `mid` groups per launch between the first and the last one --
   * measured worse: 8.83 / 9.23 / 9.54 / 9.99 ms per 16-frame 4K step for 1 / 2 / 3 / 4, because the copy stream, not the
   * traversal, is the critical path and larger launches release their frames later.) */
  const int mid = group * (ctx->host_mid_groups > 0 ? ctx->host_mid_groups : 1);
  std::vector<int> sizes;
  {
    int remaining = n_frames;
    const int tail = group < remaining ? group : remaining;
    remaining -= tail;
    if (remaining > 0) {
      const int first = group < remaining ? group : remaining;
      sizes.push_back(first);
      remaining -= first;
    }
    while (remaining > 0) {
      const int k = mid < remaining ? mid : remaining;
      sizes.push_back(k);
      remaining -= k;
    }
    sizes.push_back(tail);
  }
  const int n_launches = (int)sizes.size();
  /* The device->host copy of the LAST launch is the only one nothing overlaps.  When that launch is a single frame, it is
   * cut into kLastParts tile ranges, each followed by its own copy, so that only a quarter of a frame's copy stays exposed
   * (4K: 0.45 ms -> 0.11 ms per call). */
  constexpr int kLastParts = 4;
  const int n_tiles = (H + HMRT_ROW_TILE - 1) / HMRT_ROW_TILE;
  const int local_tiles = opts->tile_first < n_tiles ? (n_tiles - opts->tile_first + stride - 1) / stride : 0;
  const int last_frames = sizes.back();
  const bool split_last = last_frames == 1 && local_tiles >= 4 * kLastParts && rays_per_frame >= 2000000;
  /* (Releasing the FIRST frame in quarters as well, to start the copy stream earlier, measured 1 % slower: 8.85 vs 8.78 ms.) */
  int base = 0;
  int rc = hmrt::prepare_trace(ctx, n_launches + (split_last ? kLastParts - 1 : 0), &base);
  if (rc) return rc;
  rc = hmrt::ensure_frames(ctx, n_frames); /* sized up front: no reallocation while launches are in flight */
  if (rc) return rc;
  HMRT_CUDA(cudaEventRecord(ctx->prep_event, ctx->stream));
  for (int i = 0; i < 2; ++i) HMRT_CUDA(cudaStreamWaitEvent(ctx->frame_stream[i], ctx->prep_event, 0));
  int slot = 0; /* one work-counter slot per launch; launches alternate between the two frame streams */
  auto launch_parts = [&](int f0) -> int { /* one frame as kLastParts tile ranges, each followed by its copy */
    const size_t row_bytes = (size_t)W * 3, rows_total = frame_bytes / row_bytes;
    for (int k = 0; k < kLastParts; ++k, ++slot) {
      const int t0 = (int)((long long)local_tiles * k / kLastParts), t1 = (int)((long long)local_tiles * (k + 1) / kLastParts);
      const size_t r0 = (size_t)t0 * HMRT_ROW_TILE, r1 = (size_t)t1 * HMRT_ROW_TILE < rows_total ? (size_t)t1 * HMRT_ROW_TILE : rows_total;
      const size_t off = (size_t)f0 * frame_bytes + r0 * row_bytes;
      cudaStream_t st = ctx->frame_stream[slot & 1];
      int prc = hmrt::launch_trace(ctx, st, base, slot, f0, W, H, h_cameras + f0, 1, opts, d_fb + off, nullptr, t0, t1 - t0);
      if (prc) return prc;
      HMRT_CUDA(cudaEventRecord(ctx->frame_event[slot & 1], st));
      HMRT_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->frame_event[slot & 1], 0));
      HMRT_CUDA(cudaMemcpyAsync(h_rgb + off, d_fb + off, (r1 - r0) * row_bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
    }
    return 0;
  };
  const int whole_end = split_last ? n_launches - 1 : n_launches;
  int f_next = 0;
  for (int l = 0; l < whole_end; ++l) {
    const int f0 = f_next, nf = sizes[(size_t)l];
    f_next += nf;
    cudaStream_t st = ctx->frame_stream[slot & 1];
    /* the launches alternate between two streams, i.e. each one's drain runs under the next one's head: the lean kernel
     * without the tile-granular tail (see launch_trace), except for the very last whole launch of the call */
    const bool overlapped = l + 1 < whole_end || split_last;
    rc = hmrt::launch_trace(ctx, st, base, slot, f0, W, H, h_cameras + f0, nf, opts, d_fb + (size_t)f0 * frame_bytes, nullptr, 0, 0, nullptr,
                            overlapped);
    if (rc) return rc;
    HMRT_CUDA(cudaEventRecord(ctx->frame_event[slot & 1], st));
    HMRT_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->frame_event[slot & 1], 0));
    HMRT_CUDA(cudaMemcpyAsync(h_rgb + (size_t)f0 * frame_bytes, d_fb + (size_t)f0 * frame_bytes, frame_bytes * (size_t)nf,
                              cudaMemcpyDeviceToHost, ctx->copy_stream));
    ++slot;
  }
  if (split_last) {
    rc = launch_parts(n_frames - 1);
    if (rc) return rc;
  }
  return 0;
}

/* Everything hmrt_trace_host_begin enqueued -- kernels writing d_fb, copies into the callers' buffers -- has finished. */
static int trace_host_drain(hmrt_ctx* ctx, bool collect = true) {
  cudaError_t e0 = cudaStreamSynchronize(ctx->stream);
  cudaError_t e = ctx->copy_stream ? cudaStreamSynchronize(ctx->frame_stream[0]) : cudaSuccess;
  cudaError_t e1 = ctx->copy_stream ? cudaStreamSynchronize(ctx->frame_stream[1]) : cudaSuccess;
  cudaError_t e2 = ctx->copy_stream ? cudaStreamSynchronize(ctx->copy_stream) : cudaSuccess;
  if (collect) ctx->host_waited = ctx->host_begun; /* else: the calls stay "in flight" for their hmrt_trace_host_wait */
  if (e0 != cudaSuccess) return (int)e0;
  if (e != cudaSuccess) return (int)e;
  if (e1 != cudaSuccess) return (int)e1;
  return (int)e2;
}

int hmrt_trace_host_begin(hmrt_ctx* ctx, int W, int H, const hmrt_camera* h_cameras, int n_frames, const hmrt_trace_opts* opts,
                          uint8_t* h_rgb) {
  int rc = hmrt::check_trace_args(ctx, W, H, h_cameras, n_frames, opts);
  if (rc) return rc;
  if (!h_rgb || opts->full_frame_output) return HMRT_E_ARG; /* the host copy is always compact */
  if (ctx->host_begun - ctx->host_waited >= (unsigned)HMRT_MAX_HOST_CALLS_IN_FLIGHT) return HMRT_E_STATE;
  hmrt::DeviceGuard guard(ctx->device);
  const int stride = opts->tile_stride > 0 ? opts->tile_stride : 1;
  const size_t frame_bytes = (size_t)hmrt::rows_local(H, opts->tile_first, stride) * (size_t)W * 3;
  const size_t bytes = frame_bytes * (size_t)n_frames;
  const size_t half = (bytes + 255) & ~(size_t)255;
  if (!ctx->copy_stream) {
    HMRT_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      HMRT_CUDA(cudaStreamCreateWithFlags(&ctx->frame_stream[i], cudaStreamNonBlocking));
      HMRT_CUDA(cudaEventCreateWithFlags(&ctx->frame_event[i], cudaEventDisableTiming));
      HMRT_CUDA(cudaEventCreateWithFlags(&ctx->host_done[i], cudaEventDisableTiming));
    }
    HMRT_CUDA(cudaEventCreateWithFlags(&ctx->prep_event, cudaEventDisableTiming));
  }
  if (ctx->fb_half < half) { /* grow: nothing may be running in the old halves (the calls keep their place in the wait order) */
    rc = trace_host_drain(ctx, false);
    if (rc) return rc;
    if (ctx->d_fb) HMRT_CUDA(cudaFree(ctx->d_fb));
    ctx->d_fb = nullptr;
    ctx->fb_cap = ctx->fb_half = 0;
    HMRT_CUDA(cudaMalloc(&ctx->d_fb, half * HMRT_MAX_HOST_CALLS_IN_FLIGHT));
    ctx->fb_cap = half * HMRT_MAX_HOST_CALLS_IN_FLIGHT;
    ctx->fb_half = half;
  }
  const unsigned call = ctx->host_begun;
  uint8_t* d_fb = ctx->d_fb + (size_t)(call % HMRT_MAX_HOST_CALLS_IN_FLIGHT) * ctx->fb_half;
  /* Schedules (measured on B200, 4K frames, profiles/raw_r02/trace_host_variants.txt): for a batch of frames the per-group
   * schedule (one lean launch per frame on alternating streams, each followed by its copy) takes 8.83 ms per 16-frame step
   * against 8.00 ms for the traversal alone and 7.13 ms for the copies alone (55.8 GB/s, the Gen5 x16 link); the streamed
   * schedule takes 9.2-9.5 ms whatever the segment count (its kernel alone is 6 % slower: a fence and an atomic per unit of
   * work).  A SINGLE frame is latency-bound, where releasing each quarter frame as it completes wins 6 % (0.90 vs 0.96 ms).
   * Variant 0 picks accordingly. */
  const bool streamed = ctx->host_variant == 2 || (ctx->host_variant == 0 && n_frames == 1);
  stream_wait_value32_fn wait_value = streamed ? stream_wait_value32() : nullptr;
  if (bytes == 0) {
    rc = 0; /* a rank that owns no tile of the frame: nothing to enqueue, the call still counts */
  } else if (wait_value) {
    /* the segment flags are one set per context: a streamed call starts with nothing else in flight */
    if (ctx->host_begun != ctx->host_waited) {
      rc = trace_host_drain(ctx, false);
      if (rc) return rc;
    }
    rc = trace_host_streamed(ctx, W, H, h_cameras, n_frames, opts, h_rgb, frame_bytes, wait_value, d_fb);
  } else {
    rc = trace_host_enqueue(ctx, W, H, h_cameras, n_frames, opts, h_rgb, frame_bytes, d_fb);
  }
  if (rc) { /* whatever was enqueued before the failure finishes before the caller sees the error */
    trace_host_drain(ctx);
    return rc;
  }
  /* every kernel of the call is followed by its copy on the copy stream: one event behind the last copy covers the call */
  cudaError_t e = cudaEventRecord(ctx->host_done[call % HMRT_MAX_HOST_CALLS_IN_FLIGHT], ctx->copy_stream);
  if (e != cudaSuccess) {
    trace_host_drain(ctx);
    return (int)e;
  }
  ctx->host_begun = call + 1;
  return 0;
}

int hmrt_trace_host_wait(hmrt_ctx* ctx) {
  if (!ctx) return HMRT_E_ARG;
  if (ctx->host_begun == ctx->host_waited) return HMRT_E_STATE;
  hmrt::DeviceGuard guard(ctx->device);
  cudaError_t e = cudaEventSynchronize(ctx->host_done[ctx->host_waited % HMRT_MAX_HOST_CALLS_IN_FLIGHT]);
  ++ctx->host_waited;
  if (e != cudaSuccess) { /* a failed call poisons the streams: leave nothing in flight */
    trace_host_drain(ctx);
    return (int)e;
  }
  return 0;
}

int hmrt_trace_host(hmrt_ctx* ctx, int W, int H, const hmrt_camera* h_cameras, int n_frames,
                    const hmrt_trace_opts* opts, uint8_t* h_rgb) {
  int rc = hmrt::check_trace_args(ctx, W, H, h_cameras, n_frames, opts);
  if (rc) return rc;
  if (!h_rgb || opts->full_frame_output) return HMRT_E_ARG;
  hmrt::DeviceGuard guard(ctx->device);
  const int stride = opts->tile_stride > 0 ? opts->tile_stride : 1;
  if ((size_t)hmrt::rows_local(H, opts->tile_first, stride) * (size_t)W * 3 * (size_t)n_frames == 0) return 0;
  /* the synchronous form collects everything, also calls begun earlier and not yet waited for */
  if (ctx->host_begun - ctx->host_waited >= (unsigned)HMRT_MAX_HOST_CALLS_IN_FLIGHT) {
    rc = trace_host_drain(ctx);
    if (rc) return rc;
  }
  rc = hmrt_trace_host_begin(ctx, W, H, h_cameras, n_frames, opts, h_rgb);
  /* one exit: whatever was enqueued -- kernels writing d_fb, copies into h_rgb -- has finished when the caller gets
   * control back, also on the error paths */
  const int drc = trace_host_drain(ctx);
  return rc ? rc : drc;
}

/* Development knobs of the traversal launcher (not part of include/hmrt.h): key 0 = tile-granular tail (-1 auto, 0 off,
 * 1 on), key 1 = chunks per resident warp handed out tile by tile (0 = built-in 8). */
int hmrt_debug_trace_knob(int key, int value) {
  if (key == 0 && value >= -1 && value <= 1) hmrt::g_knob_tailed = value;
  else if (key == 1 && value >= 0 && value <= 64) hmrt::g_knob_tail_per_warp = value;
  else return HMRT_E_ARG;
  return 0;
}

/* Counters of the INSTRUMENTED kernels (d_hits != NULL) since the last reset: out[0] = rays, out[1] = loop iterations
 * (== height fetches of the reference algorithm, CudaKernel.cu:153-176), out[2] = the share of them the production walk
 * spent in its air phase (no memory access), out[3] = 0.  Synchronises the context's stream. */
int hmrt_trace_stats(hmrt_ctx* ctx, uint64_t out[4], int reset) {
  if (!ctx || !out) return HMRT_E_ARG;
  hmrt::DeviceGuard guard(ctx->device);
  out[0] = out[1] = out[2] = out[3] = 0;
  if (!ctx->d_stats) return 0;
  HMRT_CUDA(cudaMemcpyAsync(out, ctx->d_stats, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  if (reset) HMRT_CUDA(cudaMemsetAsync(ctx->d_stats, 0, 4 * sizeof(uint64_t), ctx->stream));
  HMRT_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // extern "C"
