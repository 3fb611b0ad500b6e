/*
 * raster.cu -- point-cloud -> heightmap rasterisation and the max-mipmap build.
 *
 * CPU semantics being replaced: the inner loop of loadLASToSection
 * (GPUHeightmapRaytracer/src/main.cpp:193-234).  Per point the reference computes the finest
 * cell, skips out-of-section / class-7 points, writes the colour (last writer in file order
 * wins, :223-224) and pushes max(z) up all pyramid levels with an early break (:227-233).
 * Because parent >= child is an invariant of that loop and sections start at +0.0f (:259),
 * its fixed point is
 *     finest = max over points (floored at +0),   level i+1 = max of 2x2 children,
 * independent of point order.  Hence two kernels:
 *   (a) scatter: one thread per point, RED.MAX on the int view of the finest level
 *       (non-negative floats order like their bit patterns) -> bit-exact, order-independent;
 *       colours go through a 64-bit (file index, rgb) key so "last writer wins" is deterministic;
 *   (b) mip build: one 128x128 finest tile per CTA, all 7 coarser levels in one pass.
 */
#include "raster_common.cuh"

namespace hmrt {

/* main.cpp:223-233 at level 0 */
__device__ __forceinline__ void bin_point(const ScatterParams& sp, double gx, double gy, double gz, int cls,
                                          uint32_t rgb, int64_t file_index, int* __restrict__ finest,
                                          unsigned long long* __restrict__ keys) {
  uint32_t cell;
  float fZ;
  if (!point_to_cell(sp, gx, gy, gz, cls, cell, fZ)) return;
  HMRT_DCHECK(cell < (uint32_t)sp.res0 * (uint32_t)sp.res0);
  if (keys) /* :223-224; +1 so that key 0 means "never written" */
    atomicMax(keys + cell, ((unsigned long long)(file_index + 1) << 24) | rgb);
  /* :227-233.  Negative or NaN heights never replace the +0 floor in the reference (`buf <= fZ` is
   * false), -0.0f maps to INT_MIN and is a no-op here. */
  if (fZ >= 0.0f) atomicMax(finest + cell, __float_as_int(fZ));
}

/*
 * LAS records: a CTA streams its 256 consecutive records (256 * record_len contiguous bytes)
 * into shared memory with 16-byte loads, then each thread decodes its own record from there:
 * the global reads are fully coalesced even for the odd 26/34-byte record sizes.
 */
__global__ void __launch_bounds__(kScatterThreads)
scatter_las_kernel(const uint8_t* __restrict__ records, int64_t n, int record_len, const __grid_constant__ ScatterParams sp,
                   int64_t first_index, int* __restrict__ finest, unsigned long long* __restrict__ keys, int staged) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int64_t first = (int64_t)blockIdx.x * kScatterThreads;
  const int64_t i = first + threadIdx.x;
  const uint8_t* rec;
  if (staged) {
    const int64_t remaining = n - first;
    const int count = remaining < kScatterThreads ? (int)remaining : kScatterThreads;
    const int bytes = count * record_len;
    const uint8_t* src = records + first * record_len; /* 16-byte aligned: 256*record_len % 16 == 0 */
    const int vec = bytes >> 4;
    for (int v = threadIdx.x; v < vec; v += kScatterThreads)
      reinterpret_cast<uint4*>(smem)[v] = __ldg(reinterpret_cast<const uint4*>(src) + v);
    for (int b = (vec << 4) + threadIdx.x; b < bytes; b += kScatterThreads) smem[b] = __ldg(src + b);
    __syncthreads();
    rec = smem + threadIdx.x * record_len;
  } else {
    rec = records + i * record_len;
  }
  if (i >= n) return;
  /* libLAS 1.8.0 Point::GetX(): raw * scale + offset, two roundings in double */
  const RecordHead rh = load_record_head(rec);
  const double gx = __dadd_rn(__dmul_rn(int_to_double(rh.x), sp.scale[0]), sp.offset[0]);
  const double gy = __dadd_rn(__dmul_rn(int_to_double(rh.y), sp.scale[1]), sp.offset[1]);
  const double gz = __dadd_rn(__dmul_rn(int_to_double(rh.z), sp.scale[2]), sp.offset[2]);
  const int cls = sp.cls_off == 15 ? (int)((rh.tail >> 24) & 0x1f) : 0; /* Classification::GetClass(): low 5 bits of byte 15 */
  uint32_t rgb = 0;
  if (keys && sp.rgb_off >= 0) {
    const uint8_t* c = rec + sp.rgb_off;
    const uint32_t r = c[0] | (c[1] << 8), g = c[2] | (c[3] << 8), b = c[4] | (c[5] << 8);
    rgb = (color16(r) << 16) | (color16(g) << 8) | color16(b);
  }
  bin_point(sp, gx, gy, gz, cls, rgb, first_index + i, finest, keys);
}

/* PointdataGenerator output: float32 x y z triples; 3 coalesced loads per thread */
__global__ void __launch_bounds__(kScatterThreads)
scatter_xyz_kernel(const float* __restrict__ xyz, int64_t n, const __grid_constant__ ScatterParams sp,
                   int* __restrict__ finest) {
  __shared__ float s[kScatterThreads * 3];
  const int64_t first = (int64_t)blockIdx.x * kScatterThreads;
  const int64_t remaining = n - first;
  const int count = remaining < kScatterThreads ? (int)remaining : kScatterThreads;
  for (int k = threadIdx.x; k < count * 3; k += kScatterThreads) s[k] = __ldg(xyz + first * 3 + k);
  __syncthreads();
  if (threadIdx.x >= count) return;
  const float x = s[threadIdx.x * 3], y = s[threadIdx.x * 3 + 1], z = s[threadIdx.x * 3 + 2];
  bin_point(sp, (double)x, (double)y, (double)z, 0, 0, 0, finest, nullptr);
}

/* How spatially ordered is the input?  Eight windows of 65536 consecutive records are sampled (2048 records
 * each) and every sample sets the bit of its tile in the window's bitmap.  A survey-ordered file touches a
 * handful of tiles per window (direct atomics stay L2-resident), an unordered cloud touches most of them. */
constexpr int kProbeWindows = 8, kProbeSamples = 2048, kProbeWords = 128; /* 4096 tile bits per window */
__global__ void __launch_bounds__(256) locality_probe_kernel(const uint8_t* __restrict__ records, int64_t n, int record_len,
                                                             const __grid_constant__ ScatterParams sp, int tile_shift, int tiles_x, uint32_t* out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x; /* kProbeWindows * kProbeSamples threads */
  const int w = s / kProbeSamples, j = s % kProbeSamples;
  if (w >= kProbeWindows) return;
  const int64_t span = n < 65536 ? n : 65536;
  const int64_t start = (n - span) / (kProbeWindows - 1 > 0 ? kProbeWindows - 1 : 1) * w;
  const int64_t i = start + (int64_t)j * span / kProbeSamples;
  if (i >= n) return;
  const uint8_t* rec = records + i * record_len;
  const RecordHead rh = load_record_head(rec);
  const double gx = __dadd_rn(__dmul_rn((double)rh.x, sp.scale[0]), sp.offset[0]);
  const double gy = __dadd_rn(__dmul_rn((double)rh.y, sp.scale[1]), sp.offset[1]);
  uint32_t cell;
  float fZ;
  if (!point_to_cell(sp, gx, gy, 0.0, 0, cell, fZ)) return;
  const uint32_t tile = ((cell / (uint32_t)sp.res0) >> tile_shift) * (uint32_t)tiles_x + ((cell % (uint32_t)sp.res0) >> tile_shift);
  atomicOr(out + w * kProbeWords + (tile >> 5), 1u << (tile & 31));
}

__global__ void resolve_colors_kernel(const unsigned long long* __restrict__ keys, uint8_t* __restrict__ cmap,
                                      int64_t n_cells) {
  /* each thread packs 4 cells = 12 bytes = 3 words, so the stores are 4-byte and coalesced */
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t c0 = q * 4;
  if (c0 >= n_cells) return;
  if (c0 + 4 <= n_cells) {
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (uint32_t)(__ldg(keys + c0 + k) & 0xffffffull);
    /* byte order r,g,b per cell; v = r<<16 | g<<8 | b */
    uint8_t bytes[12];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      bytes[3 * k] = (v[k] >> 16) & 0xff;
      bytes[3 * k + 1] = (v[k] >> 8) & 0xff;
      bytes[3 * k + 2] = v[k] & 0xff;
    }
    uint32_t* dst = reinterpret_cast<uint32_t*>(cmap + c0 * 3); /* c0*3 is a multiple of 12 */
#pragma unroll
    for (int w = 0; w < 3; ++w)
      dst[w] = bytes[4 * w] | (bytes[4 * w + 1] << 8) | (bytes[4 * w + 2] << 16) | ((uint32_t)bytes[4 * w + 3] << 24);
  } else {
    for (int64_t c = c0; c < n_cells; ++c) {
      const uint32_t v = (uint32_t)(__ldg(keys + c) & 0xffffffull);
      cmap[c * 3] = (v >> 16) & 0xff;
      cmap[c * 3 + 1] = (v >> 8) & 0xff;
      cmap[c * 3 + 2] = v & 0xff;
    }
  }
}

/* ---------------------------------------------------------------------------------------------
 * Max-mipmap build.  Fused kernel: CTA = 512 threads = one 128 x 128 tile of the finest level.
 * Warp w owns finest rows [8w, 8w+8), lane l owns columns [4l, 4l+4): eight 16-byte loads per
 * thread (each warp-load is one fully coalesced 512-byte row segment), levels 1 and 2 come
 * straight out of registers, level 3 needs one warp shuffle, levels 4..7 are reduced by the
 * first warps through shared memory.  Every level is written once, nothing is re-read from HBM:
 * traffic = 4*R0^2 read + 4*R0^2/3 written (the algorithmic minimum).
 */
__global__ void __launch_bounds__(512) build_mips_fused_kernel(const __grid_constant__ MipParams mp) {
  const float* l0 = mp.pyramid + mp.idx[0];
  mips_tile(mp, l0 + (size_t)blockIdx.y * 128 * mp.res0 + (size_t)blockIdx.x * 128, (size_t)mp.res0, false, blockIdx.x, blockIdx.y);
}

/* generic one-level step (grids that do not tile by 128, and levels beyond the 8th) */
__global__ void build_mip_level_kernel(const float* __restrict__ fine, float* __restrict__ coarse, int rc) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= rc || z >= rc) return;
  const size_t rf = (size_t)rc * 2;
  /* scalar loads: a level base is only guaranteed 4-byte alignment (odd coarse_res) */
  const float* r0 = fine + (size_t)(2 * z) * rf + 2 * x;
  const float* r1 = r0 + rf;
  coarse[(size_t)z * rc + x] = max4(__ldg(r0), __ldg(r0 + 1), __ldg(r1), __ldg(r1 + 1));
}

}  // namespace hmrt

extern "C" {

int hmrt_clear_section(hmrt_ctx* ctx, float* d_pyramid, int coarse_res, int levels, uint64_t* d_color_keys,
                       hmrt_color* d_color_map) {
  if (!ctx || !d_pyramid) return HMRT_E_ARG;
  int res[HMRT_MAX_LEVELS];
  int64_t total = 0;
  int rc = hmrt::pyramid_layout(coarse_res, levels, res, nullptr, &total);
  if (rc) return rc;
  hmrt::DeviceGuard guard(ctx->device);
  const size_t cells = (size_t)res[0] * res[0];
  HMRT_CUDA(cudaMemsetAsync(d_pyramid, 0, sizeof(float) * (size_t)total, ctx->stream)); /* main.cpp:259 */
  if (d_color_keys) HMRT_CUDA(cudaMemsetAsync(d_color_keys, 0, sizeof(uint64_t) * cells, ctx->stream));
  if (d_color_map) HMRT_CUDA(cudaMemsetAsync(d_color_map, 0, 3 * cells, ctx->stream)); /* main.cpp:260 */
  return 0;
}

int hmrt_scatter_las(hmrt_ctx* ctx, const uint8_t* d_records, int64_t n, int record_len, int point_format,
                     const hmrt_las_transform* xf, int64_t first_index, float* d_pyramid, int coarse_res, int levels,
                     uint64_t* d_color_keys) {
  if (!ctx || !xf || !d_pyramid || n < 0 || first_index < 0) return HMRT_E_ARG;
  if (point_format < 0 || point_format > 3 || record_len < hmrt::kLasMinLen[point_format] || record_len > 65535)
    return HMRT_E_ARG;
  if (n > 0 && !d_records) return HMRT_E_ARG;
  if (first_index + n >= ((int64_t)1 << 39)) return HMRT_E_ARG; /* key = (index + 1) << 24 | rgb stays below 2^63: signed and unsigned max agree */
  int res[HMRT_MAX_LEVELS];
  int64_t idx[HMRT_MAX_LEVELS];
  int rc = hmrt::pyramid_layout(coarse_res, levels, res, idx, nullptr);
  if (rc) return rc;
  if (n == 0) return 0;
  hmrt::ScatterParams sp;
  rc = hmrt::fill_scatter_params(xf, res[0], sp);
  if (rc) return rc;
  sp.cls_off = 15;
  sp.rgb_off = hmrt::kLasRgbOff[point_format];
  hmrt::DeviceGuard guard(ctx->device);
  int* finest = reinterpret_cast<int*>(d_pyramid + idx[0]);

  /* ---- choose the path: tile-binned (rasterx.cu) for large, spatially unordered clouds on grids beyond L2 ---- */
  const int tile_shift = hmrt::binned_tile_shift(res[0], 1);
  const int tiles_x = (res[0] + (1 << tile_shift) - 1) >> tile_shift;
  const int n_tiles = tiles_x * tiles_x;
  bool binned = false;
  const bool binnable = n_tiles >= 16 && (reinterpret_cast<uintptr_t>(d_records) & 15) == 0 && record_len <= 64;
  if (binnable && ctx->scatter_mode == 2) binned = true; /* forced (tests, benchmarks) */
  if (binnable && ctx->scatter_mode == 0 && n >= (int64_t)1 << 22 && (size_t)res[0] * res[0] * 4 > ((size_t)96 << 20)) {
    /* the probe costs one stream synchronisation: once per input (first_index == 0), later chunks of the same file reuse the verdict */
    if (first_index == 0 || ctx->probe_verdict < 0) {
      const size_t probe_bytes = sizeof(uint32_t) * hmrt::kProbeWindows * hmrt::kProbeWords;
      if (!ctx->d_probe) HMRT_CUDA(cudaMalloc(&ctx->d_probe, probe_bytes));
      HMRT_CUDA(cudaMemsetAsync(ctx->d_probe, 0, probe_bytes, ctx->stream));
      hmrt::locality_probe_kernel<<<hmrt::kProbeWindows * hmrt::kProbeSamples / 256, 256, 0, ctx->stream>>>(d_records, n, record_len, sp,
                                                                                                         tile_shift, tiles_x, ctx->d_probe);
      HMRT_LAUNCHED(ctx);
      uint32_t bits[hmrt::kProbeWindows * hmrt::kProbeWords];
      HMRT_CUDA(cudaMemcpyAsync(bits, ctx->d_probe, probe_bytes, cudaMemcpyDeviceToHost, ctx->stream));
      HMRT_CUDA(cudaStreamSynchronize(ctx->stream));
      int distinct = 0;
      for (uint32_t w : bits) distinct += __builtin_popcount(w);
      /* mean number of distinct tiles per window vs what 2048 uniformly random samples would touch */
      const double mean = (double)distinct / hmrt::kProbeWindows;
      const double random_expect = n_tiles * (1.0 - exp(-(double)hmrt::kProbeSamples / n_tiles));
      ctx->probe_verdict = mean > 0.5 * random_expect ? 1 : 0;
    }
    binned = ctx->probe_verdict == 1;
  }
  if (binned)
    return hmrt::scatter_binned_single(ctx, d_records, n, record_len, sp, finest, reinterpret_cast<unsigned long long*>(d_color_keys), first_index);

  const int64_t blocks = (n + hmrt::kScatterThreads - 1) / hmrt::kScatterThreads;
  if (blocks > 0x7fffffffLL) return HMRT_E_SHAPE;
  const size_t smem = (size_t)hmrt::kScatterThreads * record_len;
  const int staged = (smem <= 48 * 1024) && ((reinterpret_cast<uintptr_t>(d_records) & 15) == 0);
  hmrt::scatter_las_kernel<<<(unsigned)blocks, hmrt::kScatterThreads, staged ? smem : 0, ctx->stream>>>(
      d_records, n, record_len, sp, first_index, finest, reinterpret_cast<unsigned long long*>(d_color_keys), staged);
  HMRT_LAUNCHED(ctx);
  return 0;
}

int hmrt_set_scatter_mode(hmrt_ctx* ctx, int mode) {
  if (!ctx || mode < 0 || mode > 2) return HMRT_E_ARG;
  ctx->scatter_mode = mode;
  return 0;
}

int hmrt_scatter_xyz(hmrt_ctx* ctx, const float* d_xyz, int64_t n, const hmrt_las_transform* xf, float* d_pyramid,
                     int coarse_res, int levels) {
  if (!ctx || !xf || !d_pyramid || n < 0) return HMRT_E_ARG;
  if (n > 0 && !d_xyz) return HMRT_E_ARG;
  int res[HMRT_MAX_LEVELS];
  int64_t idx[HMRT_MAX_LEVELS];
  int rc = hmrt::pyramid_layout(coarse_res, levels, res, idx, nullptr);
  if (rc) return rc;
  if (n == 0) return 0;
  hmrt::ScatterParams sp;
  rc = hmrt::fill_scatter_params(xf, res[0], sp);
  if (rc) return rc;
  hmrt::DeviceGuard guard(ctx->device);
  const int64_t blocks = (n + hmrt::kScatterThreads - 1) / hmrt::kScatterThreads;
  if (blocks > 0x7fffffffLL) return HMRT_E_SHAPE;
  hmrt::scatter_xyz_kernel<<<(unsigned)blocks, hmrt::kScatterThreads, 0, ctx->stream>>>(
      d_xyz, n, sp, reinterpret_cast<int*>(d_pyramid + idx[0]));
  HMRT_LAUNCHED(ctx);
  return 0;
}

int hmrt_build_mips(hmrt_ctx* ctx, float* d_pyramid, int coarse_res, int levels) {
  if (!ctx || !d_pyramid) return HMRT_E_ARG;
  int res[HMRT_MAX_LEVELS];
  int64_t idx[HMRT_MAX_LEVELS];
  int rc = hmrt::pyramid_layout(coarse_res, levels, res, idx, nullptr);
  if (rc) return rc;
  if (levels == 1) return 0;
  hmrt::DeviceGuard guard(ctx->device);
  int done = 0; /* coarser levels already built */
  const bool aligned = (reinterpret_cast<uintptr_t>(d_pyramid + idx[0]) & 15) == 0;
  if (res[0] % 128 == 0 && aligned) {
    hmrt::MipParams mp;
    memset(&mp, 0, sizeof(mp));
    mp.pyramid = d_pyramid;
    mp.res0 = res[0];
    mp.out_levels = levels - 1 < 7 ? levels - 1 : 7;
    for (int i = 0; i <= mp.out_levels; ++i) mp.idx[i] = idx[i];
    /* float2 stores into level 1 need 8-byte alignment of its base: idx[1] is a sum of even squares
     * for every level below the top, but the top level's own square may be odd */
    const bool l1_aligned = (reinterpret_cast<uintptr_t>(d_pyramid + idx[1]) & 7) == 0;
    if (l1_aligned) {
      const dim3 grid(res[0] / 128, res[0] / 128);
      hmrt::build_mips_fused_kernel<<<grid, 512, 0, ctx->stream>>>(mp);
      HMRT_LAUNCHED(ctx);
      done = mp.out_levels;
    }
  }
  for (int l = done + 1; l < levels; ++l) {
    const dim3 block(32, 8);
    const dim3 grid((res[l] + 31) / 32, (res[l] + 7) / 8);
    hmrt::build_mip_level_kernel<<<grid, block, 0, ctx->stream>>>(d_pyramid + idx[l - 1], d_pyramid + idx[l], res[l]);
    HMRT_LAUNCHED(ctx);
  }
  return 0;
}

int hmrt_resolve_colors(hmrt_ctx* ctx, const uint64_t* d_color_keys, hmrt_color* d_color_map, int64_t n_cells) {
  if (!ctx || !d_color_keys || !d_color_map || n_cells < 0) return HMRT_E_ARG;
  if (n_cells == 0) return 0;
  if (reinterpret_cast<uintptr_t>(d_color_map) & 3) return HMRT_E_ARG;
  hmrt::DeviceGuard guard(ctx->device);
  const int64_t quads = (n_cells + 3) / 4;
  const int64_t blocks = (quads + 255) / 256;
  if (blocks > 0x7fffffffLL) return HMRT_E_SHAPE;
  hmrt::resolve_colors_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(
      reinterpret_cast<const unsigned long long*>(d_color_keys), reinterpret_cast<uint8_t*>(d_color_map), n_cells);
  HMRT_LAUNCHED(ctx);
  return 0;
}

}  // extern "C"
