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
#include <math.h>
#include <string.h>

#include "hmrt_internal.cuh"

namespace hmrt {

constexpr int kScatterThreads = 256;

struct ScatterParams {
  double scale[3], offset[3], mn[3];
  float cell[3];
  float rcell[3]; /* 1 / cell when cell is a power of two (x / 2^k == x * 2^-k bit for bit, one rounding of the same value), else 0 */
  float origin[2];
  int res0;
  int cls_off; /* byte offset of the classification byte, -1: none */
  int rgb_off; /* byte offset of R (u16 x 3), -1: none */
};

/* CudaSpace::Color(unsigned short...) : floor(c / 65535.f * 255.f)  (CudaKernel.cuh:41-46) */
__device__ __forceinline__ uint32_t color16(uint32_t c) {
  return (uint32_t)__float2int_rz(floorf(__fmul_rn(__fdiv_rn((float)c, 65535.0f), 255.0f))) & 0xffu;
}

/* x / cell; a power-of-two cell size (launch-uniform) makes it one multiplication */
__device__ __forceinline__ float div_cell(float x, float cell, float rcell) {
  return rcell != 0.0f ? __fmul_rn(x, rcell) : __fdiv_rn(x, cell);
}

/* The first 16 bytes of a LAS point record -- X, Y, Z (int32 LE), intensity, flags, classification -- from five aligned
 * 32-bit loads and funnel shifts, whatever the record's alignment (26- and 34-byte records alternate between 0 and 2 mod 4):
 * 5 loads instead of 13 byte loads and their shifts.  Reads at most the 3 bytes in front of the record inside its first
 * aligned word and never past byte 19 of the record (record_len >= 20). */
struct RecordHead {
  int32_t x, y, z;
  uint32_t tail; /* intensity | flags << 16 | classification << 24 */
};
__device__ __forceinline__ RecordHead load_record_head(const uint8_t* rec) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(rec);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(a & 3) * 8u;
  const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
  RecordHead h;
  h.x = (int32_t)__funnelshift_r(w0, w1, sh);
  h.y = (int32_t)__funnelshift_r(w1, w2, sh);
  h.z = (int32_t)__funnelshift_r(w2, w3, sh);
  h.tail = __funnelshift_r(w3, w4, sh);
  return h;
}

/* main.cpp:200-209 for one decoded point (gx, gy, gz = liblas Point::GetX/Y/Z in double): finest cell and
 * height, or false when the point is rejected (outside the section, class 7). */
__device__ __forceinline__ bool point_to_cell(const ScatterParams& sp, double gx, double gy, double gz, int cls, uint32_t& cell,
                                              float& fZ, uint32_t* cx_out = nullptr, uint32_t* cy_out = nullptr) {
  const float fX = div_cell(__double2float_rn(__dsub_rn(gx, sp.mn[0])), sp.cell[0], sp.rcell[0]); /* :200 */
  const float fY = div_cell(__double2float_rn(__dsub_rn(gy, sp.mn[1])), sp.cell[1], sp.rcell[1]); /* :201 */
  fZ = div_cell(__double2float_rn(__dsub_rn(gz, sp.mn[2])), sp.cell[2], sp.rcell[2]);             /* :202 */
  const float dx = floorf(__fsub_rn(fX, sp.origin[0]));                               /* :205 */
  const float dy = floorf(__fsub_rn(fY, sp.origin[1]));                               /* :206 */
  const float r0 = (float)sp.res0;
  if (!(dx >= 0.0f && dx < r0 && dy >= 0.0f && dy < r0) || cls == 7) return false;    /* :209 */
  const uint32_t cx = (uint32_t)(int)dx, cy = (uint32_t)(int)dy;
  cell = cx + cy * (uint32_t)sp.res0;
  if (cx_out) *cx_out = cx, *cy_out = cy;
  return true;
}

/* main.cpp:223-233 at level 0 */
__device__ __forceinline__ void bin_point(const ScatterParams& sp, double gx, double gy, double gz, int cls,
                                          uint32_t rgb, int64_t file_index, int* __restrict__ finest,
                                          unsigned long long* __restrict__ keys) {
  uint32_t cell;
  float fZ;
  if (!point_to_cell(sp, gx, gy, gz, cls, cell, fZ)) return;
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
  const double gx = __dadd_rn(__dmul_rn((double)rh.x, sp.scale[0]), sp.offset[0]);
  const double gy = __dadd_rn(__dmul_rn((double)rh.y, sp.scale[1]), sp.offset[1]);
  const double gz = __dadd_rn(__dmul_rn((double)rh.z, sp.scale[2]), sp.offset[2]);
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

/* ---------------------------------------------------------------------------------------------
 * Binned scatter for point clouds WITHOUT spatial order (BASELINE config 4: uniformly random points).
 * A direct RED.MAX per point touches a random 32-byte sector of a 1 GiB grid: ~22 G points/s, bound by
 * DRAM random access.  Instead:
 *   pass 1  bin_points_kernel   decode each point once and append (cell, height bits) to the bucket of its
 *                               1024 x 1024-cell tile.  Every CTA owns a private slice of every bucket
 *                               and keeps its fill counters in shared memory, so an append is one
 *                               shared-memory atomic + one 8-byte store: no global atomics, no barriers;
 *   pass 2  apply_bins_kernel   CTAs walk the buckets tile by tile, so the 4 MB of grid a tile covers is
 *                               L2-resident while its atomics are applied.
 * Slices have a fixed capacity of 2x the mean; a point that does not fit falls back to the direct atomic
 * (max is order-independent, so the result is bit-identical either way).
 */
constexpr int kBinThreads = 256, kBinPerThread = 8, kBinChunk = kBinThreads * kBinPerThread;
constexpr int kTileShift = 10;      /* 1024 x 1024 cells = 4 MB of the finest level */
constexpr int kSlicesPerApplyCta = 8;

struct BinParams {
  ScatterParams sp;
  const uint8_t* records;
  int64_t n;
  int record_len;
  uint2* pairs;        /* [n_tiles][n_ctas][slice_cap] (cell, height bits) */
  uint32_t* counts;    /* [n_tiles][n_ctas] */
  uint32_t slice_cap;
  int tiles_x, n_tiles;
  int* finest;
};

/*
 * Pass 1.  Per step of 2048 points the CTA counting-sorts its (cell, height) pairs by tile in shared
 * memory and then writes them out in sorted order, so that the lanes of a warp store to a few contiguous
 * runs (one per tile) instead of 32 unrelated 8-byte slots: ~5x fewer L2 write transactions than one
 * scattered store per point (ncu r01: the scattered version sat in lg_throttle).
 */
__global__ void __launch_bounds__(kBinThreads) bin_points_kernel(const __grid_constant__ BinParams p) {
  extern __shared__ __align__(16) uint32_t bin_smem[];
  uint32_t* fill = bin_smem;            /* [n_tiles] entries used in this CTA's slice of each bucket (persistent) */
  uint32_t* hist = fill + p.n_tiles;    /* [n_tiles] points of this step per tile */
  uint32_t* offs = hist + p.n_tiles;    /* [n_tiles] exclusive prefix of hist */
  uint32_t* dest0 = offs + p.n_tiles;   /* [n_tiles] first destination index of this step's run, or ~0u: slice full */
  uint32_t* sdest = dest0 + p.n_tiles;  /* [kBinChunk] destination pair index per sorted slot */
  uint2* spair = reinterpret_cast<uint2*>(sdest + kBinChunk); /* [kBinChunk] sorted pairs */
  __shared__ uint32_t total_s;
  for (int t = threadIdx.x; t < p.n_tiles; t += kBinThreads) fill[t] = 0, hist[t] = 0;
  __syncthreads();
  const int64_t n_chunks = (p.n + kBinChunk - 1) / kBinChunk;
  const uint32_t n_ctas = gridDim.x;
  for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    /* A: decode, rank inside the tile */
    uint32_t cell[kBinPerThread], hb[kBinPerThread], slot[kBinPerThread];
#pragma unroll
    for (int k = 0; k < kBinPerThread; ++k) {
      const int64_t i = chunk * kBinChunk + k * kBinThreads + threadIdx.x;
      slot[k] = 0xffffffffu;
      if (i < p.n) {
        const uint8_t* rec = p.records + i * p.record_len;
        const RecordHead rh = load_record_head(rec);
        const double gx = __dadd_rn(__dmul_rn((double)rh.x, p.sp.scale[0]), p.sp.offset[0]);
        const double gy = __dadd_rn(__dmul_rn((double)rh.y, p.sp.scale[1]), p.sp.offset[1]);
        const double gz = __dadd_rn(__dmul_rn((double)rh.z, p.sp.scale[2]), p.sp.offset[2]);
        const int cls = (int)((rh.tail >> 24) & 0x1f);
        uint32_t cx, cy;
        float fZ;
        if (point_to_cell(p.sp, gx, gy, gz, cls, cell[k], fZ, &cx, &cy) && fZ >= 0.0f) {
          hb[k] = __float_as_uint(fZ);
          const uint32_t tile = (cy >> kTileShift) * (uint32_t)p.tiles_x + (cx >> kTileShift);
          slot[k] = (tile << 16) | atomicAdd(&hist[tile], 1u); /* rank < 2048 */
        }
      }
    }
    __syncthreads();
    /* B: exclusive prefix over the tiles (warp 0), destinations, slice bookkeeping */
    if (threadIdx.x < 32) {
      const int per = (p.n_tiles + 31) / 32;
      uint32_t sum = 0;
      for (int q = 0; q < per; ++q) {
        const int t = threadIdx.x * per + q;
        if (t < p.n_tiles) sum += hist[t];
      }
      uint32_t incl = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if ((int)threadIdx.x >= d) incl += v;
      }
      uint32_t run = incl - sum;
      for (int q = 0; q < per; ++q) {
        const int t = threadIdx.x * per + q;
        if (t < p.n_tiles) {
          const uint32_t h = hist[t], f = fill[t];
          offs[t] = run;
          run += h;
          if (f + h <= p.slice_cap) {
            dest0[t] = (uint32_t)(((size_t)t * n_ctas + blockIdx.x) * p.slice_cap + f); /* < 2^32: checked by the launcher */
            fill[t] = f + h;
          } else {
            dest0[t] = 0xffffffffu; /* slice full: this step's points of the tile go direct */
          }
          hist[t] = 0;
        }
      }
      if (threadIdx.x == 31) total_s = incl;
    }
    __syncthreads();
    /* C: scatter into sorted order (shared memory) */
#pragma unroll
    for (int k = 0; k < kBinPerThread; ++k) {
      if (slot[k] == 0xffffffffu) continue;
      const uint32_t tile = slot[k] >> 16, rank = slot[k] & 0xffffu;
      const uint32_t j = offs[tile] + rank, d0 = dest0[tile];
      spair[j] = make_uint2(cell[k], hb[k]);
      sdest[j] = d0 == 0xffffffffu ? d0 : d0 + rank;
    }
    __syncthreads();
    /* D: write out; consecutive lanes hit consecutive addresses inside a tile's run */
    const uint32_t total = total_s;
    for (uint32_t j = threadIdx.x; j < total; j += kBinThreads) {
      const uint32_t d = sdest[j];
      const uint2 v = spair[j];
      if (d != 0xffffffffu)
        p.pairs[d] = v;
      else
        atomicMax(p.finest + v.x, (int)v.y);
    }
    __syncthreads();
  }
  for (int t = threadIdx.x; t < p.n_tiles; t += kBinThreads) p.counts[(size_t)t * n_ctas + blockIdx.x] = fill[t];
}

/* Pass 2.  Every thread keeps four 16-byte loads (eight pairs) in flight before it issues their RED.MAXes: with one 8-byte
 * load per thread and trip the kernel sat at 29 % of the DRAM rate waiting on its own loads (ncu r01: 347 warps stalled on
 * long_scoreboard per issue).  Slices start 32-byte aligned (slice_cap % 4 == 0). */
__global__ void __launch_bounds__(kBinThreads) apply_bins_kernel(const uint2* __restrict__ pairs, const uint32_t* __restrict__ counts,
                                                                  uint32_t slice_cap, uint32_t n_ctas, uint32_t groups_per_tile,
                                                                  int* __restrict__ finest) {
  const uint32_t tile = blockIdx.x / groups_per_tile, g = blockIdx.x % groups_per_tile;
  for (uint32_t s = g * kSlicesPerApplyCta; s < min(n_ctas, (g + 1) * kSlicesPerApplyCta); ++s) {
    const uint32_t count = __ldg(counts + (size_t)tile * n_ctas + s);
    const uint2* src = pairs + ((size_t)tile * n_ctas + s) * slice_cap;
    const uint4* src4 = reinterpret_cast<const uint4*>(src);
    const uint32_t n4 = count >> 1; /* whole 16-byte pieces */
    for (uint32_t i = threadIdx.x; i < n4; i += kBinThreads * 4) {
      uint4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t j = i + (uint32_t)k * kBinThreads;
        v[k] = j < n4 ? __ldcs(src4 + j) : make_uint4(0u, 0u, 0u, 0u); /* (cell 0, +0.0f) is a no-op: heights are >= +0 */
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (i + (uint32_t)k * kBinThreads < n4) {
          atomicMax(finest + v[k].x, (int)v[k].y);
          atomicMax(finest + v[k].z, (int)v[k].w);
        }
      }
    }
    if ((count & 1u) && threadIdx.x == 0) {
      const uint2 v = __ldcs(src + (count - 1));
      atomicMax(finest + v.x, (int)v.y);
    }
  }
}

/* How spatially ordered is the input?  Eight windows of 65536 consecutive records are sampled (2048 records
 * each) and every sample sets the bit of its tile in the window's bitmap.  A survey-ordered file touches a
 * handful of tiles per window (direct atomics stay L2-resident), an unordered cloud touches most of them. */
constexpr int kProbeWindows = 8, kProbeSamples = 2048, kProbeWords = 128; /* 4096 tile bits per window */
__global__ void __launch_bounds__(256) locality_probe_kernel(const uint8_t* __restrict__ records, int64_t n, int record_len,
                                                             const __grid_constant__ ScatterParams sp, int tiles_x, uint32_t* out) {
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
  const uint32_t tile = ((cell / (uint32_t)sp.res0) >> kTileShift) * (uint32_t)tiles_x + ((cell % (uint32_t)sp.res0) >> kTileShift);
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
struct MipParams {
  float* pyramid;
  int64_t idx[8]; /* float offsets of levels 0..7 (unused entries = 0) */
  int res0;
  int out_levels; /* how many coarser levels to write: min(levels - 1, 7) */
};

__device__ __forceinline__ float max4(float a, float b, float c, float d) { return fmaxf(fmaxf(a, b), fmaxf(c, d)); }

__global__ void __launch_bounds__(512) build_mips_fused_kernel(const __grid_constant__ MipParams mp) {
  __shared__ float s3[16][17];
  __shared__ float s4[8][9];
  __shared__ float s5[4][5];
  __shared__ float s6[2][3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tile_x = blockIdx.x, tile_z = blockIdx.y;
  const int R0 = mp.res0;
  const float* l0 = mp.pyramid + mp.idx[0];
  const int z0 = tile_z * 128 + warp * 8, x0 = tile_x * 128 + lane * 4;

  float4 a[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) a[r] = __ldg(reinterpret_cast<const float4*>(l0 + (size_t)(z0 + r) * R0 + x0));

  float m2[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) { /* two groups of 4 rows */
    float m1[2][2];
#pragma unroll
    for (int rp = 0; rp < 2; ++rp) {
      const float4 u = a[h * 4 + rp * 2], v = a[h * 4 + rp * 2 + 1];
      m1[rp][0] = max4(u.x, u.y, v.x, v.y);
      m1[rp][1] = max4(u.z, u.w, v.z, v.w);
      if (mp.out_levels >= 1) {
        float* l1 = mp.pyramid + mp.idx[1];
        const int z1 = tile_z * 64 + warp * 4 + h * 2 + rp, x1 = tile_x * 64 + lane * 2;
        *reinterpret_cast<float2*>(l1 + (size_t)z1 * (R0 >> 1) + x1) = make_float2(m1[rp][0], m1[rp][1]);
      }
    }
    m2[h] = max4(m1[0][0], m1[0][1], m1[1][0], m1[1][1]);
    if (mp.out_levels >= 2) {
      float* l2 = mp.pyramid + mp.idx[2];
      const int z2 = tile_z * 32 + warp * 2 + h, x2 = tile_x * 32 + lane;
      l2[(size_t)z2 * (R0 >> 2) + x2] = m2[h];
    }
  }
  if (mp.out_levels < 3) return;
  float m3 = fmaxf(m2[0], m2[1]);
  m3 = fmaxf(m3, __shfl_xor_sync(0xffffffffu, m3, 1));
  if ((lane & 1) == 0) {
    float* l3 = mp.pyramid + mp.idx[3];
    const int z3 = tile_z * 16 + warp, x3 = tile_x * 16 + (lane >> 1);
    l3[(size_t)z3 * (R0 >> 3) + x3] = m3;
    s3[warp][lane >> 1] = m3;
  }
  if (mp.out_levels < 4) return;
  __syncthreads();
  const int t = threadIdx.x;
  if (t < 64) {
    const int z = t >> 3, x = t & 7;
    const float m = max4(s3[2 * z][2 * x], s3[2 * z][2 * x + 1], s3[2 * z + 1][2 * x], s3[2 * z + 1][2 * x + 1]);
    s4[z][x] = m;
    (mp.pyramid + mp.idx[4])[(size_t)(tile_z * 8 + z) * (R0 >> 4) + tile_x * 8 + x] = m;
  }
  if (mp.out_levels < 5) return;
  __syncthreads();
  if (t < 16) {
    const int z = t >> 2, x = t & 3;
    const float m = max4(s4[2 * z][2 * x], s4[2 * z][2 * x + 1], s4[2 * z + 1][2 * x], s4[2 * z + 1][2 * x + 1]);
    s5[z][x] = m;
    (mp.pyramid + mp.idx[5])[(size_t)(tile_z * 4 + z) * (R0 >> 5) + tile_x * 4 + x] = m;
  }
  if (mp.out_levels < 6) return;
  __syncthreads();
  if (t < 4) {
    const int z = t >> 1, x = t & 1;
    const float m = max4(s5[2 * z][2 * x], s5[2 * z][2 * x + 1], s5[2 * z + 1][2 * x], s5[2 * z + 1][2 * x + 1]);
    s6[z][x] = m;
    (mp.pyramid + mp.idx[6])[(size_t)(tile_z * 2 + z) * (R0 >> 6) + tile_x * 2 + x] = m;
  }
  if (mp.out_levels < 7) return;
  __syncthreads();
  if (t == 0)
    (mp.pyramid + mp.idx[7])[(size_t)tile_z * (R0 >> 7) + tile_x] = max4(s6[0][0], s6[0][1], s6[1][0], s6[1][1]);
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

static int fill_scatter_params(const hmrt_las_transform* xf, int res0, ScatterParams& sp) {
  for (int i = 0; i < 3; ++i) {
    sp.scale[i] = xf->scale[i];
    sp.offset[i] = xf->offset[i];
    sp.mn[i] = xf->min[i];
    sp.cell[i] = xf->cell_size[i];
    if (!(xf->cell_size[i] > 0.0f)) return HMRT_E_ARG;
    int e = 0;
    const bool pow2 = frexpf(xf->cell_size[i], &e) == 0.5f && e > -100 && e < 100; /* the reference's default is 2.0 (main.cpp:62) */
    sp.rcell[i] = pow2 ? 1.0f / xf->cell_size[i] : 0.0f;
  }
  sp.origin[0] = xf->origin[0];
  sp.origin[1] = xf->origin[1];
  sp.res0 = res0;
  sp.cls_off = -1;
  sp.rgb_off = -1;
  return 0;
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
  static const int min_len[4] = {20, 28, 26, 34};
  static const int rgb_off[4] = {-1, -1, 20, 28};
  if (!ctx || !xf || !d_pyramid || n < 0 || first_index < 0) return HMRT_E_ARG;
  if (point_format < 0 || point_format > 3 || record_len < min_len[point_format] || record_len > 65535)
    return HMRT_E_ARG;
  if (n > 0 && !d_records) return HMRT_E_ARG;
  if (first_index + n >= ((int64_t)1 << 40)) return HMRT_E_ARG; /* key = index << 24 | rgb */
  int res[HMRT_MAX_LEVELS];
  int64_t idx[HMRT_MAX_LEVELS];
  int rc = hmrt::pyramid_layout(coarse_res, levels, res, idx, nullptr);
  if (rc) return rc;
  if (n == 0) return 0;
  hmrt::ScatterParams sp;
  rc = hmrt::fill_scatter_params(xf, res[0], sp);
  if (rc) return rc;
  sp.cls_off = 15;
  sp.rgb_off = rgb_off[point_format];
  hmrt::DeviceGuard guard(ctx->device);
  int* finest = reinterpret_cast<int*>(d_pyramid + idx[0]);

  /* ---- choose the path: binned for large, spatially unordered clouds on grids beyond L2 ---- */
  const int tiles_x = (res[0] + (1 << hmrt::kTileShift) - 1) >> hmrt::kTileShift;
  const int n_tiles = tiles_x * tiles_x;
  bool binned = false;
  if (!d_color_keys && n_tiles >= 64 && n_tiles <= 4096 && n >= (int64_t)1 << 22 && ctx->scatter_mode != 1) {
    binned = ctx->scatter_mode == 2;
    if (ctx->scatter_mode == 0) {
      const size_t probe_bytes = sizeof(uint32_t) * hmrt::kProbeWindows * hmrt::kProbeWords;
      if (!ctx->d_probe) HMRT_CUDA(cudaMalloc(&ctx->d_probe, probe_bytes));
      HMRT_CUDA(cudaMemsetAsync(ctx->d_probe, 0, probe_bytes, ctx->stream));
      hmrt::locality_probe_kernel<<<hmrt::kProbeWindows * hmrt::kProbeSamples / 256, 256, 0, ctx->stream>>>(d_records, n, record_len, sp,
                                                                                                         tiles_x, ctx->d_probe);
      HMRT_LAUNCHED(ctx);
      uint32_t bits[hmrt::kProbeWindows * hmrt::kProbeWords];
      HMRT_CUDA(cudaMemcpyAsync(bits, ctx->d_probe, probe_bytes, cudaMemcpyDeviceToHost, ctx->stream));
      HMRT_CUDA(cudaStreamSynchronize(ctx->stream));
      int distinct = 0;
      for (uint32_t w : bits) distinct += __builtin_popcount(w);
      /* mean number of distinct tiles per window vs what 2048 uniformly random samples would touch */
      const double mean = (double)distinct / hmrt::kProbeWindows;
      const double random_expect = n_tiles * (1.0 - exp(-(double)hmrt::kProbeSamples / n_tiles));
      binned = mean > 0.5 * random_expect;
    }
  }
  if (binned) {
    /* sub-batches bound the workspace (~16 B per point at 2x mean slice capacity) to ~16 GB */
    const int64_t max_batch = (int64_t)1 << 30;
    for (int64_t first = 0; first < n; first += max_batch) {
      const int64_t nb = n - first < max_batch ? n - first : max_batch;
      const int64_t chunks = (nb + hmrt::kBinChunk - 1) / hmrt::kBinChunk;
      const int64_t n_ctas = chunks < (int64_t)ctx->sm_count * 5 ? chunks : (int64_t)ctx->sm_count * 5;
      int64_t slice_cap = 2 * ((nb + (int64_t)n_tiles * n_ctas - 1) / ((int64_t)n_tiles * n_ctas));
      slice_cap = (slice_cap + 3) / 4 * 4; /* keep slices 32-byte aligned */
      if (slice_cap < 64) slice_cap = 64;
      if ((unsigned long long)n_tiles * (unsigned long long)n_ctas * (unsigned long long)slice_cap >= (1ull << 32)) return HMRT_E_SHAPE;
      const size_t pair_bytes = (size_t)n_tiles * (size_t)n_ctas * (size_t)slice_cap * sizeof(uint2);
      const size_t need = pair_bytes + (size_t)n_tiles * (size_t)n_ctas * sizeof(uint32_t);
      if (ctx->ws_cap < need) {
        if (ctx->d_ws) HMRT_CUDA(cudaFree(ctx->d_ws));
        ctx->d_ws = nullptr;
        ctx->ws_cap = 0;
        HMRT_CUDA(cudaMalloc(&ctx->d_ws, need));
        ctx->ws_cap = need;
      }
      hmrt::BinParams bp;
      bp.sp = sp;
      bp.records = d_records + first * record_len;
      bp.n = nb;
      bp.record_len = record_len;
      bp.pairs = reinterpret_cast<uint2*>(ctx->d_ws);
      bp.counts = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(ctx->d_ws) + pair_bytes);
      bp.slice_cap = (uint32_t)slice_cap;
      bp.tiles_x = tiles_x;
      bp.n_tiles = n_tiles;
      bp.finest = finest;
      const size_t bin_smem = (size_t)n_tiles * 4 * sizeof(uint32_t) + (size_t)hmrt::kBinChunk * (sizeof(uint32_t) + sizeof(uint2));
      hmrt::bin_points_kernel<<<(unsigned)n_ctas, hmrt::kBinThreads, bin_smem, ctx->stream>>>(bp);
      HMRT_LAUNCHED(ctx);
      const uint32_t groups = (uint32_t)((n_ctas + hmrt::kSlicesPerApplyCta - 1) / hmrt::kSlicesPerApplyCta);
      hmrt::apply_bins_kernel<<<(unsigned)n_tiles * groups, hmrt::kBinThreads, 0, ctx->stream>>>(bp.pairs, bp.counts, bp.slice_cap,
                                                                                               (uint32_t)n_ctas, groups, finest);
      HMRT_LAUNCHED(ctx);
    }
    return 0;
  }

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
