/* Shared pieces of the rasterisation kernels (raster.cu: direct scatter + mip build; rasterx.cu: tile-binned scatter and
 * its multi-GPU form).  Reference semantics: the loadLASToSection inner loop, main.cpp:193-234. */
#pragma once
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

/* the same from shared memory (`saddr` = shared-window address of the record): five LDS */
__device__ __forceinline__ RecordHead load_record_head_shared(uint32_t saddr) {
  const uint32_t base = saddr & ~3u, sh = (saddr & 3u) * 8u;
  uint32_t w0, w1, w2, w3, w4;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(base));
  asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(base));
  asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(w2) : "r"(base));
  asm volatile("ld.shared.u32 %0, [%1+12];" : "=r"(w3) : "r"(base));
  asm volatile("ld.shared.u32 %0, [%1+16];" : "=r"(w4) : "r"(base));
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
  /* :205-209.  floor(t) >= 0 <=> t >= 0 and floor(t) < res0 <=> t < res0 for an integer res0, so the range test runs on t
   * itself, and for 0 <= t < 2^23 the significand of t + 2^23 rounded toward -inf is floor(t): no FRND / F2I (XU pipe). */
  const float tx = __fsub_rn(fX, sp.origin[0]), ty = __fsub_rn(fY, sp.origin[1]);
  const float r0 = (float)sp.res0;
  if (!(tx >= 0.0f && tx < r0 && ty >= 0.0f && ty < r0) || cls == 7) return false;
  const uint32_t cx = __float_as_uint(__fadd_rd(tx, 8388608.0f)) & 0x7fffffu;
  const uint32_t cy = __float_as_uint(__fadd_rd(ty, 8388608.0f)) & 0x7fffffu;
  cell = cx + cy * (uint32_t)sp.res0;
  if (cx_out) *cx_out = cx, *cy_out = cy;
  return true;
}

/* (double)int32, exactly, without the XU conversion: 2^52 + 2^31 + x is representable, subtract the bias */
__device__ __forceinline__ double int_to_double(int32_t x) {
  return __dsub_rn(__hiloint2double(0x43300000, (int)((uint32_t)x ^ 0x80000000u)), 4503601774854144.0);
}

/* launch-invariant parameters from the caller's transform; HMRT_E_ARG for a non-positive cell size */
static inline int fill_scatter_params(const hmrt_las_transform* xf, int res0, ScatterParams& sp) {
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


/* LAS 1.2 point data record formats 0-3: minimum record length, byte offset of R (u16 x 3) or -1 */
static const int kLasMinLen[4] = {20, 28, 26, 34};
static const int kLasRgbOff[4] = {-1, -1, 20, 28};

/* ---------------------------------------------------------------------------------------------
 * Max-mipmap build.  CTA = 512 threads = one 128 x 128 tile of the finest level.
 * Warp w owns finest rows [8w, 8w+8), lane l owns columns [4l, 4l+4): eight 16-byte loads per
 * thread (each warp-load is one fully coalesced 512-byte row segment), levels 1 and 2 come
 * straight out of registers, level 3 needs one warp shuffle, levels 4..7 are reduced by the
 * first warps through shared memory.  Every level is written once, nothing is re-read from HBM:
 * traffic = 4*R0^2 read + 4*R0^2/3 written (the algorithmic minimum).
 */
#if defined(__CUDACC__)
struct MipParams {
  float* pyramid;
  int64_t idx[8]; /* float offsets of levels 0..7 (unused entries = 0) */
  int res0;
  int out_levels; /* how many coarser levels to write: min(levels - 1, 7) */
};

__device__ __forceinline__ float max4(float a, float b, float c, float d) { return fmaxf(fmaxf(a, b), fmaxf(c, d)); }

/* One 128 x 128 finest tile (tile_x, tile_z) -> every coarser level.  `src` points at the tile's first cell in a finest
 * level of row pitch `src_pitch` floats -- the local pyramid's own level 0, or the owner's band in PEER memory (multi-GPU
 * rasterisation): with copy_l0 the tile is also stored into the local level 0, which makes this the all-gather of the
 * finest level and the mip build in one pass. */
__device__ __forceinline__ void mips_tile(const MipParams& mp, const float* __restrict__ src, size_t src_pitch, bool copy_l0, int tile_x,
                                          int tile_z) {
  __shared__ float s3[16][17];
  __shared__ float s4[8][9];
  __shared__ float s5[4][5];
  __shared__ float s6[2][3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int R0 = mp.res0;

  HMRT_DCHECK(tile_x >= 0 && tile_z >= 0 && (tile_x + 1) * 128 <= R0 && (tile_z + 1) * 128 <= R0);
  float4 a[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) a[r] = __ldg(reinterpret_cast<const float4*>(src + (size_t)(warp * 8 + r) * src_pitch + lane * 4));
  if (copy_l0) {
    float* l0 = mp.pyramid + mp.idx[0];
    const int z0 = tile_z * 128 + warp * 8, x0 = tile_x * 128 + lane * 4;
#pragma unroll
    for (int r = 0; r < 8; ++r) *reinterpret_cast<float4*>(l0 + (size_t)(z0 + r) * R0 + x0) = a[r];
  }

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

#endif

/* ---- tile-binned path (rasterx.cu), also used by hmrt_scatter_las for unordered clouds ---------------------------- */
int binned_tile_shift(int res0, int world);
/* Bin + apply n records into `finest` (single GPU: overflowing points go straight to the grid). */
int scatter_binned_single(hmrt_ctx* ctx, const uint8_t* d_records, int64_t n, int record_len, const ScatterParams& sp, int* finest,
                          unsigned long long* keys, int64_t first_index);

}  // namespace hmrt
