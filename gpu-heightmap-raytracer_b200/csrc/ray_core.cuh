/*
 * ray_core.cuh -- per-ray arithmetic of the heightfield traversal, shared by every trace kernel.
 *
 * What it computes is fixed by the reference (GPUHeightmapRaytracer/src/CudaKernel.cu):
 *   pixel -> ray        cuda_rayTrace :195-222, viewToGridSpace :183-190
 *   max-mipmap walk     castRay :121-177, calculateExitPointAndEdge :74-91,
 *                       testIntersection :96-114, getPointBufferValue :61-69
 *   colouring           getHeightColorValue :38-56, getColorMapValue :25-33
 * How it computes it is not: there are no device-heap globals, no pow() and no per-level
 * table loads.  Every fp32 operation is written with an explicit round-to-nearest intrinsic
 * (no FMA contraction), in the reference's operation order, so the result is bit-identical to
 * the reference's own code compiled for the host (oracle/_ref) -- see DESIGN.md section 3.
 * pow(2.f, LOD) is an exponent-field constant; x / 2^LOD == x * 2^-LOD exactly.
 *
 * The functions are __host__ __device__ only so that tests/hostsim can run the very same
 * arithmetic on the CPU next to the oracle; the product never executes the host instance.
 */
#pragma once
#include <stdint.h>

#include "../../include/hmrt.h"

#if defined(__CUDACC__)
#define HMRT_HD __host__ __device__ __forceinline__
#else
#define HMRT_HD inline
#include <math.h>
#include <string.h>
#endif

/* Device-side bounds checks of the CHECKED build (make libhmrt_checked.so: -DHMRT_CHECKED; compute-sanitizer is not available
 * on the GPU pool): every computed index of the rasterisation, window and traversal kernels is tested before it is used; a
 * failure prints the expression and traps, so the next CUDA call of the test fails.  The product build compiles them away. */
#if defined(HMRT_CHECKED) && defined(__CUDACC__)
#include <stdio.h>
#define HMRT_DCHECK(cond)                                                                                             \
  do {                                                                                                                \
    if (!(cond)) {                                                                                                    \
      printf("HMRT_DCHECK failed: %s (%s:%d) block %u thread %u\n", #cond, __FILE__, __LINE__, blockIdx.x, threadIdx.x); \
      __trap();                                                                                                       \
    }                                                                                                                 \
  } while (0)
#else
#define HMRT_DCHECK(cond) ((void)0)
#endif

namespace hmrt {

/* ---- exactly-rounded fp32 primitives (never contracted) --------------------------------- */
#if defined(__CUDA_ARCH__)
HMRT_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
HMRT_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
HMRT_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
HMRT_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
HMRT_HD float fsqrt(float a) { return __fsqrt_rn(a); }
HMRT_HD float ffloor(float a) { return floorf(a); }
HMRT_HD float as_float(uint32_t u) { return __uint_as_float(u); }
HMRT_HD int f2i_rz(float a) { return __float2int_rz(a); }
#else
HMRT_HD float fmul(float a, float b) { return a * b; }
HMRT_HD float fadd(float a, float b) { return a + b; }
HMRT_HD float fsub(float a, float b) { return a - b; }
HMRT_HD float fdiv(float a, float b) { return a / b; }
HMRT_HD float fsqrt(float a) { return sqrtf(a); }
HMRT_HD float ffloor(float a) { return floorf(a); }
HMRT_HD float as_float(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
HMRT_HD int f2i_rz(float a) { return (int)a; }
#endif

template <typename T>
HMRT_HD T ld_ro(const T* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

/* ---- launch-invariant description of the heightmap + frame geometry --------------------- */
struct Grid {
  const float* pyramid;       /* borrowed, reference layout (coarsest level first) */
  const uint8_t* color_map;   /* borrowed hmrt_color[res0*res0] or null */
  uint32_t coarse_sq;         /* coarse_res^2 */
  int coarse_res;
  int levels;
  int res0;                   /* finest resolution = boundary (CudaKernel.cu:260) */
  float extent;               /* (float)coarse_res * 2^(levels-1) (CudaKernel.cu:134,145) */
};

/* cuda_setParameters (CudaKernel.cu:227-240) evaluated once per frame on the host. */
struct FrameConsts {
  float m[9];   /* pixel_to_grid_matrix: m[0..2] = u, m[3..5] = v, m[6..8] = w (columns) */
  float cam[3]; /* grid_camera_position */
  float fd[3];  /* frame_dimension */
  float pad;
};

struct Shading {
  float max_height;
  int use_color_map;
  int shadows;
  float light[3];
  float bias;
};

struct Vec3 {
  float x, y, z;
};

struct RayResult {
  uint8_t r, g, b;
  uint32_t flags; /* HMRT_HIT_* | steps << HMRT_HIT_STEPS_SHIFT */
  uint32_t air;   /* how many of those steps the production walk took in its air phase (instrumentation; 0 for the exact walk) */
  Vec3 pos;       /* castRay's by-reference ray_position at return (mirrored space) */
};

/* level offset in floats: idx[l] = coarse^2 * (4^(L-1-l) - 1)/3 (closed form of
 * CudaKernel.cu:253-258); (4^s - 1)/3 is the bit pattern 0b0101..01 with s ones. */
HMRT_HD uint32_t level_offset(const Grid& g, int lod) {
  const int s = g.levels - 1 - lod;
  return g.coarse_sq * (0x55555555u & ((1u << (2 * s)) - 1u));
}

/* glm::normalize (func_geometric.inl:94): v * (1 / sqrt(dot(v, v))) */
HMRT_HD Vec3 normalize3(Vec3 v) {
  const float d = fadd(fadd(fmul(v.x, v.x), fmul(v.y, v.y)), fmul(v.z, v.z));
  const float s = fdiv(1.0f, fsqrt(d));
  Vec3 r = {fmul(v.x, s), fmul(v.y, s), fmul(v.z, s)};
  return r;
}

/* cuda_rayTrace :209-215: ray through pixel (px, py) of a W x H frame */
HMRT_HD void primary_ray(const FrameConsts& f, int W, int H, int px, int py, Vec3& pos, Vec3& dir) {
  /* viewToGridSpace :185-188 */
  const float gx = fsub(fdiv(f.fd[0], 2.0f), fdiv(fmul(f.fd[0], (float)px), (float)(W - 1)));
  const float gy = fadd(fdiv(-f.fd[1], 2.0f), fdiv(fmul(f.fd[1], (float)py), (float)(H - 1)));
  const float gz = -f.fd[2];
  /* mat3 * vec3, glm type_mat3x3.inl:430-433 */
  dir.x = fadd(fadd(fmul(f.m[0], gx), fmul(f.m[3], gy)), fmul(f.m[6], gz));
  dir.y = fadd(fadd(fmul(f.m[1], gx), fmul(f.m[4], gy)), fmul(f.m[7], gz));
  dir.z = fadd(fadd(fmul(f.m[2], gx), fmul(f.m[5], gy)), fmul(f.m[8], gz));
  pos.x = fadd(dir.x, f.cam[0]);
  pos.y = fadd(dir.y, f.cam[1]);
  pos.z = fadd(dir.z, f.cam[2]);
  dir = normalize3(dir);
}

/* getHeightColorValue :38-56; float -> unsigned char as the host build does it (truncate to
 * int, keep the low byte) */
HMRT_HD void height_color(float height, float max_height, uint8_t& r, uint8_t& g, uint8_t& b) {
  height = fdiv(fmul(height, 2.0f), max_height);
  if (height > 1.0f) {
    height = fsub(height, 1.0f);
    r = 255;
    g = (uint8_t)f2i_rz(fsub(255.0f, fmul(height, 255.0f)));
    b = 0;
  } else {
    r = (uint8_t)f2i_rz(fmul(255.0f, height));
    g = r;
    b = (uint8_t)f2i_rz(fsub(255.0f, fmul(height, 255.0f)));
  }
}

/*
 * castRay (CudaKernel.cu:121-177).  pos/dir are in/out like the reference's by-reference
 * arguments (dir comes back mirrored to the positive quadrant).  Returns true when the walk
 * ended with a hit on the finest level; mirror bits go to `flags`, loop iterations are added
 * to `steps`.  SHADE selects whether a colour is produced (primary) or not (shadow segment).
 */
template <bool SHADE>
HMRT_HD bool cast_ray(const Grid& g, const Shading& sh, Vec3& pos, Vec3& dir, uint32_t& flags, uint32_t& steps,
                      uint8_t& cr, uint8_t& cg, uint8_t& cb) {
  const int top = g.levels - 1;
  int lod = top;
  bool mirror_x = false, mirror_z = false;
  if (dir.x < 0.0f) { /* :130-135 */
    mirror_x = true;
    dir.x = -dir.x;
    pos.x = fsub(g.extent, pos.x);
  }
  if (dir.z < 0.0f) { /* :141-146 */
    mirror_z = true;
    dir.z = -dir.z;
    pos.z = fsub(g.extent, pos.z);
  }
  flags |= (mirror_x ? HMRT_HIT_MIRROR_X : 0u) | (mirror_z ? HMRT_HIT_MIRROR_Z : 0u);
  const bool rising = dir.y >= 0.0f; /* :102 */
  const bool up = dir.y > 0.0f;      /* :153 */
  uint32_t n = 0;
  bool hit_finest = false;

  while (pos.x < g.extent && pos.z < g.extent && !(up && pos.y > sh.max_height)) { /* :153 */
    ++n;
    const float c = as_float((uint32_t)(127 + lod) << 23);  /* pow(2.f, LOD) */
    const float ic = as_float((uint32_t)(127 - lod) << 23); /* 1 / pow(2.f, LOD), exact */
    /* calculateExitPointAndEdge :77-90 */
    const float fx = ffloor(fmul(pos.x, ic));
    const float fz = ffloor(fmul(pos.z, ic));
    const float bx = fmul(fadd(fx, 1.0f), c);
    const float bz = fmul(fadd(fz, 1.0f), c);
    const float tx = fdiv(fsub(bx, pos.x), dir.x);
    const float tz = fdiv(fsub(bz, pos.z), dir.z);
    /* getPointBufferValue :61-69 (issued before the divides resolve) */
    int ix = f2i_rz(fx), iz = f2i_rz(fz);
    const int res = g.coarse_res << (top - lod);
    const int ux = mirror_x ? res - 1 - ix : ix;
    const int uz = mirror_z ? res - 1 - iz : iz;
    const float height = ld_ro(g.pyramid + (size_t)(level_offset(g, lod) + (uint32_t)ux + (uint32_t)uz * (uint32_t)res));
    const bool x_first = tx <= tz;
    const float t = x_first ? tx : tz;
    Vec3 ex;
    ex.x = fadd(pos.x, fmul(t, dir.x));
    ex.y = fadd(pos.y, fmul(t, dir.y));
    ex.z = fadd(pos.z, fmul(t, dir.z));
    int edge; /* floor(exit / 2^LOD) of the snapped coordinate == cell index + 1 */
    if (x_first) {
      ex.x = bx;
      edge = ix + 1;
    } else {
      ex.z = bz;
      edge = iz + 1;
    }
    /* testIntersection :102-111 */
    bool hit;
    if (rising) {
      hit = pos.y <= height;
    } else {
      hit = ex.y <= height;
      if (hit) {
        const float q = fdiv(fsub(height, pos.y), dir.y);
        const float adv = (0.0f < q) ? q : 0.0f; /* glm::max(0.f, q) */
        pos.x = fadd(pos.x, fmul(adv, dir.x));
        pos.y = fadd(pos.y, fmul(adv, dir.y));
        pos.z = fadd(pos.z, fmul(adv, dir.z));
      }
    }
    if (hit) { /* :157-169 */
      if (lod > 0) {
        --lod;
      } else {
        hit_finest = true;
        break;
      }
    } else { /* :173-174 */
      const int nl = lod + 1 - (edge & 1);
      lod = nl < top ? nl : top;
      pos = ex;
    }
  }
  steps += n;
  if (SHADE && hit_finest) {
    if (sh.use_color_map) { /* getColorMapValue :25-33 */
      int cx = f2i_rz(ffloor(pos.x)), cz = f2i_rz(ffloor(pos.z));
      if (mirror_x) cx = g.res0 - 1 - cx;
      if (mirror_z) cz = g.res0 - 1 - cz;
      const uint8_t* p = g.color_map + ((size_t)cx + (size_t)cz * (size_t)g.res0) * 3;
      cr = ld_ro(p);
      cg = ld_ro(p + 1);
      cb = ld_ro(p + 2);
    } else {
      height_color(pos.y, sh.max_height, cr, cg, cb);
    }
  }
  return hit_finest;
}

/* One pixel of cuda_rayTrace (:195-222) plus the shadow extension. */
HMRT_HD RayResult trace_pixel(const Grid& g, const Shading& sh, const FrameConsts& f, int W, int H, int px, int py) {
  RayResult out;
  out.r = out.g = out.b = 200; /* :204 */
  uint32_t flags = 0, steps = 0;
  Vec3 pos, dir;
  primary_ray(f, W, H, px, py, pos, dir);
  const Vec3 dir0 = dir;
  bool hit = false;
  /* the reference indexes out of bounds when the ray starts below the grid origin in mirrored
   * space; defined here (and in the oracle restatement) as "background" */
  const float mx = dir.x < 0.0f ? fsub(g.extent, pos.x) : pos.x;
  const float mz = dir.z < 0.0f ? fsub(g.extent, pos.z) : pos.z;
  if (mx < 0.0f || mz < 0.0f) {
    flags = (dir.x < 0.0f ? HMRT_HIT_MIRROR_X : 0u) | (dir.z < 0.0f ? HMRT_HIT_MIRROR_Z : 0u);
    pos.x = mx;
    pos.z = mz;
  } else {
    hit = cast_ray<true>(g, sh, pos, dir, flags, steps, out.r, out.g, out.b);
  }
  if (hit) flags |= HMRT_HIT_HIT;
  if (hit && sh.shadows) {
    Vec3 gp = pos, org;
    if (dir0.x < 0.0f) gp.x = fsub(g.extent, gp.x);
    if (dir0.z < 0.0f) gp.z = fsub(g.extent, gp.z);
    org.x = fsub(gp.x, fmul(sh.bias, dir0.x));
    org.y = fsub(gp.y, fmul(sh.bias, dir0.y));
    org.z = fsub(gp.z, fmul(sh.bias, dir0.z));
    if (org.x >= 0.0f && org.x < g.extent && org.z >= 0.0f && org.z < g.extent) {
      Vec3 ldir = {sh.light[0], sh.light[1], sh.light[2]};
      uint32_t sflags = 0;
      uint8_t d0, d1, d2;
      if (cast_ray<false>(g, sh, org, ldir, sflags, steps, d0, d1, d2)) {
        flags |= HMRT_HIT_SHADOWED;
        out.r >>= 1;
        out.g >>= 1;
        out.b >>= 1;
      }
    }
  }
  out.flags = flags | (steps << HMRT_HIT_STEPS_SHIFT);
  out.air = 0;
  out.pos = pos;
  return out;
}

/* cuda_setParameters (CudaKernel.cu:227-240): the view->grid basis of one frame.  Evaluated once
 * per frame on the HOST by the launcher (the reference spends a <<<1,1>>> launch + device sync on
 * it, :300-301); strict fp32, reference operation order. */
HMRT_HD void make_frame_consts(const hmrt_camera& c, FrameConsts& f) {
  const float upx = 0.f, upy = 100.f, upz = 0.f;
  const float wx = -c.forward[0], wy = -c.forward[1], wz = -c.forward[2]; /* :236 */
  /* cross(up, w), glm func_geometric.inl:80-83 */
  float ux = fsub(fmul(upy, wz), fmul(wy, upz));
  float uy = fsub(fmul(upz, wx), fmul(wz, upx));
  float uz = fsub(fmul(upx, wy), fmul(wx, upy));
  /* normalize: v * (1 / sqrt(dot)) */
  const float d = fadd(fadd(fmul(ux, ux), fmul(uy, uy)), fmul(uz, uz));
  const float s = fdiv(1.0f, fsqrt(d));
  ux = fmul(ux, s);
  uy = fmul(uy, s);
  uz = fmul(uz, s);
  /* v = cross(w, u) :238 */
  const float vx = fsub(fmul(wy, uz), fmul(uy, wz));
  const float vy = fsub(fmul(wz, ux), fmul(uz, wx));
  const float vz = fsub(fmul(wx, uy), fmul(ux, wy));
  f.m[0] = ux, f.m[1] = uy, f.m[2] = uz; /* :239 mat3(u, v, w): columns */
  f.m[3] = vx, f.m[4] = vy, f.m[5] = vz;
  f.m[6] = wx, f.m[7] = wy, f.m[8] = wz;
  for (int i = 0; i < 3; ++i) {
    f.cam[i] = c.position[i];
    f.fd[i] = c.frame_dim[i];
  }
  f.pad = 0.f;
}

/* tile bookkeeping shared by host and device: rows selected by (first, stride) */
HMRT_HD int rows_local(int H, int tile_first, int tile_stride) {
  if (tile_stride <= 0) tile_stride = 1;
  const int n_tiles = (H + HMRT_ROW_TILE - 1) / HMRT_ROW_TILE;
  int rows = 0;
  for (int t = tile_first; t < n_tiles; t += tile_stride) {
    const int r0 = t * HMRT_ROW_TILE;
    rows += (H - r0 < HMRT_ROW_TILE) ? (H - r0) : HMRT_ROW_TILE;
  }
  return rows;
}

}  // namespace hmrt
