/*
 * TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference's hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load the library built from this file, and only as the checker / reported CPU baseline.
 * The product path (gpu-heightmap-raytracer_b200/csrc) never links, loads or calls it.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors,
 * so the ray traversal below is pinned against the reference's own device code compiled for
 * the host (oracle/_ref/libhmrt_ref.so, build_ref.sh) -- tests/test_oracle_vs_ref.py compares
 * colours and castRay's final ray position bit for bit -- and against fixtures generated from
 * that library (tests/golden/, tests/golden/make_golden.py).  The rasteriser restatement
 * (main.cpp:193-234) cannot be pinned the same way: the original needs the libLAS 1.8.0 binary,
 * <Windows.h> and GL, none of which exist here ("parity unpinned" for the LAS decode step; the
 * arithmetic after decode is restated line by line).
 *
 * Arithmetic: fp32 throughout, no FMA contraction (compile with -ffp-contract=off), i.e. the
 * original MSVC 2015 / CUDA 8 meaning of `pow(2.f, int)` and `floor(float)` (see
 * oracle/shim/device_launch_parameters.h).  All paths below are relative to
 * /root/reference/GPUHeightmapRaytracer/src/ unless stated.
 */
#include "hmrt_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  float x, y, z;
} v3;

/* device globals of the reference (CudaKernel.cu:8-20) + tables of :250-258 */
typedef struct {
  const float* point_buffer;
  const hmrt_color* color_map;
  int levels;
  int res[HMRT_MAX_LEVELS];
  int64_t idx[HMRT_MAX_LEVELS];
  int coarse_res;
  int boundary; /* CudaKernel.cu:260 */
  int W, H;
  float frame_dim[3];
  v3 cam;
  float m[3][3]; /* pixel_to_grid_matrix, m[col][row] like glm (CudaKernel.cu:239) */
  int use_color_map;
  float max_height;
} state_t;

/* ---- pyramid tables: main.cpp:995-1003 == CudaKernel.cu:250-258 ------------------------- */
int hmrt_oracle_pyramid_layout(int coarse_res, int levels, int* res, int64_t* idx,
                               int64_t* total) {
  int r[HMRT_MAX_LEVELS];
  int64_t ix[HMRT_MAX_LEVELS];
  if (coarse_res < 1 || levels < 1 || levels > HMRT_MAX_LEVELS) return -1;
  r[levels - 1] = coarse_res; /* main.cpp:995 */
  ix[levels - 1] = 0;         /* main.cpp:996 */
  for (int i = levels - 2; i >= 0; i--) {
    ix[i] = ix[i + 1] + (int64_t)r[i + 1] * r[i + 1]; /* main.cpp:1000 */
    r[i] = r[i + 1] * 2;                              /* main.cpp:1001 */
  }
  for (int i = 0; i < levels; i++) {
    if (res) res[i] = r[i];
    if (idx) idx[i] = ix[i];
  }
  if (total) *total = ix[0] + (int64_t)r[0] * r[0];
  return 0;
}

/* ---- glm pieces with the reference's operation order ----------------------------------- */
static float dot3(v3 a, v3 b) { /* glm/detail/func_geometric.inl:58-59 */
  float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z;
  return tx + ty + tz;
}
static v3 cross3(v3 x, v3 y) { /* func_geometric.inl:80-83 */
  v3 r;
  r.x = x.y * y.z - y.y * x.z;
  r.y = x.z * y.x - y.z * x.x;
  r.z = x.x * y.y - y.x * x.y;
  return r;
}
static v3 normalize3(v3 v) { /* func_geometric.inl:94 + func_exponential.inl:46 */
  float s = 1.0f / sqrtf(dot3(v, v));
  v3 r = {v.x * s, v.y * s, v.z * s};
  return r;
}

/* ---- cuda_setParameters: CudaKernel.cu:227-240 ------------------------------------------ */
static void set_parameters(state_t* s, const hmrt_camera* c, int use_color, float max_height) {
  v3 up = {0.f, 100.f, 0.f};
  v3 w = {-c->forward[0], -c->forward[1], -c->forward[2]}; /* :236 */
  v3 u = normalize3(cross3(up, w));                        /* :237 */
  v3 v = cross3(w, u);                                     /* :238 */
  s->m[0][0] = u.x, s->m[0][1] = u.y, s->m[0][2] = u.z;    /* :239 columns u, v, w */
  s->m[1][0] = v.x, s->m[1][1] = v.y, s->m[1][2] = v.z;
  s->m[2][0] = w.x, s->m[2][1] = w.y, s->m[2][2] = w.z;
  s->frame_dim[0] = c->frame_dim[0];
  s->frame_dim[1] = c->frame_dim[1];
  s->frame_dim[2] = c->frame_dim[2];
  s->cam.x = c->position[0], s->cam.y = c->position[1], s->cam.z = c->position[2];
  s->use_color_map = use_color;
  s->max_height = max_height;
}

/* ---- getPointBufferValue: CudaKernel.cu:61-69 ------------------------------------------- */
static float point_buffer_value(const state_t* s, int posX, int posZ, int mirrorX, int mirrorZ,
                                int LOD) {
  if (mirrorX) posX = s->res[LOD] - 1 - posX;
  if (mirrorZ) posZ = s->res[LOD] - 1 - posZ;
  return s->point_buffer[s->idx[LOD] + posX + (int64_t)posZ * s->res[LOD]];
}

/* ---- getHeightColorValue: CudaKernel.cu:38-56 (float -> unsigned char truncation) ------- */
static hmrt_color height_color(const state_t* s, float height) {
  hmrt_color c;
  height = height * 2 / s->max_height; /* :41 */
  if (height > 1) {
    height -= 1;
    c.r = 255;
    c.g = (uint8_t)(int)(255 - height * 255); /* :46 */
    c.b = 0;
  } else {
    c.r = (uint8_t)(int)(255 * height); /* :51 */
    c.g = c.r;
    c.b = (uint8_t)(int)(255 - height * 255); /* :53 */
  }
  return c;
}

/* ---- getColorMapValue: CudaKernel.cu:25-33 ---------------------------------------------- */
static hmrt_color color_map_value(const state_t* s, int posX, int posZ, int mirrorX,
                                  int mirrorZ) {
  if (mirrorX) posX = s->res[0] - 1 - posX;
  if (mirrorZ) posZ = s->res[0] - 1 - posZ;
  return s->color_map[posX + (int64_t)posZ * s->res[0]];
}

/*
 * castRay: CudaKernel.cu:121-177 with calculateExitPointAndEdge (:74-91) and testIntersection
 * (:96-114) in line.  Returns 1 when it returned from the finest level (:161-167).  pos and dir
 * are in/out exactly like the reference's by-reference arguments.  *steps counts loop
 * iterations (= height fetches).
 */
static int cast_ray(const state_t* s, v3* pos, v3* dir, hmrt_color* result, int* mirror_flags,
                    uint32_t* steps) {
  int mirrorX, mirrorZ, edge, LOD = s->levels - 1;
  v3 ex;
  uint32_t n = 0;

  if (dir->x < 0) { /* :130-135 */
    mirrorX = 1;
    dir->x = -dir->x;
    pos->x = (float)s->coarse_res * powf(2.f, (float)LOD) - pos->x;
  } else
    mirrorX = 0;
  if (dir->z < 0) { /* :141-146 */
    mirrorZ = 1;
    dir->z = -dir->z;
    pos->z = (float)s->coarse_res * powf(2.f, (float)LOD) - pos->z;
  } else
    mirrorZ = 0;
  *mirror_flags = (mirrorX ? HMRT_HIT_MIRROR_X : 0) | (mirrorZ ? HMRT_HIT_MIRROR_Z : 0);

  while (pos->x < (float)s->boundary && pos->z < (float)s->boundary &&
         !(dir->y > 0 && pos->y > s->max_height)) { /* :153 */
    const float c = powf(2.f, (float)LOD);
    float tX, tZ, height;
    int hit;
    n++;
    /* calculateExitPointAndEdge :77-90 */
    tX = ((floorf(pos->x / c) + 1) * c - pos->x) / dir->x;
    tZ = ((floorf(pos->z / c) + 1) * c - pos->z) / dir->z;
    if (tX <= tZ) {
      ex.x = pos->x + tX * dir->x;
      ex.y = pos->y + tX * dir->y;
      ex.z = pos->z + tX * dir->z;
      ex.x = (floorf(pos->x / c) + 1) * c;
      edge = (int)floorf(ex.x / c);
    } else {
      ex.x = pos->x + tZ * dir->x;
      ex.y = pos->y + tZ * dir->y;
      ex.z = pos->z + tZ * dir->z;
      ex.z = (floorf(pos->z / c) + 1) * c;
      edge = (int)floorf(ex.z / c);
    }
    /* testIntersection :101-111 */
    height = point_buffer_value(s, (int)floorf(pos->x / c), (int)floorf(pos->z / c), mirrorX,
                                mirrorZ, LOD);
    if (dir->y >= 0) {
      hit = pos->y <= height;
    } else {
      hit = ex.y <= height;
      if (hit) {
        /* glm::max(0.f, q) = (0.f < q) ? q : 0.f  (:110, glm/detail/func_common.inl) */
        float q = (height - pos->y) / dir->y;
        float t = (0.f < q) ? q : 0.f;
        pos->x += t * dir->x;
        pos->y += t * dir->y;
        pos->z += t * dir->z;
      }
    }
    if (hit) { /* :157-169 */
      if (LOD > 0)
        LOD--;
      else {
        if (s->use_color_map)
          *result = color_map_value(s, (int)floorf(pos->x), (int)floorf(pos->z), mirrorX,
                                    mirrorZ);
        else
          *result = height_color(s, pos->y);
        *steps += n;
        return 1;
      }
    } else { /* :171-175; edge % 2 uses C remainder like the reference */
      int up = LOD + 1 - (edge % 2);
      LOD = up < s->levels - 1 ? up : s->levels - 1;
      *pos = ex;
    }
  }
  *steps += n;
  return 0;
}

/* ---- viewToGridSpace: CudaKernel.cu:183-190 --------------------------------------------- */
static v3 view_to_grid(const state_t* s, int px, int py) {
  v3 r;
  r.x = s->frame_dim[0] / 2.0f - s->frame_dim[0] * (float)px / (float)(s->W - 1);
  r.y = -s->frame_dim[1] / 2.0f + s->frame_dim[1] * (float)py / (float)(s->H - 1);
  r.z = -s->frame_dim[2];
  return r;
}

/* ---- cuda_rayTrace body for one pixel: CudaKernel.cu:195-222 (+ shadow extension) ------- */
static void trace_pixel(const state_t* s, const hmrt_trace_opts* o, int px, int py, uint8_t* rgb,
                        hmrt_hit* hits) {
  hmrt_color color = {200, 200, 200}; /* :204 */
  v3 g = view_to_grid(s, px, py);
  v3 dir, pos, dir0;
  int mflags = 0, hit;
  uint32_t steps = 0, flags;
  /* glm mat3 * vec3, type_mat3x3.inl:430-433 */
  dir.x = s->m[0][0] * g.x + s->m[1][0] * g.y + s->m[2][0] * g.z;
  dir.y = s->m[0][1] * g.x + s->m[1][1] * g.y + s->m[2][1] * g.z;
  dir.z = s->m[0][2] * g.x + s->m[1][2] * g.y + s->m[2][2] * g.z;
  pos.x = dir.x + s->cam.x, pos.y = dir.y + s->cam.y, pos.z = dir.z + s->cam.z; /* :214 */
  dir = normalize3(dir);                                                        /* :215 */
  dir0 = dir;
  /* defined behaviour where the reference reads out of bounds (negative cell index): a ray
   * whose start lies below the grid origin in castRay's mirrored space returns background */
  {
    float ext = (float)s->boundary;
    float mx = dir.x < 0 ? ext - pos.x : pos.x, mz = dir.z < 0 ? ext - pos.z : pos.z;
    if (mx < 0.f || mz < 0.f) {
      hit = 0;
      mflags = (dir.x < 0 ? HMRT_HIT_MIRROR_X : 0) | (dir.z < 0 ? HMRT_HIT_MIRROR_Z : 0);
      if (dir.x < 0) dir.x = -dir.x, pos.x = mx;
      if (dir.z < 0) dir.z = -dir.z, pos.z = mz;
      goto done;
    }
  }
  hit = cast_ray(s, &pos, &dir, &color, &mflags, &steps); /* :216 */

  if (hit && o->shadows) { /* extension: DESIGN.md section 5 */
    const float bias = o->shadow_bias > 0.f ? o->shadow_bias : 0.0625f;
    const float ext = (float)s->boundary;
    v3 gp = pos, org;
    if (dir0.x < 0) gp.x = ext - gp.x;
    if (dir0.z < 0) gp.z = ext - gp.z;
    org.x = gp.x - bias * dir0.x;
    org.y = gp.y - bias * dir0.y;
    org.z = gp.z - bias * dir0.z;
    if (org.x >= 0.f && org.x < ext && org.z >= 0.f && org.z < ext) {
      v3 ldir = {o->light_dir[0], o->light_dir[1], o->light_dir[2]};
      hmrt_color dummy;
      int smf;
      state_t ramp = *s; /* the shadow segment never reads the colour map */
      ramp.use_color_map = 0;
      if (cast_ray(&ramp, &org, &ldir, &dummy, &smf, &steps)) {
        mflags |= HMRT_HIT_SHADOWED;
        color.r >>= 1, color.g >>= 1, color.b >>= 1;
      }
    }
  }
done:
  flags = (hit ? HMRT_HIT_HIT : 0u) | (uint32_t)mflags | (steps << HMRT_HIT_STEPS_SHIFT);
  {
    size_t t = (size_t)px + (size_t)py * s->W; /* :202 */
    rgb[t * 3] = color.r, rgb[t * 3 + 1] = color.g, rgb[t * 3 + 2] = color.b; /* :219-221 */
    if (hits) {
      hits[t].x = pos.x, hits[t].y = pos.y, hits[t].z = pos.z;
      hits[t].flags = flags;
    }
  }
}

typedef struct {
  const state_t* s;
  const hmrt_trace_opts* o;
  int row_begin, row_end, t, nt;
  uint8_t* rgb;
  hmrt_hit* hits;
} job_t;

static void* worker(void* p) {
  job_t* j = (job_t*)p;
  for (int py = j->row_begin + j->t; py < j->row_end; py += j->nt)
    for (int px = 0; px < j->s->W; px++) trace_pixel(j->s, j->o, px, py, j->rgb, j->hits);
  return NULL;
}

int hmrt_oracle_trace(const float* pyramid, const hmrt_color* color_map, int coarse_res,
                      int levels, int W, int H, const hmrt_camera* cam,
                      const hmrt_trace_opts* opts, int n_threads, int row_begin, int row_end,
                      uint8_t* rgb, hmrt_hit* hits) {
  state_t s;
  if (!pyramid || !cam || !opts || !rgb || W < 2 || H < 2) return -1;
  if (opts->use_color_map && !color_map) return -1;
  memset(&s, 0, sizeof s);
  if (hmrt_oracle_pyramid_layout(coarse_res, levels, s.res, s.idx, NULL)) return -1;
  s.point_buffer = pyramid;
  s.color_map = color_map;
  s.levels = levels;
  s.coarse_res = coarse_res;
  s.boundary = s.res[0]; /* CudaKernel.cu:260 */
  s.W = W, s.H = H;
  set_parameters(&s, cam, opts->use_color_map != 0, opts->max_height);
  if (row_begin < 0) row_begin = 0;
  if (row_end > H) row_end = H;
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 1024) n_threads = 1024;
  {
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    job_t* jobs = (job_t*)malloc(sizeof(job_t) * (size_t)n_threads);
    if (!th || !jobs) return -4;
    for (int t = 0; t < n_threads; t++) {
      job_t j = {&s, opts, row_begin, row_end, t, n_threads, rgb, hits};
      jobs[t] = j;
      pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
  }
  return 0;
}

/* ======================= rasterisation: main.cpp:193-234 ================================= */

/* CudaSpace::Color(unsigned short r,g,b): CudaKernel.cuh:41-46 */
static uint8_t color16(uint16_t c) { return (uint8_t)(int)floorf((float)c / 65535.f * 255.f); }

static uint16_t rd_u16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static int32_t rd_i32(const uint8_t* p) {
  return (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) |
                   ((uint32_t)p[3] << 24));
}

/* One point after decode: main.cpp:200-233.  gx/gy/gz are liblas Point::GetX/Y/Z (double). */
static void rasterise_point(double gx, double gy, double gz, int cls, int has_color, uint16_t cr,
                            uint16_t cg, uint16_t cb, const hmrt_las_transform* xf,
                            float* pyramid, const int* res, const int64_t* idx, int levels,
                            hmrt_color* color_map) {
  int64_t index[HMRT_MAX_LEVELS];
  float fX = (float)(gx - xf->min[0]) / xf->cell_size[0]; /* :200 */
  float fY = (float)(gy - xf->min[1]) / xf->cell_size[1]; /* :201 */
  float fZ = (float)(gz - xf->min[2]) / xf->cell_size[2]; /* :202 */
  float dx = floorf(fX - xf->origin[0]), dy = floorf(fY - xf->origin[1]);
  int x, y;
  /* static_cast<int> of an out-of-range float is undefined; reject it the way the range test
   * at :209 would reject any representable value */
  if (!(dx >= 0.f && dx < (float)res[0] && dy >= 0.f && dy < (float)res[0])) return;
  x = (int)dx; /* :205 */
  y = (int)dy; /* :206 */
  if (x < 0 || x >= res[0] || y < 0 || y >= res[0] || cls == 7) return; /* :209 */
  for (int i = levels - 1; i >= 0; i--) /* :213-216 */
    index[i] = idx[i] + x / (1 << i) + (int64_t)(y / (1 << i)) * res[i];
  if (color_map) { /* :223-224, last writer wins */
    hmrt_color c = {0, 0, 0};
    if (has_color) c.r = color16(cr), c.g = color16(cg), c.b = color16(cb);
    color_map[x + (int64_t)y * res[0]] = c;
  }
  for (int i = 0; i < levels; i++) { /* :227-233 */
    if (pyramid[index[i]] <= fZ)
      pyramid[index[i]] = fZ;
    else
      break;
  }
}

int hmrt_oracle_rasterise_las(const uint8_t* records, int64_t n, int record_len,
                              int point_format, const hmrt_las_transform* xf, float* pyramid,
                              int coarse_res, int levels, hmrt_color* color_map) {
  int res[HMRT_MAX_LEVELS];
  int64_t idx[HMRT_MAX_LEVELS];
  /* LAS 1.2 point data record formats 0-3: X,Y,Z int32 at 0,4,8; classification byte at 15;
   * RGB uint16 at 20 (format 2) or 28 (format 3) */
  static const int min_len[4] = {20, 28, 26, 34};
  static const int rgb_off[4] = {-1, -1, 20, 28};
  if (!records || !xf || !pyramid || point_format < 0 || point_format > 3) return -1;
  if (record_len < min_len[point_format]) return -1;
  if (hmrt_oracle_pyramid_layout(coarse_res, levels, res, idx, NULL)) return -1;
  for (int64_t i = 0; i < n; i++) {
    const uint8_t* p = records + i * record_len;
    /* libLAS 1.8.0 Point::GetX(): raw * scale + offset, in double */
    double gx = rd_i32(p) * xf->scale[0] + xf->offset[0];
    double gy = rd_i32(p + 4) * xf->scale[1] + xf->offset[1];
    double gz = rd_i32(p + 8) * xf->scale[2] + xf->offset[2];
    int cls = p[15] & 0x1f; /* liblas Classification::GetClass(): low 5 bits */
    int ro = rgb_off[point_format];
    uint16_t r = 0, g = 0, b = 0;
    if (ro >= 0) r = rd_u16(p + ro), g = rd_u16(p + ro + 2), b = rd_u16(p + ro + 4);
    rasterise_point(gx, gy, gz, cls, ro >= 0, r, g, b, xf, pyramid, res, idx, levels,
                    color_map);
  }
  return 0;
}

int hmrt_oracle_rasterise_xyz(const float* xyz, int64_t n, const hmrt_las_transform* xf,
                              float* pyramid, int coarse_res, int levels) {
  int res[HMRT_MAX_LEVELS];
  int64_t idx[HMRT_MAX_LEVELS];
  if (!xyz || !xf || !pyramid) return -1;
  if (hmrt_oracle_pyramid_layout(coarse_res, levels, res, idx, NULL)) return -1;
  for (int64_t i = 0; i < n; i++)
    rasterise_point((double)xyz[3 * i], (double)xyz[3 * i + 1], (double)xyz[3 * i + 2], 0, 0, 0,
                    0, 0, xf, pyramid, res, idx, levels, NULL);
  return 0;
}

int hmrt_oracle_build_mips(float* pyramid, int coarse_res, int levels) {
  int res[HMRT_MAX_LEVELS];
  int64_t idx[HMRT_MAX_LEVELS];
  if (!pyramid) return -1;
  if (hmrt_oracle_pyramid_layout(coarse_res, levels, res, idx, NULL)) return -1;
  for (int l = 1; l < levels; l++) {
    const float* fine = pyramid + idx[l - 1];
    float* coarse = pyramid + idx[l];
    int rf = res[l - 1], rc = res[l];
    for (int z = 0; z < rc; z++)
      for (int x = 0; x < rc; x++) {
        float a = fine[(int64_t)(2 * z) * rf + 2 * x], b = fine[(int64_t)(2 * z) * rf + 2 * x + 1];
        float c = fine[(int64_t)(2 * z + 1) * rf + 2 * x],
              d = fine[(int64_t)(2 * z + 1) * rf + 2 * x + 1];
        float m = a;
        if (m <= b) m = b; /* same comparison as main.cpp:229 */
        if (m <= c) m = c;
        if (m <= d) m = d;
        coarse[(int64_t)z * rc + x] = m;
      }
  }
  return 0;
}

/* ======================= PointdataGenerator: PointdataGenerator/main.cpp:72-184 ========== */

typedef struct {
  uint64_t s;
} lcg_t;
/* uniform float in [0,1): top 24 bits of a 64-bit LCG (Knuth MMIX constants) */
static float lcg_uniform(lcg_t* g) {
  g->s = g->s * 6364136223846793005ULL + 1442695040888963407ULL;
  return (float)(g->s >> 40) * (1.0f / 16777216.0f);
}

#define GZ(i, j) z[(size_t)(i) * (size_t)G + (size_t)(j)]

static void diamond_step(float* z, int G, int x, int y, int step, float r) { /* PDG:117-130 */
  GZ(x, y) = (GZ(x + step, y + step) + GZ(x - step, y + step) + GZ(x + step, y - step) +
              GZ(x - step, y - step)) /
                 4 +
             r;
}

static void square_step(float* z, int G, int x, int y, int step, float r) { /* PDG:132-158 */
  float left = 0, right = 0, top = 0, bottom = 0;
  int count = 0;
  if (x > 0) left = GZ(x - step, y), count++;
  if (x < G - 1) right = GZ(x + step, y), count++;
  if (y > 0) top = GZ(x, y - step), count++;
  if (y < G - 1) bottom = GZ(x, y + step), count++;
  GZ(x, y) = (left + right + top + bottom) / count + r;
}

int hmrt_oracle_pdg_generate(int n, uint64_t seed, float* out_xyz) {
  const int G = n + 1; /* PDG:10,52 */
  float* z;
  lcg_t gen = {seed};
  int count = 1;
  if (n < 2 || (n & (n - 1)) || !out_xyz) return -1;
  z = (float*)calloc((size_t)G * G, sizeof(float));
  if (!z) return -4;
  GZ(0, 0) = lcg_uniform(&gen); /* PDG:91-94, startDeviation = 1 */
  GZ(0, G - 1) = lcg_uniform(&gen);
  GZ(G - 1, 0) = lcg_uniform(&gen);
  GZ(G - 1, G - 1) = lcg_uniform(&gen);
  for (int i = G; i > 1; i /= 2) { /* PDG:100-112 */
    for (int x = i / 2; x < G; x += i)
      for (int y = i / 2; y < G; y += i) {
        /* pow(dist(gen), count): float, int -> double pow, narrowed to the float parameter */
        float r0 = (float)pow((double)lcg_uniform(&gen), (double)count);
        diamond_step(z, G, x, y, i / 2, r0);
        r0 = (float)pow((double)lcg_uniform(&gen), (double)count);
        square_step(z, G, x, y - i / 2, i / 2, r0);
        r0 = (float)pow((double)lcg_uniform(&gen), (double)count);
        square_step(z, G, x - i / 2, y, i / 2, r0);
        r0 = (float)pow((double)lcg_uniform(&gen), (double)count);
        square_step(z, G, x + i / 2, y, i / 2, r0);
        r0 = (float)pow((double)lcg_uniform(&gen), (double)count);
        square_step(z, G, x, y + i / 2, i / 2, r0);
      }
    ++count;
  }
  for (int i = 0; i < G - 1; i++) /* scaleData, PDG:175-184: last row/column untouched */
    for (int j = 0; j < G - 1; j++) GZ(i, j) *= 10.f;
  for (int i = 0; i < G; i++) /* PDG:86: Point((float)i, (float)j, z) */
    for (int j = 0; j < G; j++) {
      size_t k = ((size_t)i * G + j) * 3;
      out_xyz[k] = (float)i, out_xyz[k + 1] = (float)j, out_xyz[k + 2] = GZ(i, j);
    }
  free(z);
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * The camera window over the section grid: preparePointBuffer, main.cpp:459-618, restated loop by loop.
 * ---------------------------------------------------------------------------------------------- */

int hmrt_oracle_window_place(const float camera_position[3], const float* section_origins, int grid, int coarse_res, int levels,
                             hmrt_window_placement* out) {
  if (!camera_position || !section_origins || !out || grid < 1 || coarse_res < 1 || levels < 1) return -1;
#define ORIGIN(i, j, a) section_origins[((size_t)(i) * grid + (j)) * 2 + (a)]
  /* glm::pow(2.0f, LOD_levels - 1): float under the original toolchain, and exact either way */
  const float top = powf(2.0f, (float)(levels - 1));
  const float off_x = top * (float)coarse_res / 2.0f, off_y = top * (float)coarse_res / 2.0f; /* :465-467 */
  const float bl_x = camera_position[0] - off_x, bl_y = camera_position[2] - off_y;            /* :469 */
  const float tr_x = camera_position[0] + off_x - FLT_MIN, tr_y = camera_position[2] + off_y - FLT_MIN; /* :470 */
  int minX, minY, maxX, maxY;
  /* :472-502.  `minX < grid` is tested first here: the reference reads origins[grid] before it looks at the index. */
  minX = 0;
  while (minX < grid && bl_x > ORIGIN(minX, 0, 0)) minX++;
  minX--;
  minY = 0;
  while (minY < grid && bl_y > ORIGIN(0, minY, 1)) minY++;
  minY--;
  maxX = 0;
  while (maxX < grid && tr_x > ORIGIN(maxX, 0, 0)) maxX++;
  maxX--;
  maxY = 0;
  while (maxY < grid && tr_y > ORIGIN(0, maxY, 1)) maxY++;
  maxY--;
  if (minX < 0 || minY < 0 || maxX < 0 || maxY < 0) return -1; /* the reference would index [-1] */
  const float sp_x = bl_x - ORIGIN(minX, minY, 0), sp_y = bl_y - ORIGIN(minX, minY, 1); /* :509 */
  const int cell_x = (int)floorf(sp_x / top), cell_y = (int)floorf(sp_y / top);       /* :510 */
  if (cell_x < 0 || cell_y < 0 || cell_x >= coarse_res || cell_y >= coarse_res) return -1;
  out->min_x = minX, out->min_y = minY, out->max_x = maxX, out->max_y = maxY;
  out->cell_x = cell_x, out->cell_y = cell_y;
  const float half_top = powf(2.0f, (float)(levels - 2));
  out->camera[0] = (sp_x - cell_x * top) + (coarse_res - 1) * half_top; /* :513-516 */
  out->camera[1] = camera_position[1];
  out->camera[2] = (sp_y - cell_y * top) + (coarse_res - 1) * half_top;
#undef ORIGIN
  return 0;
}

/* sections[x][y]: x 0 = minX, 1 = maxX; y 0 = minY, 1 = maxY (host pointers).  colors may be NULL (then out_colors too). */
int hmrt_oracle_compose_window(const float* const sections[2][2], const hmrt_color* const colors[2][2], int coarse_res, int levels,
                               int cell_x, int cell_y, float* h_point_buffer, hmrt_color* h_color_map) {
  int LOD_resolutions[HMRT_MAX_LEVELS];
  int64_t LOD_indexes[HMRT_MAX_LEVELS];
  if (hmrt_oracle_pyramid_layout(coarse_res, levels, LOD_resolutions, LOD_indexes, NULL)) return -1;
  int cpx = cell_x, cpy = cell_y; /* cell_position */
  int row_index, row_offset;
  for (int i = levels - 1; i >= 0; i--) { /* :519 */
    const int R = LOD_resolutions[i];
    const int64_t I = LOD_indexes[i];
    /* lower left section, :521-529 */
    row_offset = 0;
    for (row_index = cpy; row_index < R; row_index++) {
      memcpy(h_point_buffer + I + (int64_t)row_offset * R, sections[0][0] + I + cpx + (int64_t)row_index * R, sizeof(float) * (size_t)(R - cpx));
      row_offset++;
    }
    /* bottom right section, :531-541 */
    row_offset = 0;
    row_index = cpx == 0 ? R : cpy;
    for (; row_index < R; row_index++) {
      memcpy(h_point_buffer + I + (R - cpx) + (int64_t)row_offset * R, sections[1][0] + I + (int64_t)row_index * R, sizeof(float) * (size_t)cpx);
      row_offset++;
    }
    /* top left section, :543-553 */
    row_offset = 0;
    row_index = cpy == 0 ? cpy : 0;
    for (; row_index < cpy; row_index++) {
      memcpy(h_point_buffer + I + (int64_t)(row_index + R - cpy) * R, sections[0][1] + I + cpx + (int64_t)row_offset * R,
             sizeof(float) * (size_t)(R - cpx));
      row_offset++;
    }
    /* top right section, :555-565 */
    row_offset = 0;
    row_index = (cpy == 0 || cpx == 0) ? cpy : 0;
    for (; row_index < cpy; row_index++) {
      memcpy(h_point_buffer + I + (R - cpx) + (int64_t)(row_index + R - cpy) * R, sections[1][1] + I + (int64_t)row_offset * R,
             sizeof(float) * (size_t)cpx);
      row_offset++;
    }
    if (i > 0) cpx *= 2, cpy *= 2; /* :566-567 */
  }
  if (!h_color_map) return 0;
  if (!colors) return -1;
  const int R = LOD_resolutions[0];
  /* colour data, :570-618 (cell_position is now at the finest level) */
  row_offset = 0;
  for (row_index = cpy; row_index < R; row_index++) {
    memcpy(h_color_map + (int64_t)row_offset * R, colors[0][0] + cpx + (int64_t)row_index * R, sizeof(hmrt_color) * (size_t)(R - cpx));
    row_offset++;
  }
  row_offset = 0;
  row_index = cpx == 0 ? R : cpy;
  for (; row_index < R; row_index++) {
    memcpy(h_color_map + (R - cpx) + (int64_t)row_offset * R, colors[1][0] + (int64_t)row_index * R, sizeof(hmrt_color) * (size_t)cpx);
    row_offset++;
  }
  row_offset = 0;
  row_index = cpy == 0 ? cpy : 0;
  for (; row_index < cpy; row_index++) {
    memcpy(h_color_map + (int64_t)(row_index + R - cpy) * R, colors[0][1] + cpx + (int64_t)row_offset * R, sizeof(hmrt_color) * (size_t)(R - cpx));
    row_offset++;
  }
  row_offset = 0;
  row_index = (cpy == 0 || cpx == 0) ? cpy : 0;
  for (; row_index < cpy; row_index++) {
    memcpy(h_color_map + (R - cpx) + (int64_t)(row_index + R - cpy) * R, colors[1][1] + (int64_t)row_offset * R, sizeof(hmrt_color) * (size_t)cpx);
    row_offset++;
  }
  return 0;
}
