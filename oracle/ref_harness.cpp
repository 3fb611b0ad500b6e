/*
 * TEST INFRASTRUCTURE ONLY -- host "launcher" for the reference's own device code.
 *
 * build_ref.sh compiles lines 1-286 of /root/reference/GPUHeightmapRaytracer/src/CudaKernel.cu
 * (everything before the <<<>>> host wrappers) into an object file with plain g++; this file is
 * linked against it and calls the reference's functions through their external symbols:
 *   cuda_initializeDeviceVariables (CudaKernel.cu:245-272), cuda_setParameters (:227-240),
 *   cuda_rayTrace (:195-222), castRay (:121-177), viewToGridSpace (:183-190),
 *   cuda_freeDeviceVariables (:277-286).
 * No reference source is copied: the only restated lines are the 8-line pixel preamble of
 * cuda_rayTrace (:204-216) in trace_pixel_glue(), needed to read castRay's by-reference
 * ray_position (hit point) -- and every call checks that glue against the verbatim kernel.
 */
#include "CudaKernel.cuh"  // the reference header (cast-patched temp copy, see build_ref.sh)

#include <atomic>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "hmrt_oracle.h"

thread_local uint3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0};
thread_local dim3 blockDim(1, 1, 1), gridDim(1, 1, 1);

namespace CudaSpace {
extern bool use_color_map;
extern float max_height;
extern int LOD_levels;
extern int* LOD_resolutions;
extern glm::vec3* grid_camera_position;
extern glm::mat3x3* pixel_to_grid_matrix;
extern glm::ivec2* boundary;
void castRay(glm::vec3& ray_position, glm::vec3& ray_direction, Color& result);
glm::vec3 viewToGridSpace(glm::ivec2& pixel_position);
void cuda_rayTrace(unsigned char* color_buffer);
void cuda_setParameters(glm::vec3 frame_dim, glm::vec3 camera_for, glm::vec3 grid_camera_pos,
                        bool use_color, float max_height);
void cuda_initializeDeviceVariables(glm::ivec2 point_buffer_resolution,
                                    glm::ivec2 texture_resolution, float* point_buffer,
                                    CudaSpace::Color* color_map, int LOD_levels, int stride_x,
                                    float max_height);
void cuda_freeDeviceVariables();
}  // namespace CudaSpace

namespace {

std::mutex g_lock;  // the reference keeps its state in namespace-scope globals

struct PixelResult {
  CudaSpace::Color color;
  hmrt_hit hit;
};

/* Pixel preamble of cuda_rayTrace (CudaKernel.cu:204-216) followed by castRay, keeping what
 * castRay leaves in its by-reference arguments.  With shadows: the extension defined in
 * DESIGN.md section 5, built from a second call of the reference's castRay. */
PixelResult trace_pixel_glue(int px, int py, const hmrt_trace_opts& o, float grid_extent) {
  PixelResult r;
  CudaSpace::Color color(static_cast<unsigned char>(200), static_cast<unsigned char>(200),
                         static_cast<unsigned char>(200));
  glm::ivec2 pixel(px, py);
  glm::vec3 dir = *CudaSpace::pixel_to_grid_matrix * CudaSpace::viewToGridSpace(pixel);
  glm::vec3 pos = dir + *CudaSpace::grid_camera_position;
  dir = glm::normalize(dir);
  const glm::vec3 dir0 = dir;

  /* hit detection: a sentinel the height ramp cannot produce (r != g and r != 255) */
  CudaSpace::Color probe(static_cast<unsigned char>(1), static_cast<unsigned char>(2),
                         static_cast<unsigned char>(3));
  CudaSpace::castRay(pos, dir, probe);
  const bool hit = !(probe.r == 1 && probe.g == 2 && probe.b == 3);
  if (hit) color = probe;
  uint32_t flags = (hit ? HMRT_HIT_HIT : 0u) | (dir0.x < 0 ? HMRT_HIT_MIRROR_X : 0u) |
                   (dir0.z < 0 ? HMRT_HIT_MIRROR_Z : 0u);

  if (hit && o.shadows) {
    const float bias = o.shadow_bias > 0.f ? o.shadow_bias : 0.0625f;
    glm::vec3 g = pos;  // un-mirror (same expression shape as CudaKernel.cu:134,145)
    if (dir0.x < 0) g.x = grid_extent - g.x;
    if (dir0.z < 0) g.z = grid_extent - g.z;
    glm::vec3 org;
    org.x = g.x - bias * dir0.x;
    org.y = g.y - bias * dir0.y;
    org.z = g.z - bias * dir0.z;
    if (org.x >= 0.f && org.x < grid_extent && org.z >= 0.f && org.z < grid_extent) {
      glm::vec3 ldir(o.light_dir[0], o.light_dir[1], o.light_dir[2]);
      CudaSpace::Color sprobe(static_cast<unsigned char>(1), static_cast<unsigned char>(2),
                              static_cast<unsigned char>(3));
      CudaSpace::castRay(org, ldir, sprobe);
      if (!(sprobe.r == 1 && sprobe.g == 2 && sprobe.b == 3)) {
        flags |= HMRT_HIT_SHADOWED;
        color.r = color.r >> 1;
        color.g = color.g >> 1;
        color.b = color.b >> 1;
      }
    }
  }
  r.color = color;
  r.hit.x = pos.x;
  r.hit.y = pos.y;
  r.hit.z = pos.z;
  r.hit.flags = flags;
  return r;
}

template <typename F>
void parallel_rows(int row_begin, int row_end, int n_threads, F&& body) {
  if (n_threads < 1) n_threads = 1;
  std::vector<std::thread> pool;
  for (int t = 0; t < n_threads; ++t)
    pool.emplace_back([=, &body] {
      blockDim = dim3(1, 1, 1);
      gridDim = dim3(1, 1, 1);
      threadIdx = {0, 0, 0};
      for (int py = row_begin + t; py < row_end; py += n_threads) body(py);
    });
  for (auto& th : pool) th.join();
}

}  // namespace

extern "C" int hmrt_ref_float_math(void) {
#ifdef HMRT_REF_FLOAT_MATH
  return 1;
#else
  return 0;
#endif
}

extern "C" int hmrt_ref_trace(const float* pyramid, const hmrt_color* color_map, int coarse_res,
                              int levels, int W, int H, const hmrt_camera* cam,
                              const hmrt_trace_opts* opts, int n_threads, int row_begin,
                              int row_end, uint8_t* rgb, hmrt_hit* hits) {
  if (!pyramid || !cam || !opts || !rgb || coarse_res < 1 || levels < 1 || W < 2 || H < 2)
    return -1;
  if (opts->use_color_map && !color_map) return -1;
  if (row_begin < 0) row_begin = 0;
  if (row_end > H) row_end = H;
  std::lock_guard<std::mutex> guard(g_lock);

  int stride_x = 0;  // main.cpp:997-1002
  for (int i = 0, p = 1; i < levels; ++i, p *= 4) stride_x += p;

  CudaSpace::cuda_initializeDeviceVariables(
      glm::ivec2(coarse_res, coarse_res), glm::ivec2(W, H), const_cast<float*>(pyramid),
      reinterpret_cast<CudaSpace::Color*>(const_cast<hmrt_color*>(color_map)), levels, stride_x,
      opts->max_height);
  CudaSpace::cuda_setParameters(
      glm::vec3(cam->frame_dim[0], cam->frame_dim[1], cam->frame_dim[2]),
      glm::vec3(cam->forward[0], cam->forward[1], cam->forward[2]),
      glm::vec3(cam->position[0], cam->position[1], cam->position[2]), opts->use_color_map != 0,
      opts->max_height);
  const float grid_extent = static_cast<float>(CudaSpace::boundary->x);

  /* (1) the reference kernel verbatim, one "thread" per pixel */
  parallel_rows(row_begin, row_end, n_threads, [&](int py) {
    for (int px = 0; px < W; ++px) {
      blockIdx.x = px;
      blockIdx.y = py;
      CudaSpace::cuda_rayTrace(rgb);
    }
  });

  int status = 0;
  if (hits || opts->shadows) {
    /* (2) instrumented pass in height-ramp mode (hit point, flags, shadow term) */
    std::vector<PixelResult> tmp(static_cast<size_t>(W) * (row_end - row_begin));
    CudaSpace::use_color_map = false;
    parallel_rows(row_begin, row_end, n_threads, [&](int py) {
      for (int px = 0; px < W; ++px)
        tmp[static_cast<size_t>(py - row_begin) * W + px] =
            trace_pixel_glue(px, py, *opts, grid_extent);
    });
    CudaSpace::use_color_map = opts->use_color_map != 0;
    std::atomic<int> bad(0);
    for (int py = row_begin; py < row_end; ++py)
      for (int px = 0; px < W; ++px) {
        const PixelResult& r = tmp[static_cast<size_t>(py - row_begin) * W + px];
        uint8_t* out = rgb + (static_cast<size_t>(px) + static_cast<size_t>(py) * W) * 3;
        if (hits) hits[static_cast<size_t>(px) + static_cast<size_t>(py) * W] = r.hit;
        const bool shadowed = (r.hit.flags & HMRT_HIT_SHADOWED) != 0;
        if (!opts->use_color_map) {
          /* glue colour (before shading) must equal the verbatim kernel's */
          const uint8_t er = shadowed ? out[0] >> 1 : out[0];
          const uint8_t eg = shadowed ? out[1] >> 1 : out[1];
          const uint8_t eb = shadowed ? out[2] >> 1 : out[2];
          if (er != r.color.r || eg != r.color.g || eb != r.color.b) bad++;
          out[0] = r.color.r;
          out[1] = r.color.g;
          out[2] = r.color.b;
        } else {
          /* colour-map mode: keep the verbatim kernel's colour, apply the shadow term;
           * a hit in ramp mode must be a non-background pixel unless the map holds 200s */
          if (shadowed) {
            out[0] >>= 1;
            out[1] >>= 1;
            out[2] >>= 1;
          }
        }
      }
    if (bad.load()) status = -100;
  }

  CudaSpace::cuda_freeDeviceVariables();
  delete CudaSpace::boundary;  // never freed by the reference (CudaKernel.cu:260 vs :277-286)
  CudaSpace::boundary = nullptr;
  return status;
}
