/*
 * CudaKernel.cuh -- drop-in replacement for the reference header of the same name
 * (GPUHeightmapRaytracer/src/CudaKernel.cuh:35-52).
 *
 * The reference's main.cpp includes "CudaKernel.cuh" and calls exactly three functions of
 * namespace CudaSpace plus the Color struct:
 *     CudaSpace::initializeDeviceVariables(...)   main.cpp:1014
 *     CudaSpace::rayTrace(...)                    main.cpp:686
 *     CudaSpace::freeDeviceVariables()            main.cpp:1024
 *     CudaSpace::Color                            main.cpp:66,78,224,260,...
 * This header provides the same names, argument order and meaning, implemented on top of the C ABI
 * of libhmrt.so (include/hmrt.h).  The vector arguments are templates over "anything with .x/.y/.z",
 * so the reference's glm::ivec2 / glm::vec3 objects bind without this header depending on GLM.
 * Semantics kept: borrowed device pointers, synchronous rayTrace (CudaKernel.cu:307), errors are
 * logged to stderr and execution continues (inc/helper_cuda.h:985-995).
 */
#pragma once
#include <climits>
#include <cmath>
#include <cstdio>

#include "../../include/hmrt.h"

namespace CudaSpace {

struct Color { /* CudaKernel.cuh:37-48 */
  Color() : r(0), g(0), b(0) {}
  Color(unsigned char r_, unsigned char g_, unsigned char b_) : r(r_), g(g_), b(b_) {}
  Color(unsigned short r_, unsigned short g_, unsigned short b_) {
    r = (unsigned char)(std::floor(r_ / static_cast<float>(USHRT_MAX) * 255.f));
    g = (unsigned char)(std::floor(g_ / static_cast<float>(USHRT_MAX) * 255.f));
    b = (unsigned char)(std::floor(b_ / static_cast<float>(USHRT_MAX) * 255.f));
  }
  unsigned char r, g, b;
};
static_assert(sizeof(Color) == sizeof(hmrt_color), "Color must stay a packed RGB8 triple");

namespace detail {
struct State {
  hmrt_ctx* ctx = nullptr;
  int width = 0, height = 0; /* texture_resolution captured at init (CudaKernel.cu:266) */
};
inline State& state() {
  static State s;
  return s;
}
inline bool check(int code, const char* what, const char* file, int line) {
  if (code != 0) /* checkCudaErrors: report and carry on */
    std::fprintf(stderr, "CUDA error at %s:%d code=%d(%s) \"%s\" \n", file, line, code, hmrt_error_string(code), what);
  return code == 0;
}
}  // namespace detail
#define HMRT_SHIM_CHECK(call) ::CudaSpace::detail::check((call), #call, __FILE__, __LINE__)

/* CudaKernel.cuh:50 / CudaKernel.cu:313-317.  stride_x is implied by (resolution, LOD_levels). */
template <class IVec2>
inline void initializeDeviceVariables(IVec2& point_buffer_res, IVec2& texture_res, float* d_gpu_pointBuffer,
                                      Color* d_color_map, int LOD_levels, int /*stride_x*/, float max_height) {
  detail::State& s = detail::state();
  if (!s.ctx) {
    int dev = 0;
    if (!HMRT_SHIM_CHECK(hmrt_create(dev, &s.ctx))) return;
  }
  s.width = texture_res.x;
  s.height = texture_res.y;
  HMRT_SHIM_CHECK(hmrt_set_heightmap(s.ctx, d_gpu_pointBuffer, reinterpret_cast<const hmrt_color*>(d_color_map),
                                     point_buffer_res.x, LOD_levels, max_height));
}

/* CudaKernel.cuh:49 / CudaKernel.cu:291-308.  colorBuffer is a DEVICE pointer (the mapped PBO,
 * main.cpp:681-683); the call returns when the frame is complete. */
template <class IVec2, class Vec3>
inline void rayTrace(IVec2& texture_resolution, Vec3& frame_dimensions, Vec3& camera_forward, Vec3& grid_camera_position,
                     unsigned char* colorBuffer, bool use_color, float max_height) {
  detail::State& s = detail::state();
  if (!s.ctx) {
    std::fprintf(stderr, "CudaSpace::rayTrace called before initializeDeviceVariables\n");
    return;
  }
  (void)texture_resolution; /* the kernel uses the resolution captured at init (CudaKernel.cu:266) */
  hmrt_camera cam;
  cam.frame_dim[0] = frame_dimensions.x, cam.frame_dim[1] = frame_dimensions.y, cam.frame_dim[2] = frame_dimensions.z;
  cam.forward[0] = camera_forward.x, cam.forward[1] = camera_forward.y, cam.forward[2] = camera_forward.z;
  cam.position[0] = grid_camera_position.x, cam.position[1] = grid_camera_position.y, cam.position[2] = grid_camera_position.z;
  hmrt_trace_opts opts;
  hmrt_trace_opts_default(&opts, max_height);
  opts.use_color_map = use_color ? 1 : 0;
  if (HMRT_SHIM_CHECK(hmrt_trace(s.ctx, s.width, s.height, &cam, 1, &opts, colorBuffer, nullptr)))
    HMRT_SHIM_CHECK(hmrt_synchronize(s.ctx)); /* cudaDeviceSynchronize, CudaKernel.cu:307 */
}

/* CudaKernel.cuh:51 / CudaKernel.cu:322-326 */
inline void freeDeviceVariables() {
  detail::State& s = detail::state();
  if (!s.ctx) return;
  HMRT_SHIM_CHECK(hmrt_clear_heightmap(s.ctx));
  HMRT_SHIM_CHECK(hmrt_destroy(s.ctx));
  s.ctx = nullptr;
}

}  // namespace CudaSpace
