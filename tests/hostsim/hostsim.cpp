/*
 * TEST INFRASTRUCTURE ONLY.  Compiles the product's per-ray arithmetic
 * (gpu-heightmap-raytracer_b200/csrc/ray_core.cuh, the __host__ instance of the same template
 * the CUDA kernels instantiate) with plain g++, so that CPU-only tests can check the kernel's
 * formulation against the oracle without a GPU.  It is never loaded by the product package and
 * is not a fallback: hmrt/_abi.py only ever loads csrc/libhmrt.so.
 */
#include <thread>
#include <vector>

#include "../../gpu-heightmap-raytracer_b200/csrc/ray_core.cuh"

extern "C" int hostsim_trace(const float* pyramid, const hmrt_color* color_map, int coarse_res, int levels, int W,
                             int H, const hmrt_camera* cam, const hmrt_trace_opts* opts, int n_threads,
                             int row_begin, int row_end, uint8_t* rgb, hmrt_hit* hits) {
  if (!pyramid || !cam || !opts || !rgb) return -1;
  hmrt::Grid g;
  g.pyramid = pyramid;
  g.color_map = reinterpret_cast<const uint8_t*>(color_map);
  g.coarse_res = coarse_res;
  g.coarse_sq = (uint32_t)coarse_res * (uint32_t)coarse_res;
  g.levels = levels;
  g.res0 = coarse_res << (levels - 1);
  g.extent = (float)coarse_res * (float)(1 << (levels - 1));
  hmrt::Shading sh;
  sh.max_height = opts->max_height;
  sh.use_color_map = opts->use_color_map;
  sh.shadows = opts->shadows;
  for (int i = 0; i < 3; ++i) sh.light[i] = opts->light_dir[i];
  sh.bias = opts->shadow_bias > 0.f ? opts->shadow_bias : 0.0625f;
  hmrt::FrameConsts f;
  hmrt::make_frame_consts(*cam, f);
  if (row_begin < 0) row_begin = 0;
  if (row_end > H) row_end = H;
  if (n_threads < 1) n_threads = 1;
  std::vector<std::thread> pool;
  for (int t = 0; t < n_threads; ++t)
    pool.emplace_back([=] {
      for (int py = row_begin + t; py < row_end; py += n_threads)
        for (int px = 0; px < W; ++px) {
          const hmrt::RayResult r = hmrt::trace_pixel(g, sh, f, W, H, px, py);
          const size_t k = (size_t)px + (size_t)py * W;
          rgb[k * 3] = r.r, rgb[k * 3 + 1] = r.g, rgb[k * 3 + 2] = r.b;
          if (hits) {
            hits[k].x = r.pos.x, hits[k].y = r.pos.y, hits[k].z = r.pos.z;
            hits[k].flags = r.flags;
          }
        }
    });
  for (auto& th : pool) th.join();
  return 0;
}
