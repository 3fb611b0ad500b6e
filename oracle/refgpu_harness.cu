/*
 * TEST / BASELINE INFRASTRUCTURE ONLY -- the reference's OWN CUDA kernel recompiled for sm_100a.
 *
 * build_ref.sh compiles this translation unit with nvcc; REF_CU is the path of a temp copy of
 * /root/reference/GPUHeightmapRaytracer/src/CudaKernel.cu (header cast-patched, nothing else changed) that is
 * #included below, so the reference's __device__ globals, cuda_initializeDeviceVariables / cuda_setParameters /
 * cuda_rayTrace / cuda_freeDeviceVariables are used verbatim.  The only difference from the reference's host wrapper
 * (CudaKernel.cu:291-308) is the launch shape: its block (1, H/2) exceeds 1024 threads at H = 2160, so the block
 * height is the largest divisor of H that is <= 512 (results do not depend on the shape).
 * This is the "reference kernel on B200" the new kernel is measured against; note that
 * nvcc-on-Linux gives it the double-pow / FMA-contracted meaning.
 */
#include REF_CU

#include "../include/hmrt.h"

extern "C" int hmrt_refgpu_trace(float* d_pyramid, hmrt_color* d_color_map, int coarse_res, int levels, int W, int H,
                                 const hmrt_camera* cam, int use_color_map, float max_height, unsigned char* d_rgb, int n_repeat,
                                 float* ms_out) {
  if (!d_pyramid || !cam || !d_rgb || W < 2 || H < 2) return -1;
  int stride_x = 0;
  for (int i = 0, p = 1; i < levels; ++i, p *= 4) stride_x += p;
  glm::ivec2 pbr(coarse_res, coarse_res), tex(W, H);
  CudaSpace::initializeDeviceVariables(pbr, tex, d_pyramid, reinterpret_cast<CudaSpace::Color*>(d_color_map), levels, stride_x, max_height);
  int by = 1;
  for (int d = 1; d <= 512 && d <= H; ++d)
    if (H % d == 0) by = d;
  const dim3 block(1, by), grid(W, H / by);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  if (n_repeat < 1) n_repeat = 1;
  cudaEventRecord(e0);
  for (int r = 0; r < n_repeat; ++r) {
    CudaSpace::cuda_setParameters<<<1, 1>>>(glm::vec3(cam->frame_dim[0], cam->frame_dim[1], cam->frame_dim[2]),
                                            glm::vec3(cam->forward[0], cam->forward[1], cam->forward[2]),
                                            glm::vec3(cam->position[0], cam->position[1], cam->position[2]), use_color_map != 0, max_height);
    CudaSpace::cuda_rayTrace<<<grid, block>>>(d_rgb);
  }
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  if (ms_out) *ms_out = ms / n_repeat;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  CudaSpace::freeDeviceVariables();
  return (int)err;
}
