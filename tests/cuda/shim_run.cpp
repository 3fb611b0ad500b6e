/*
 * TEST INFRASTRUCTURE ONLY: the reference's own call ORDER around its three CudaSpace entry points, executed on a GPU against
 * the drop-in header gpu-heightmap-raytracer_b200/host/CudaKernel.cuh, with the reference's own GLM types:
 *     cudaMalloc x2 (main.cpp:1012-1013) -> cudaMemcpy x2 (:623-624) -> CudaSpace::initializeDeviceVariables (:1014)
 *     -> per frame: CudaSpace::rayTrace on a device colour buffer (:686; the mapped PBO of :681-683 is a plain device
 *        buffer here) -> cudaFree x2 (:1022-1023) -> CudaSpace::freeDeviceVariables (:1024)
 * Usage: shim_run <scene.bin> <frames.bin>
 *   scene.bin : int32 coarse_res, levels, W, H, n_frames, use_color; float max_height; then per frame 9 floats
 *               (frame_dimension, camera_forward, grid_camera_position); then the pyramid floats; then the colour map bytes
 *   frames.bin: n_frames * W * H * 3 bytes (what the reference hands to glTexSubImage2D, main.cpp:701-703)
 * Built where /root/reference is present (GLM comes from its include tree; nothing is copied); the binary travels to the GPU box.
 */
#include <cstdint>
#include <cstdio>
#include <vector>

#include <cuda_runtime.h>
#include <glm/glm.hpp>

#include "CudaKernel.cuh"

/* globals with the reference's names and types (main.cpp:52-58, 71, 83-85, 110-111) */
glm::ivec2 texture_resolution(1920, 1080);
glm::vec3 camera_point_buffer(0, 0, 0), camera_forward(0, 0, 1), frame_dimension(32, 18, 20);
glm::ivec2 point_buffer_resolution(32, 32);
int LOD_levels = 8;
int stride_x = 0;
float max_height = 0;
bool use_color_map = false;
float* d_point_buffer = nullptr;
CudaSpace::Color* d_color_map = nullptr;

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  int32_t hdr[6];
  if (std::fread(hdr, sizeof(hdr), 1, f) != 1 || std::fread(&max_height, 4, 1, f) != 1) return 2;
  point_buffer_resolution = glm::ivec2(hdr[0], hdr[0]);
  LOD_levels = hdr[1];
  texture_resolution = glm::ivec2(hdr[2], hdr[3]);
  const int n_frames = hdr[4];
  use_color_map = hdr[5] != 0;
  std::vector<float> cams((size_t)n_frames * 9);
  if (std::fread(cams.data(), 4, cams.size(), f) != cams.size()) return 2;
  stride_x = 0; /* main.cpp:997-1002 */
  for (int i = 0; i < LOD_levels; ++i) stride_x += 1 << (2 * i);
  const size_t n_floats = (size_t)point_buffer_resolution.x * point_buffer_resolution.y * stride_x;
  const size_t res0 = (size_t)point_buffer_resolution.x << (LOD_levels - 1);
  std::vector<float> h_point_buffer(n_floats);
  std::vector<CudaSpace::Color> h_color_map(res0 * res0);
  if (std::fread(h_point_buffer.data(), 4, n_floats, f) != n_floats) return 2;
  if (std::fread(h_color_map.data(), 3, res0 * res0, f) != res0 * res0) return 2;
  std::fclose(f);

  if (cudaMalloc(&d_point_buffer, sizeof(float) * n_floats) != cudaSuccess) return 3;                          /* main.cpp:1012 */
  if (cudaMalloc(&d_color_map, sizeof(CudaSpace::Color) * res0 * res0) != cudaSuccess) return 3;               /* main.cpp:1013 */
  CudaSpace::initializeDeviceVariables(point_buffer_resolution, texture_resolution, d_point_buffer, d_color_map, LOD_levels, stride_x, max_height); /* :1014 */
  const size_t frame_bytes = (size_t)texture_resolution.x * texture_resolution.y * 3;
  unsigned char* devPtr = nullptr; /* the mapped PBO */
  if (cudaMalloc(&devPtr, frame_bytes) != cudaSuccess) return 3;
  std::vector<unsigned char> out(frame_bytes);
  FILE* o = std::fopen(argv[2], "wb");
  if (!o) return 2;
  for (int i = 0; i < n_frames; ++i) {
    /* copyPointBuffer, main.cpp:623-624: the reference refills both buffers every frame; the library must not cache their contents */
    if (cudaMemcpy(d_point_buffer, h_point_buffer.data(), sizeof(float) * n_floats, cudaMemcpyHostToDevice) != cudaSuccess) return 3;
    if (cudaMemcpy(d_color_map, h_color_map.data(), sizeof(CudaSpace::Color) * res0 * res0, cudaMemcpyHostToDevice) != cudaSuccess) return 3;
    const float* c = &cams[(size_t)i * 9];
    frame_dimension = glm::vec3(c[0], c[1], c[2]);
    camera_forward = glm::vec3(c[3], c[4], c[5]);
    camera_point_buffer = glm::vec3(c[6], c[7], c[8]);
    CudaSpace::rayTrace(texture_resolution, frame_dimension, camera_forward, camera_point_buffer, devPtr, use_color_map, max_height); /* main.cpp:686 */
    /* rayTrace is synchronous (CudaKernel.cu:307): the frame is complete here */
    if (cudaMemcpy(out.data(), devPtr, frame_bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return 3;
    std::fwrite(out.data(), 1, frame_bytes, o);
    /* between frames the reference's heightmap changes under the library's feet: scale it once to prove nothing is cached */
    if (i == 0 && n_frames > 1)
      for (float& v : h_point_buffer) v *= 0.5f;
  }
  std::fclose(o);
  cudaFree(devPtr);
  cudaFree(d_point_buffer);          /* main.cpp:1022 */
  cudaFree(d_color_map);             /* main.cpp:1023 */
  CudaSpace::freeDeviceVariables();  /* main.cpp:1024 */
  CudaSpace::freeDeviceVariables();  /* idempotent */
  std::puts("shim_run: ok");
  return 0;
}
