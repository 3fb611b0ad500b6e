/*
 * TEST INFRASTRUCTURE ONLY: compiles the reference's three CudaSpace call sites
 * (GPUHeightmapRaytracer/src/main.cpp:1014, :686, :1024), with the reference's own GLM types, against
 * the drop-in header gpu-heightmap-raytracer_b200/host/CudaKernel.cuh.  Built only where
 * /root/reference is present (GLM comes from its include tree; nothing is copied).
 */
#include <glm/glm.hpp>

#include "CudaKernel.cuh"

/* globals with the reference's names and types (main.cpp:52-58, 71, 83-85, 110-111) */
glm::ivec2 texture_resolution(1920, 1080);
glm::vec3 camera_point_buffer(0, 0, 0), camera_forward(glm::normalize(glm::vec3(0, -.9, 1))), frame_dimension(16 * 2, 9 * 2, 20);
glm::ivec2 point_buffer_resolution(32, 32);
const int LOD_levels = 8;
int stride_x = 21845;
float max_height = 0;
bool use_color_map = false;
float* d_point_buffer = nullptr;
CudaSpace::Color* d_color_map = nullptr;

int main() {
  unsigned char* devPtr = nullptr;
  CudaSpace::Color c(static_cast<unsigned short>(65535), static_cast<unsigned short>(32768), static_cast<unsigned short>(0));
  /* main.cpp:1014 */
  CudaSpace::initializeDeviceVariables(point_buffer_resolution, texture_resolution, d_point_buffer, d_color_map, LOD_levels, stride_x, max_height);
  /* main.cpp:686 */
  CudaSpace::rayTrace(texture_resolution, frame_dimension, camera_forward, camera_point_buffer, devPtr, use_color_map, max_height);
  /* main.cpp:1024 */
  CudaSpace::freeDeviceVariables();
  return c.r == 255 ? 0 : 1;
}
