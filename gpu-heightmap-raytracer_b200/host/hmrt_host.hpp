/*
 * hmrt_host.hpp -- C++ host side above the C ABI (include/hmrt.h), mirroring the reference's
 * host-side flow in GPUHeightmapRaytracer/src/main.cpp:
 *
 *   reference (main.cpp)                              here
 *   ---------------------------------------------------------------------------------------------
 *   initialize(): LOD tables            :995-1003     PyramidLayout
 *   readLASHeader()                     :124-168      LasFile::open + SceneInfo (cell size 2.0,
 *                                                     boundaries, camera on the first point, max_height)
 *   allocateSection + loadLASToSection  :256-269,     Heightmap::clear + Heightmap::rasterise*()
 *                                       :174-244      (GPU scatter + max-mipmap build; points are
 *                                                     streamed through pinned host memory)
 *   PointdataGenerator output           PDG:186-205   read_pdg_text / pdg_generate (seeded)
 *   moveCamera / rotateCamera           :753-781      Camera::move / Camera::rotate
 *   updateTexture -> CudaSpace::rayTrace :675-690     Renderer::render (device or host framebuffer)
 *   GL PBO + quad                       :635-743      write_ppm (headless; GL is out of scope)
 *   initializeSections / manageSections :276-448      SectionGrid::initialize / manage (the sections are resident
 *                                                     DEVICE heightmaps, filled by the GPU rasteriser -- no loader
 *                                                     threads, no host copies)
 *   preparePointBuffer + copyPointBuffer :459-625     SectionGrid::prepare_window (hmrt_window_place +
 *                                                     hmrt_compose_window: a device-side gather, no PCIe upload)
 */
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "../../include/hmrt.h"

namespace hmrt_host {

struct Vec3 {
  float x, y, z;
};

/* main.cpp:995-1003 */
struct PyramidLayout {
  int coarse_res = 0, levels = 0;
  std::vector<int> res;
  std::vector<int64_t> idx;
  int64_t total = 0;
  PyramidLayout() = default;
  PyramidLayout(int coarse_res, int levels);
  int finest() const { return res.empty() ? 0 : res[0]; }
};

/* LAS 1.2 public header block + raw point records (formats 0-3), the subset of libLAS the reference uses */
struct LasFile {
  int point_format = 0, record_len = 0;
  uint64_t n_points = 0;
  uint32_t offset_to_points = 0;
  double scale[3] = {1, 1, 1}, offset[3] = {0, 0, 0}, min[3] = {0, 0, 0}, max[3] = {0, 0, 0};
  std::string path;
  static bool open(const std::string& path, LasFile& out, std::string* error = nullptr);
  /* read records [first, first + count) into dst (count * record_len bytes) */
  bool read_records(uint64_t first, uint64_t count, uint8_t* dst) const;
  /* first point's world coordinates (the reference places the camera there, main.cpp:159-161) */
  bool first_point(double xyz[3]) const;
  hmrt_las_transform transform(const float cell_size[3], const float origin[2]) const;
};

/* what readLASHeader derives (main.cpp:153-164) */
struct SceneInfo {
  float cell_size[3] = {2.f, 2.f, 2.f}; /* main.cpp:154-155 */
  float boundaries[2] = {0, 0};         /* main.cpp:156 */
  Vec3 camera_position{0, 0, 0};        /* main.cpp:161 */
  float max_height = 0;                 /* main.cpp:164 */
};
SceneInfo read_las_header(const LasFile& las);

/* PointdataGenerator (PointdataGenerator/main.cpp:72-184), seeded: (n+1)^2 xyz triples */
std::vector<float> pdg_generate(int n, uint64_t seed);
bool write_pdg_text(const std::string& path, const std::vector<float>& xyz); /* PDG:186-205, 7 decimals */
bool read_pdg_text(const std::string& path, std::vector<float>& xyz);

/* A resident heightmap: pyramid (+ optional colour map) in device memory, owned by this object
 * (the reference's d_point_buffer / d_color_map, main.cpp:1012-1013). */
class Heightmap {
 public:
  Heightmap(hmrt_ctx* ctx, int coarse_res, int levels, bool with_colors);
  ~Heightmap();
  Heightmap(const Heightmap&) = delete;
  Heightmap& operator=(const Heightmap&) = delete;
  bool ok() const { return status_ == 0; }
  int status() const { return status_; }
  const PyramidLayout& layout() const { return layout_; }
  float* d_pyramid() const { return d_pyramid_; }
  hmrt_color* d_color_map() const { return d_color_map_; }

  int clear(); /* allocateSection's zero-init, main.cpp:259-260 */
  /* loadLASToSection for one section origin: stream the file's records to the GPU in chunks */
  int rasterise_las(const LasFile& las, const float cell_size[3], const float origin[2], uint64_t chunk_points = 1u << 22);
  int rasterise_xyz(const std::vector<float>& xyz, const float cell_size[3], const float origin[2]);
  int finish(); /* build the mip levels, resolve colours */
  int max_height(float* out) const; /* maximum of the coarsest level (device -> host) */

 private:
  hmrt_ctx* ctx_;
  PyramidLayout layout_;
  float* d_pyramid_ = nullptr;
  hmrt_color* d_color_map_ = nullptr;
  uint64_t* d_color_keys_ = nullptr;
  uint64_t points_seen_ = 0;
  int status_ = 0;
};

/* The origins of the section grid and the reference's shift rules (manageSections, main.cpp:407-448; rearrangeSections*,
 * :329-402) as pure host logic: `tag[i][j]` identifies which section object sits at (i, j). */
struct SectionLayout {
  int grid = 0, coarse_res = 0, levels = 0;
  std::vector<float> origin; /* [(i * grid + j) * 2 + {x, y}] == point_sections_origins[i][j] */
  std::vector<int> tag;      /* [i * grid + j] */
  struct Load {
    int i, j, tag;   /* the section object `tag`, now at (i, j), must be (re)filled ... */
    float origin[2]; /* ... for this origin (allocateSection, main.cpp:256-269) */
  };
  float section_size() const; /* pow(2, levels - 1) * coarse_res: world extent of a section in cells */
  /* initializeSections (main.cpp:276-288): grid x grid sections centred on the camera; every section is a Load */
  std::vector<Load> initialize(int grid, int coarse_res, int levels, const Vec3& camera_position);
  /* manageSections (main.cpp:407-448): shift the grid when the camera leaves the inner sections; returns the sections to fill */
  std::vector<Load> manage(const Vec3& camera_position);
};

/* The section grid resident in device memory + the per-frame camera window. */
class SectionGrid {
 public:
  /* fills one section for an origin: the reference's loadLASToSection(file, origin, ...) (main.cpp:174) */
  using Loader = std::function<int(Heightmap& section, const float origin[2])>;
  SectionGrid(hmrt_ctx* ctx, int coarse_res, int levels, int grid, bool with_colors, Loader loader);
  bool ok() const { return status_ == 0; }
  int status() const { return status_; }
  int initialize(const Vec3& camera_position); /* initializeSections */
  /* manageSections: returns the number of sections (re)loaded, or a negative / CUDA error code */
  int manage(const Vec3& camera_position);
  /* preparePointBuffer + copyPointBuffer (main.cpp:459-625): compose the camera-centred window into `window`
   * (a Heightmap of the same layout) and return camera_point_buffer, the camera in window coordinates */
  int prepare_window(const Vec3& camera_position, Heightmap& window, Vec3* camera_point_buffer);
  const SectionLayout& layout() const { return layout_; }
  const Heightmap& section(int i, int j) const { return *sections_[layout_.tag[(size_t)i * layout_.grid + j]]; }

 private:
  int fill(const std::vector<SectionLayout::Load>& loads);
  hmrt_ctx* ctx_;
  int coarse_res_, levels_, grid_;
  Loader loader_;
  SectionLayout layout_;
  std::vector<std::unique_ptr<Heightmap>> sections_; /* indexed by tag */
  int status_ = 0;
};

/* camera state + the reference's motion rules */
struct Camera {
  Vec3 position{0, 0, 0};
  Vec3 forward{0.f, -0.6689647f, 0.7432941f}; /* normalize(0, -.9, 1), main.cpp:56 */
  Vec3 frame_dimension{32.f, 18.f, 20.f};     /* main.cpp:57 */
  /* moveCamera (main.cpp:756-772): dt-scaled translation, clamped to the boundaries / 4*max_height */
  void move(float fwd, float right, float up, float dt, const float boundaries[2], float max_height);
  /* rotateCamera (main.cpp:779-780): yaw about +y, pitch about the right vector (radians * dt) */
  void rotate(float yaw, float pitch, float dt);
  hmrt_camera abi() const;
};

/* rayTrace + framebuffer ownership (the mapped PBO of main.cpp:675-690) */
class Renderer {
 public:
  Renderer(hmrt_ctx* ctx, int width, int height);
  ~Renderer();
  int set_heightmap(const Heightmap& hm, float max_height);
  /* one frame into the device framebuffer (synchronous like the reference) */
  int render(const Camera& cam, const hmrt_trace_opts& opts);
  /* n frames straight into host memory through hmrt_trace_host */
  int render_to_host(const std::vector<Camera>& cams, const hmrt_trace_opts& opts, uint8_t* h_rgb);
  int download(std::vector<uint8_t>& rgb) const;
  int width() const { return w_; }
  int height() const { return h_; }
  uint8_t* d_rgb() const { return d_rgb_; }

 private:
  hmrt_ctx* ctx_;
  int w_, h_;
  uint8_t* d_rgb_ = nullptr;
};

/* binary PPM; flip = true writes the last row first (the reference's row 0 is the bottom GL row) */
bool write_ppm(const std::string& path, const uint8_t* rgb, int width, int height, bool flip = true);

}  // namespace hmrt_host
