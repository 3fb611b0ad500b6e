/*
 * TEST INFRASTRUCTURE ONLY -- host harness around the reference's OWN host-side hot-path code.
 *
 * build_ref.sh cuts these line ranges VERBATIM out of /root/reference/GPUHeightmapRaytracer/src/main.cpp
 * (sed -n, from where the file lies; nothing is copied into the repository) and this file #includes them:
 *
 *   REFHOST_CUT_MAIN   main.cpp:44-618   globals (:47-112), readLASHeader (:124-168), loadLASToSection (:174-244),
 *                                        allocateSection (:256-269), initializeSections (:276-289),
 *                                        unloadSectionsColumn/Row (:294-323), rearrangeSectionsX/Y (:329-402),
 *                                        manageSections (:408-448), preparePointBuffer (:459-618)
 *                      main.cpp:745-781  moveCamera (:753-772), rotateCamera (:776-781)
 *   REFHOST_CUT_TABLES main.cpp:995-1003 the pyramid tables inside initialize()
 *
 * What is NOT the reference's and is stubbed below, because it is third-party / OS code that is absent here:
 *   - liblas::{ReaderFactory, Reader, Header, Point, Classification, Color}: libLAS 1.8.0 ships as headers only in the
 *     reference tree (no .cpp / .lib).  The stub serves LAS 1.2 records (formats 0-3) from memory and decodes
 *     Point::GetX/Y/Z as `raw * scale + offset` in double, Classification::GetClass as the low 5 bits and the colour as
 *     three uint16 -- libLAS's published behaviour (inc/liblas/point.hpp:113-117, classification.hpp:83-88,159,
 *     color.hpp:59,118-140).  THAT decode step stays "parity unpinned"; everything after it is the reference's text.
 *   - std::ifstream (the file under ../Data): replaced by an always-open memory stream through a macro.
 *   - std::thread + Win32 SetThreadPriority: allocateSection's `new std::thread(loadLASToSection, ...)` is recorded, not
 *     run; the harness runs the verbatim loadLASToSection itself (synchronously) when a test asks for a section's
 *     content.  The loader's trailing `while (!*exit_control) yield; delete[] ...` (main.cpp:240-243) is left through an
 *     exception thrown from the stub stream's close() (main.cpp:237), so the buffers stay readable.
 *   - GLuint, cudaGraphicsResource: typedef / forward declaration for two unused globals.
 *   - `float glm::pow(float, int)`: the MSVC <cmath> overload the text relies on (see below).
 * LOD_levels (8) and point_sections_size (4) are compile-time constants of the reference (main.cpp:74,83); the
 * `_g3l4` variant of this library is built from a temp copy in which ONLY those two literals are changed by sed.
 */
#include <chrono>
#include <cfloat>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include <glm/glm.hpp>
#include <glm/gtc/constants.hpp>
#include <glm/gtx/rotate_vector.hpp>

#include "CudaKernel.cuh"  // the reference header (cast-patched temp copy, see build_ref.sh)

#include "hmrt_oracle.h"

/* MSVC 2015's <cmath> has the overload `float pow(float, int)` (reached through glm's `using std::pow`); g++ promotes the
 * same call to double, and then `glm::vec2 * glm::pow(2.0f, LOD_levels - 1)` (main.cpp:286-287,419,...) does not even
 * compile.  Same original-toolchain meaning as oracle/shim/device_launch_parameters.h; all uses are exact powers of two. */
namespace glm {
inline float pow(float a, int b) { return ::powf(a, static_cast<float>(b)); }
}
typedef unsigned int GLuint;
struct cudaGraphicsResource;

thread_local uint3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0};
thread_local dim3 blockDim(1, 1, 1), gridDim(1, 1, 1);

/* ------------------------------------------------------------------ the in-memory "LAS file" */
namespace refhost {

struct LasFile {
  const uint8_t* records = nullptr;
  int64_t n = 0;
  int record_len = 0, point_format = 0;
  double scale[3] = {1, 1, 1}, offset[3] = {0, 0, 0}, mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
};
LasFile g_file;
bool g_throw_on_close = false;  // set around synchronous loader runs
struct LoaderFinished {};

struct MemStream {
  void open(const std::string&, std::ios::openmode) {}
  bool is_open() const { return true; }
  void close() {
    if (g_throw_on_close) throw LoaderFinished();
  }
};

struct ThreadRecord {  // what allocateSection handed to `new std::thread(loadLASToSection, ...)`
  glm::vec2 origin;
  bool* exit_control;
  float* point_section;
  CudaSpace::Color* color_section;
};
std::vector<ThreadRecord> g_spawn_log;

struct LoaderThread {
  ThreadRecord rec;
  bool attached = true;
  template <class F>
  LoaderThread(F, std::string, glm::vec2 origin, bool* exit_control, float* point_section, CudaSpace::Color* color_section)
      : rec{origin, exit_control, point_section, color_section} {
    g_spawn_log.push_back(rec);
  }
  /* the loader thread frees its section when it is told to exit (main.cpp:240-243); exit flags are leaked on purpose:
   * the reference never shifts thread_exit[][] in rearrangeSections*, so later unloads write through stale flags */
  ~LoaderThread() {
    delete[] rec.point_section;
    delete[] rec.color_section;
  }
  bool joinable() const { return attached; }
  void detach() { attached = false; }
  int native_handle() const { return 0; }
};

}  // namespace refhost

namespace liblas {
struct Classification {
  uint8_t flags;
  uint8_t GetClass() const { return flags & 0x1f; }
};
struct Color {
  uint16_t c[3];
  uint16_t GetRed() const { return c[0]; }
  uint16_t GetGreen() const { return c[1]; }
  uint16_t GetBlue() const { return c[2]; }
};
struct Header {
  bool Compressed() const { return false; }
  uint32_t GetPointRecordsCount() const { return (uint32_t)refhost::g_file.n; }
  double GetMinX() const { return refhost::g_file.mn[0]; }
  double GetMinY() const { return refhost::g_file.mn[1]; }
  double GetMinZ() const { return refhost::g_file.mn[2]; }
  double GetMaxX() const { return refhost::g_file.mx[0]; }
  double GetMaxY() const { return refhost::g_file.mx[1]; }
  double GetMaxZ() const { return refhost::g_file.mx[2]; }
  double GetScaleX() const { return refhost::g_file.scale[0]; }
  double GetScaleY() const { return refhost::g_file.scale[1]; }
  double GetScaleZ() const { return refhost::g_file.scale[2]; }
  double GetOffsetX() const { return refhost::g_file.offset[0]; }
  double GetOffsetY() const { return refhost::g_file.offset[1]; }
  double GetOffsetZ() const { return refhost::g_file.offset[2]; }
};
struct Point {
  int32_t raw[3];
  Classification cls;
  Color color;
  double GetX() const { return raw[0] * refhost::g_file.scale[0] + refhost::g_file.offset[0]; }
  double GetY() const { return raw[1] * refhost::g_file.scale[1] + refhost::g_file.offset[1]; }
  double GetZ() const { return raw[2] * refhost::g_file.scale[2] + refhost::g_file.offset[2]; }
  Classification const& GetClassification() const { return cls; }
  Color const& GetColor() const { return color; }
};
struct Reader {
  int64_t next = 0;
  Point cur;
  Header hdr;
  Header const& GetHeader() const { return hdr; }
  bool ReadNextPoint() {
    const refhost::LasFile& f = refhost::g_file;
    if (next >= f.n) return false;
    const uint8_t* p = f.records + next * f.record_len;
    std::memcpy(cur.raw, p, 12);  // X, Y, Z: little-endian int32 at 0, 4, 8 (LAS 1.2 formats 0-3)
    cur.cls.flags = p[15];
    static const int rgb_off[4] = {-1, -1, 20, 28};
    if (rgb_off[f.point_format] >= 0)
      std::memcpy(cur.color.c, p + rgb_off[f.point_format], 6);
    else
      cur.color = Color{{0, 0, 0}};
    next++;
    return true;
  }
  Point const& GetPoint() const { return cur; }
};
struct ReaderFactory {
  Reader CreateWithStream(refhost::MemStream&) { return Reader(); }
};
}  // namespace liblas

namespace std {
using hmrt_refhost_memstream = ::refhost::MemStream;
using hmrt_refhost_thread = ::refhost::LoaderThread;
}  // namespace std
static inline void SetThreadPriority(int, int) {}

#define ifstream hmrt_refhost_memstream
#define thread hmrt_refhost_thread
#include REFHOST_CUT_MAIN
#undef ifstream
#undef thread

static void refhost_tables() {
#include REFHOST_CUT_TABLES
}

/* ------------------------------------------------------------------------------ C interface */
namespace {
std::mutex g_lock;  // the reference keeps its state in namespace-scope globals
bool g_sections_live = false;

void drop_sections() {
  if (!g_sections_live) return;
  for (int i = 0; i < point_sections_size; i++)
    for (int j = 0; j < point_sections_size; j++) {
      delete thread_pool[i][j];  // frees the section buffers (see LoaderThread)
      thread_pool[i][j] = nullptr;
      point_sections[i][j] = nullptr;
      color_sections[i][j] = nullptr;
    }
  g_sections_live = false;
}

void run_loader(glm::vec2 origin, float* point_section, CudaSpace::Color* color_section) {
  bool* never = new bool(false);
  refhost::g_throw_on_close = true;
  try {
    loadLASToSection(point_cloud_file, origin, never, point_section, color_section);
  } catch (refhost::LoaderFinished&) {
  }
  refhost::g_throw_on_close = false;
  delete never;
}

struct QuietCout {
  std::streambuf* old;
  std::ostringstream sink;
  QuietCout() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~QuietCout() { std::cout.rdbuf(old); }
};
}  // namespace

extern "C" {

int hmrt_refhost_levels(void) { return LOD_levels; }
int hmrt_refhost_grid(void) { return point_sections_size; }

/* point_buffer_resolution = (coarse, coarse); then main.cpp:995-1003.  Drops any live sections and buffers. */
int hmrt_refhost_config(int coarse_res, int* res, int64_t* idx, int64_t* total) {
  std::lock_guard<std::mutex> lk(g_lock);
  if (coarse_res < 1) return -1;
  drop_sections();
  delete[] h_point_buffer;
  delete[] h_color_map;
  h_point_buffer = nullptr, h_color_map = nullptr;
  point_buffer_resolution = glm::ivec2(coarse_res, coarse_res);
  refhost_tables();
  for (int i = 0; i < LOD_levels; i++) {
    if (res) res[i] = LOD_resolutions[i];
    if (idx) idx[i] = LOD_indexes[i];
  }
  if (total) *total = (int64_t)stride_x * coarse_res * coarse_res;
  return 0;
}

/* The records stay borrowed until the next call. */
int hmrt_refhost_set_las(const uint8_t* records, int64_t n, int record_len, int point_format, const double scale[3],
                         const double offset[3], const double mn[3], const double mx[3]) {
  std::lock_guard<std::mutex> lk(g_lock);
  if (point_format < 0 || point_format > 3 || (n > 0 && !records)) return -1;
  refhost::LasFile& f = refhost::g_file;
  f.records = records, f.n = n, f.record_len = record_len, f.point_format = point_format;
  for (int a = 0; a < 3; a++) f.scale[a] = scale[a], f.offset[a] = offset[a], f.mn[a] = mn[a], f.mx[a] = mx[a];
  return 0;
}

void hmrt_refhost_set_cell_size(float x, float y, float z) {
  std::lock_guard<std::mutex> lk(g_lock);
  cell_size = glm::vec3(x, y, z);
}

/* readLASHeader (main.cpp:124-168): sets cell_size = 2, boundaries, camera_position, max_height. */
int hmrt_refhost_read_header(float camera[3], float bounds[2], float* max_h, float cell[3]) {
  std::lock_guard<std::mutex> lk(g_lock);
  if (refhost::g_file.n < 1) return -1;
  {
    QuietCout quiet;
    readLASHeader(point_cloud_file);
  }
  camera[0] = camera_position.x, camera[1] = camera_position.y, camera[2] = camera_position.z;
  bounds[0] = boundaries.x, bounds[1] = boundaries.y;
  *max_h = max_height;
  cell[0] = cell_size.x, cell[1] = cell_size.y, cell[2] = cell_size.z;
  return 0;
}

/* One section the way allocateSection (main.cpp:259-260) makes it, filled by loadLASToSection (:174-238). */
int hmrt_refhost_rasterise(const float origin[2], float* pyramid_out, hmrt_color* color_out) {
  std::lock_guard<std::mutex> lk(g_lock);
  const size_t nf = (size_t)stride_x * point_buffer_resolution.x * point_buffer_resolution.y;
  const size_t nc = (size_t)LOD_resolutions[0] * LOD_resolutions[0];
  float* sec = new float[nf]();
  CudaSpace::Color* col = new CudaSpace::Color[nc];
  run_loader(glm::vec2(origin[0], origin[1]), sec, col);
  std::memcpy(pyramid_out, sec, nf * sizeof(float));
  if (color_out) std::memcpy(color_out, col, nc * sizeof(CudaSpace::Color));
  delete[] sec;
  delete[] col;
  return 0;
}

static void report_sections(float* origins, int* loaded_slots, float* loaded_origins, int* n_loaded) {
  const int g = point_sections_size;
  for (int i = 0; i < g; i++)
    for (int j = 0; j < g; j++) {
      origins[(i * g + j) * 2 + 0] = point_sections_origins[i][j].x;
      origins[(i * g + j) * 2 + 1] = point_sections_origins[i][j].y;
    }
  /* which slots were (re)allocated by this call, in call order: match the spawn log against the slot buffers */
  int n = 0;
  for (const refhost::ThreadRecord& r : refhost::g_spawn_log)
    for (int i = 0; i < g; i++)
      for (int j = 0; j < g; j++)
        if (point_sections[i][j] == r.point_section) {
          loaded_slots[2 * n] = i, loaded_slots[2 * n + 1] = j;
          loaded_origins[2 * n] = r.origin.x, loaded_origins[2 * n + 1] = r.origin.y;
          n++;
        }
  *n_loaded = n;
  refhost::g_spawn_log.clear();
}

/* camera_position = camera; initializeSections() (main.cpp:276-289).  origins: [grid][grid][2]; loaded_*: up to grid*grid. */
int hmrt_refhost_init_sections(const float camera[3], float* origins, int* loaded_slots, float* loaded_origins, int* n_loaded) {
  std::lock_guard<std::mutex> lk(g_lock);
  drop_sections();
  refhost::g_spawn_log.clear();
  camera_position = glm::vec3(camera[0], camera[1], camera[2]);
  initializeSections();
  g_sections_live = true;
  report_sections(origins, loaded_slots, loaded_origins, n_loaded);
  return 0;
}

/* camera_position = camera; manageSections() (main.cpp:408-448). */
int hmrt_refhost_manage(const float camera[3], float* origins, int* loaded_slots, float* loaded_origins, int* n_loaded) {
  std::lock_guard<std::mutex> lk(g_lock);
  if (!g_sections_live) return -1;
  camera_position = glm::vec3(camera[0], camera[1], camera[2]);
  manageSections();
  report_sections(origins, loaded_slots, loaded_origins, n_loaded);
  return 0;
}

/* Fill section (i, j): from the LAS file through the verbatim loader (pyramid == NULL), or with the given content. */
int hmrt_refhost_fill_section(int i, int j, const float* pyramid, const hmrt_color* colors) {
  std::lock_guard<std::mutex> lk(g_lock);
  const int g = point_sections_size;
  if (!g_sections_live || i < 0 || j < 0 || i >= g || j >= g) return -1;
  const size_t nf = (size_t)stride_x * point_buffer_resolution.x * point_buffer_resolution.y;
  const size_t nc = (size_t)LOD_resolutions[0] * LOD_resolutions[0];
  if (!pyramid) {
    run_loader(point_sections_origins[i][j], point_sections[i][j], color_sections[i][j]);
    return 0;
  }
  std::memcpy(point_sections[i][j], pyramid, nf * sizeof(float));
  if (colors) std::memcpy(color_sections[i][j], colors, nc * sizeof(CudaSpace::Color));
  return 0;
}

int hmrt_refhost_read_section(int i, int j, float* pyramid, hmrt_color* colors) {
  std::lock_guard<std::mutex> lk(g_lock);
  const int g = point_sections_size;
  if (!g_sections_live || i < 0 || j < 0 || i >= g || j >= g) return -1;
  const size_t nf = (size_t)stride_x * point_buffer_resolution.x * point_buffer_resolution.y;
  const size_t nc = (size_t)LOD_resolutions[0] * LOD_resolutions[0];
  if (pyramid) std::memcpy(pyramid, point_sections[i][j], nf * sizeof(float));
  if (colors) std::memcpy(colors, color_sections[i][j], nc * sizeof(CudaSpace::Color));
  return 0;
}

/* camera_position = camera; preparePointBuffer() (main.cpp:459-618); returns camera_point_buffer and both buffers. */
int hmrt_refhost_prepare(const float camera[3], float camera_point_buffer_out[3], float* point_buffer_out, hmrt_color* color_map_out) {
  std::lock_guard<std::mutex> lk(g_lock);
  if (!g_sections_live) return -1;
  const size_t nf = (size_t)stride_x * point_buffer_resolution.x * point_buffer_resolution.y;
  const size_t nc = (size_t)LOD_resolutions[0] * LOD_resolutions[0];
  if (!h_point_buffer) h_point_buffer = new float[nf];
  if (!h_color_map) h_color_map = new CudaSpace::Color[nc];
  camera_position = glm::vec3(camera[0], camera[1], camera[2]);
  preparePointBuffer();
  camera_point_buffer_out[0] = camera_point_buffer.x, camera_point_buffer_out[1] = camera_point_buffer.y,
  camera_point_buffer_out[2] = camera_point_buffer.z;
  if (point_buffer_out) std::memcpy(point_buffer_out, h_point_buffer, nf * sizeof(float));
  if (color_map_out) std::memcpy(color_map_out, h_color_map, nc * sizeof(CudaSpace::Color));
  return 0;
}

/* moveCamera (main.cpp:753-772): state in, state out.  move = (movement_rht, movement_up, movement_fwd). */
int hmrt_refhost_move_camera(float position[3], const float forward[3], const float move[3], float dt, const float bounds[2], float max_h) {
  std::lock_guard<std::mutex> lk(g_lock);
  camera_position = glm::vec3(position[0], position[1], position[2]);
  camera_forward = glm::vec3(forward[0], forward[1], forward[2]);
  movement_rht = move[0], movement_up = move[1], movement_fwd = move[2];
  delta_time = std::chrono::duration<float>(dt);
  boundaries = glm::vec2(bounds[0], bounds[1]);
  max_height = max_h;
  moveCamera();
  position[0] = camera_position.x, position[1] = camera_position.y, position[2] = camera_position.z;
  return 0;
}

/* rotateCamera (main.cpp:776-781): forward in/out. */
int hmrt_refhost_rotate_camera(float forward[3], float rot_up, float rot_right, float dt) {
  std::lock_guard<std::mutex> lk(g_lock);
  camera_forward = glm::vec3(forward[0], forward[1], forward[2]);
  rotation_up = rot_up, rotation_right = rot_right;
  delta_time = std::chrono::duration<float>(dt);
  rotateCamera();
  forward[0] = camera_forward.x, forward[1] = camera_forward.y, forward[2] = camera_forward.z;
  return 0;
}

}  // extern "C"
