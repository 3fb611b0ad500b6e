/* hmrt_host.cpp -- see hmrt_host.hpp.  Host logic only; every computation on the traced /
 * rasterised path happens in libhmrt.so's CUDA kernels. */
#include "hmrt_host.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>

namespace hmrt_host {

/* ---------------------------------------------------------------------------------------------- */
PyramidLayout::PyramidLayout(int coarse, int lv) : coarse_res(coarse), levels(lv) {
  if (lv < 1 || lv > HMRT_MAX_LEVELS) return;
  res.assign(lv, 0);
  idx.assign(lv, 0);
  if (hmrt_pyramid_layout(coarse, lv, res.data(), idx.data(), &total) != 0) {
    res.clear();
    idx.clear();
    total = 0;
  }
}

/* ---------------------------------------------------------------------------------------------- */
static uint16_t rd_u16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static uint32_t rd_u32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static double rd_f64(const uint8_t* p) {
  double d;
  std::memcpy(&d, p, 8);
  return d;
}

bool LasFile::open(const std::string& path, LasFile& out, std::string* error) {
  auto fail = [&](const char* why) {
    if (error) *error = why;
    return false;
  };
  std::ifstream f(path, std::ios::binary);
  if (!f) return fail("cannot open file");
  uint8_t h[227];
  if (!f.read(reinterpret_cast<char*>(h), sizeof h)) return fail("short LAS header");
  if (std::memcmp(h, "LASF", 4) != 0) return fail("not a LAS file");
  if (h[104] & 0xC0) return fail("compressed (LAZ) point data is not supported");
  out.point_format = h[104] & 0x3F;
  static const int min_len[4] = {20, 28, 26, 34};
  if (out.point_format > 3) return fail("unsupported point data format (0-3 supported)");
  out.record_len = rd_u16(h + 105);
  if (out.record_len < min_len[out.point_format]) return fail("record length shorter than the point format");
  out.offset_to_points = rd_u32(h + 96);
  out.n_points = rd_u32(h + 107);
  for (int i = 0; i < 3; ++i) {
    out.scale[i] = rd_f64(h + 131 + 8 * i);
    out.offset[i] = rd_f64(h + 155 + 8 * i);
    out.max[i] = rd_f64(h + 179 + 16 * i);
    out.min[i] = rd_f64(h + 187 + 16 * i);
  }
  out.path = path;
  return true;
}

bool LasFile::read_records(uint64_t first, uint64_t count, uint8_t* dst) const {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  f.seekg((std::streamoff)offset_to_points + (std::streamoff)(first * (uint64_t)record_len));
  return (bool)f.read(reinterpret_cast<char*>(dst), (std::streamsize)(count * (uint64_t)record_len));
}

bool LasFile::first_point(double xyz[3]) const {
  if (n_points == 0) return false;
  std::vector<uint8_t> rec(record_len);
  if (!read_records(0, 1, rec.data())) return false;
  for (int i = 0; i < 3; ++i) xyz[i] = (double)(int32_t)rd_u32(rec.data() + 4 * i) * scale[i] + offset[i]; /* libLAS GetX() */
  return true;
}

hmrt_las_transform LasFile::transform(const float cell_size[3], const float origin[2]) const {
  hmrt_las_transform t;
  for (int i = 0; i < 3; ++i) {
    t.scale[i] = scale[i];
    t.offset[i] = offset[i];
    t.min[i] = min[i];
    t.cell_size[i] = cell_size[i];
  }
  t.origin[0] = origin[0];
  t.origin[1] = origin[1];
  return t;
}

SceneInfo read_las_header(const LasFile& las) { /* main.cpp:153-164 */
  SceneInfo s;
  const double dX = las.max[0] - las.min[0], dY = las.max[1] - las.min[1];
  s.boundaries[0] = (float)(dX / s.cell_size[0]);
  s.boundaries[1] = (float)(dY / s.cell_size[1]);
  double p[3];
  if (las.first_point(p)) {
    s.camera_position.x = (float)((p[0] - las.min[0]) / s.cell_size[0]);
    s.camera_position.y = (float)((las.max[2] - las.min[2]) / s.cell_size[2]);
    s.camera_position.z = (float)((p[1] - las.min[1]) / s.cell_size[0]);
  }
  s.max_height = static_cast<float>(las.max[2] - las.min[2]) / s.cell_size[2];
  return s;
}

/* ---------------------------------------------------------------------------------------------- */
/* PointdataGenerator/main.cpp:72-184 with an explicit seed (64-bit LCG, top 24 bits -> [0,1)) */
namespace {
struct Lcg {
  uint64_t s;
  float next() {
    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
    return (float)(s >> 40) * (1.0f / 16777216.0f);
  }
};
}  // namespace

std::vector<float> pdg_generate(int n, uint64_t seed) {
  const int G = n + 1;
  if (n < 2 || (n & (n - 1))) return {};
  std::vector<float> z((size_t)G * G, 0.f);
  auto at = [&](int i, int j) -> float& { return z[(size_t)i * G + j]; };
  Lcg gen{seed};
  at(0, 0) = gen.next(); /* PDG:91-94 */
  at(0, G - 1) = gen.next();
  at(G - 1, 0) = gen.next();
  at(G - 1, G - 1) = gen.next();
  auto rough = [&](int count) { return (float)std::pow((double)gen.next(), (double)count); }; /* PDG:105 */
  auto diamond = [&](int x, int y, int s, float r) {                                          /* PDG:117-130 */
    at(x, y) = (at(x + s, y + s) + at(x - s, y + s) + at(x + s, y - s) + at(x - s, y - s)) / 4 + r;
  };
  auto square = [&](int x, int y, int s, float r) { /* PDG:132-158 */
    float left = 0, right = 0, top = 0, bottom = 0;
    int count = 0;
    if (x > 0) left = at(x - s, y), count++;
    if (x < G - 1) right = at(x + s, y), count++;
    if (y > 0) top = at(x, y - s), count++;
    if (y < G - 1) bottom = at(x, y + s), count++;
    at(x, y) = (left + right + top + bottom) / count + r;
  };
  int count = 1;
  for (int i = G; i > 1; i /= 2) { /* PDG:100-112 */
    for (int x = i / 2; x < G; x += i)
      for (int y = i / 2; y < G; y += i) {
        float r = rough(count);
        diamond(x, y, i / 2, r);
        r = rough(count);
        square(x, y - i / 2, i / 2, r);
        r = rough(count);
        square(x - i / 2, y, i / 2, r);
        r = rough(count);
        square(x + i / 2, y, i / 2, r);
        r = rough(count);
        square(x, y + i / 2, i / 2, r);
      }
    ++count;
  }
  for (int i = 0; i < G - 1; ++i) /* scaleData, PDG:175-184 */
    for (int j = 0; j < G - 1; ++j) at(i, j) *= 10.f;
  std::vector<float> xyz((size_t)G * G * 3);
  for (int i = 0; i < G; ++i)
    for (int j = 0; j < G; ++j) {
      const size_t k = ((size_t)i * G + j) * 3;
      xyz[k] = (float)i, xyz[k + 1] = (float)j, xyz[k + 2] = at(i, j);
    }
  return xyz;
}

bool write_pdg_text(const std::string& path, const std::vector<float>& xyz) {
  FILE* f = std::fopen(path.c_str(), "w");
  if (!f) return false;
  for (size_t i = 0; i + 2 < xyz.size(); i += 3) std::fprintf(f, "%.7f %.7f %.7f\n", xyz[i], xyz[i + 1], xyz[i + 2]);
  return std::fclose(f) == 0;
}

bool read_pdg_text(const std::string& path, std::vector<float>& xyz) {
  FILE* f = std::fopen(path.c_str(), "r");
  if (!f) return false;
  xyz.clear();
  float x, y, z;
  while (std::fscanf(f, "%f %f %f", &x, &y, &z) == 3) {
    xyz.push_back(x);
    xyz.push_back(y);
    xyz.push_back(z);
  }
  std::fclose(f);
  return !xyz.empty();
}

/* ---------------------------------------------------------------------------------------------- */
Heightmap::Heightmap(hmrt_ctx* ctx, int coarse_res, int levels, bool with_colors) : ctx_(ctx), layout_(coarse_res, levels) {
  if (layout_.total == 0) {
    status_ = HMRT_E_SHAPE;
    return;
  }
  const size_t cells = (size_t)layout_.finest() * layout_.finest();
  status_ = (int)cudaMalloc(&d_pyramid_, sizeof(float) * (size_t)layout_.total); /* main.cpp:1012 */
  if (status_ == 0 && with_colors) {
    status_ = (int)cudaMalloc(&d_color_map_, 3 * cells); /* main.cpp:1013 */
    if (status_ == 0) status_ = (int)cudaMalloc(&d_color_keys_, sizeof(uint64_t) * cells);
  }
  if (status_ == 0) status_ = clear();
}

Heightmap::~Heightmap() {
  cudaFree(d_pyramid_);
  cudaFree(d_color_map_);
  cudaFree(d_color_keys_);
}

int Heightmap::clear() {
  points_seen_ = 0;
  return hmrt_clear_section(ctx_, d_pyramid_, layout_.coarse_res, layout_.levels, d_color_keys_, d_color_map_);
}

int Heightmap::rasterise_las(const LasFile& las, const float cell_size[3], const float origin[2], uint64_t chunk_points) {
  const hmrt_las_transform xf = las.transform(cell_size, origin);
  const size_t chunk_bytes = (size_t)chunk_points * las.record_len;
  /* copies and events run on the CONTEXT's stream, like the scatter they feed (hmrt_set_stream may have changed it) */
  cudaStream_t stream = (cudaStream_t)hmrt_get_stream(ctx_);
  uint8_t* h_buf[2] = {nullptr, nullptr};
  uint8_t* d_buf[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  int rc = 0;
  for (int i = 0; i < 2 && rc == 0; ++i) { /* double-buffered pinned staging */
    rc = (int)cudaMallocHost(&h_buf[i], chunk_bytes);
    if (rc == 0) rc = (int)cudaMalloc(&d_buf[i], chunk_bytes);
    if (rc == 0) rc = (int)cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming);
  }
  uint64_t first = 0;
  for (int k = 0; rc == 0 && first < las.n_points; ++k) {
    const int b = k & 1;
    const uint64_t n = std::min<uint64_t>(chunk_points, las.n_points - first);
    if (k >= 2) rc = (int)cudaEventSynchronize(done[b]); /* the scatter that read buffer b has finished: both halves are free again */
    if (rc == 0 && !las.read_records(first, n, h_buf[b])) rc = HMRT_E_ARG;
    if (rc == 0) rc = (int)cudaMemcpyAsync(d_buf[b], h_buf[b], (size_t)n * las.record_len, cudaMemcpyHostToDevice, stream);
    /* first_index == 0 only for the first chunk of a freshly cleared heightmap: the library probes the file's spatial
     * order there (one stream synchronisation) and keeps the verdict for the following chunks */
    if (rc == 0)
      rc = hmrt_scatter_las(ctx_, d_buf[b], (int64_t)n, las.record_len, las.point_format, &xf, (int64_t)(points_seen_ + first),
                            d_pyramid_, layout_.coarse_res, layout_.levels, d_color_keys_);
    if (rc == 0) rc = (int)cudaEventRecord(done[b], stream);
    first += n;
  }
  const int sync_rc = hmrt_synchronize(ctx_); /* also on the error paths: nothing may still read the staging buffers */
  if (rc == 0) rc = sync_rc;
  points_seen_ += las.n_points;
  for (int i = 0; i < 2; ++i) {
    if (done[i]) cudaEventDestroy(done[i]);
    if (h_buf[i]) cudaFreeHost(h_buf[i]);
    if (d_buf[i]) cudaFree(d_buf[i]);
  }
  return rc;
}

int Heightmap::rasterise_xyz(const std::vector<float>& xyz, const float cell_size[3], const float origin[2]) {
  hmrt_las_transform xf;
  for (int i = 0; i < 3; ++i) {
    xf.scale[i] = 1.0;
    xf.offset[i] = 0.0;
    xf.min[i] = 0.0;
    xf.cell_size[i] = cell_size[i];
  }
  xf.origin[0] = origin[0];
  xf.origin[1] = origin[1];
  const int64_t n = (int64_t)(xyz.size() / 3);
  float* d = nullptr;
  int rc = (int)cudaMalloc(&d, sizeof(float) * xyz.size());
  if (rc == 0) rc = (int)cudaMemcpy(d, xyz.data(), sizeof(float) * xyz.size(), cudaMemcpyHostToDevice);
  if (rc == 0) rc = hmrt_scatter_xyz(ctx_, d, n, &xf, d_pyramid_, layout_.coarse_res, layout_.levels);
  if (rc == 0) rc = hmrt_synchronize(ctx_);
  cudaFree(d);
  return rc;
}

int Heightmap::finish() {
  int rc = hmrt_build_mips(ctx_, d_pyramid_, layout_.coarse_res, layout_.levels);
  if (rc == 0 && d_color_keys_)
    rc = hmrt_resolve_colors(ctx_, d_color_keys_, d_color_map_, (int64_t)layout_.finest() * layout_.finest());
  if (rc == 0) rc = hmrt_synchronize(ctx_);
  return rc;
}

int Heightmap::max_height(float* out) const {
  const size_t n = (size_t)layout_.coarse_res * layout_.coarse_res;
  std::vector<float> top(n);
  int rc = (int)cudaMemcpy(top.data(), d_pyramid_, n * sizeof(float), cudaMemcpyDeviceToHost);
  if (rc == 0) *out = *std::max_element(top.begin(), top.end());
  return rc;
}

/* ---------------------------------------------------------------------------------------------- */
float SectionLayout::section_size() const { return std::ldexp(1.0f, levels - 1) * (float)coarse_res; }

std::vector<SectionLayout::Load> SectionLayout::initialize(int grid_, int coarse_res_, int levels_, const Vec3& cam) {
  grid = grid_, coarse_res = coarse_res_, levels = levels_;
  origin.assign((size_t)grid * grid * 2, 0.0f);
  tag.assign((size_t)grid * grid, 0);
  std::vector<Load> loads;
  const float top = std::ldexp(1.0f, levels - 1);
  for (int i = 0; i < grid; ++i)
    for (int j = 0; j < grid; ++j) { /* main.cpp:280-286 */
      float* o = &origin[((size_t)i * grid + j) * 2];
      o[0] = cam.x + ((float)i - (float)grid / 2.0f) * top * (float)coarse_res;
      o[1] = cam.z + ((float)j - (float)grid / 2.0f) * top * (float)coarse_res;
      tag[(size_t)i * grid + j] = i * grid + j;
      loads.push_back({i, j, i * grid + j, {o[0], o[1]}});
    }
  return loads;
}

std::vector<SectionLayout::Load> SectionLayout::manage(const Vec3& cam) {
  std::vector<Load> loads;
  if (grid < 2) return loads;
  const float size = section_size();
  auto org = [&](int i, int j) { return &origin[((size_t)i * grid + j) * 2]; };
  auto tg = [&](int i, int j) -> int& { return tag[(size_t)i * grid + j]; };
  auto move = [&](int di, int dj, int si, int sj) { /* point_sections[di][dj] = point_sections[si][sj] etc. (:336-339) */
    org(di, dj)[0] = org(si, sj)[0], org(di, dj)[1] = org(si, sj)[1];
    tg(di, dj) = tg(si, sj);
  };
  auto place = [&](int i, int j, int t, float ox, float oy) { /* allocateSection (:256-269) on a recycled section object */
    org(i, j)[0] = ox, org(i, j)[1] = oy;
    tg(i, j) = t;
    loads.push_back({i, j, t, {ox, oy}});
  };
  std::vector<int> freed((size_t)grid);
  /* allocate left, move sections right (:410-417) */
  if (cam.x < org(1, 0)[0]) {
    for (int j = 0; j < grid; ++j) freed[j] = tg(grid - 1, j); /* unloadSectionsColumn(size - 1) */
    for (int i = grid - 1; i >= 1; --i)                         /* rearrangeSectionsX(1), :334-345 */
      for (int j = 0; j < grid; ++j) move(i, j, i - 1, j);
    for (int j = 0; j < grid; ++j) place(0, j, freed[j], org(1, j)[0] - size, org(1, j)[1]);
  }
  /* allocate right, move sections left (:420-427) */
  if (cam.x >= org(grid - 1, grid - 1)[0]) {
    for (int j = 0; j < grid; ++j) freed[j] = tg(0, j);
    for (int i = 0; i < grid - 1; ++i) /* rearrangeSectionsX(-1), :348-361 */
      for (int j = 0; j < grid; ++j) move(i, j, i + 1, j);
    for (int j = 0; j < grid; ++j) place(grid - 1, j, freed[j], org(grid - 2, j)[0] + size, org(grid - 2, j)[1]);
  }
  /* allocate down, move sections up (:430-437) */
  if (cam.z < org(0, 1)[1]) {
    for (int i = 0; i < grid; ++i) freed[i] = tg(i, grid - 1);
    for (int i = 0; i < grid; ++i) /* rearrangeSectionsY(1), :371-383 */
      for (int j = grid - 1; j >= 1; --j) move(i, j, i, j - 1);
    for (int i = 0; i < grid; ++i) place(i, 0, freed[i], org(i, 1)[0], org(i, 1)[1] - size);
  }
  /* allocate up, move sections down (:440-447) */
  if (cam.z >= org(0, grid - 1)[1]) {
    for (int i = 0; i < grid; ++i) freed[i] = tg(i, 0);
    for (int i = 0; i < grid; ++i) /* rearrangeSectionsY(-1), :386-399 */
      for (int j = 0; j < grid - 1; ++j) move(i, j, i, j + 1);
    for (int i = 0; i < grid; ++i) place(i, grid - 1, freed[i], org(i, grid - 2)[0], org(i, grid - 2)[1] + size);
  }
  return loads;
}

SectionGrid::SectionGrid(hmrt_ctx* ctx, int coarse_res, int levels, int grid, bool with_colors, Loader loader)
    : ctx_(ctx), coarse_res_(coarse_res), levels_(levels), grid_(grid), loader_(std::move(loader)) {
  if (grid < 2 || !loader_) {
    status_ = HMRT_E_ARG;
    return;
  }
  for (int k = 0; k < grid * grid && status_ == 0; ++k) {
    sections_.emplace_back(new Heightmap(ctx, coarse_res, levels, with_colors));
    status_ = sections_.back()->status();
  }
}

int SectionGrid::fill(const std::vector<SectionLayout::Load>& loads) {
  int filled = 0;
  for (size_t k = 0; k < loads.size(); ++k) {
    const SectionLayout::Load& l = loads[k];
    bool superseded = false; /* a later shift of the same call recycled this section again: only its last origin counts */
    for (size_t m = k + 1; m < loads.size(); ++m) superseded |= loads[m].tag == l.tag;
    if (superseded) continue;
    ++filled;
    Heightmap& s = *sections_[(size_t)l.tag];
    int rc = s.clear(); /* `new float[n]()`, main.cpp:259-260 */
    if (rc == 0) rc = loader_(s, l.origin);
    if (rc == 0) rc = s.finish();
    if (rc) return rc > 0 ? -1000 - rc : rc;
  }
  return filled;
}

int SectionGrid::initialize(const Vec3& cam) {
  if (status_) return status_;
  return fill(layout_.initialize(grid_, coarse_res_, levels_, cam));
}

int SectionGrid::manage(const Vec3& cam) {
  if (status_) return status_;
  return fill(layout_.manage(cam));
}

int SectionGrid::prepare_window(const Vec3& cam, Heightmap& window, Vec3* camera_point_buffer) {
  if (status_) return status_;
  const float c[3] = {cam.x, cam.y, cam.z};
  hmrt_window_placement pl;
  int rc = hmrt_window_place(c, layout_.origin.data(), grid_, coarse_res_, levels_, &pl);
  if (rc) return rc;
  hmrt_window_sections secs;
  const int xs[2] = {pl.min_x, pl.max_x}, ys[2] = {pl.min_y, pl.max_y};
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      const Heightmap& s = section(xs[a], ys[b]);
      secs.d_pyramid[a][b] = s.d_pyramid();
      secs.d_color_map[a][b] = s.d_color_map();
    }
  rc = hmrt_compose_window(ctx_, &secs, coarse_res_, levels_, pl.cell_x, pl.cell_y, window.d_pyramid(), window.d_color_map());
  if (rc == 0 && camera_point_buffer) *camera_point_buffer = {pl.camera[0], pl.camera[1], pl.camera[2]};
  return rc;
}

/* ---------------------------------------------------------------------------------------------- */
static Vec3 normalize(Vec3 v) {
  const float s = 1.0f / std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
  return {v.x * s, v.y * s, v.z * s};
}
static Vec3 cross(Vec3 a, Vec3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
/* glm::rotate(v, angle, axis) as GLM 0.9.8.3 evaluates it (inc/glm/gtx/rotate_vector.inl:45-53): the upper 3x3 of
 * rotate(mat4(1), angle, axis) (gtc/matrix_transform.inl:19-47) times v (detail/type_mat3x3.inl:430-433), in that
 * operation order, so that a fly-through replays bit for bit (tests/test_refhost.py). */
static Vec3 rotate(Vec3 v, float angle, Vec3 normal) {
  const float c = std::cos(angle), s = std::sin(angle);
  const Vec3 axis = normalize(normal);
  const Vec3 temp = {(1.0f - c) * axis.x, (1.0f - c) * axis.y, (1.0f - c) * axis.z};
  float R[3][3]; /* Rotate[column][row] */
  R[0][0] = c + temp.x * axis.x;
  R[0][1] = temp.x * axis.y + s * axis.z;
  R[0][2] = temp.x * axis.z - s * axis.y;
  R[1][0] = temp.y * axis.x - s * axis.z;
  R[1][1] = c + temp.y * axis.y;
  R[1][2] = temp.y * axis.z + s * axis.x;
  R[2][0] = temp.z * axis.x + s * axis.y;
  R[2][1] = temp.z * axis.y - s * axis.x;
  R[2][2] = c + temp.z * axis.z;
  float M[3][3]; /* Result[k] = m[0] * R[k][0] + m[1] * R[k][1] + m[2] * R[k][2] with m = identity */
  for (int k = 0; k < 3; ++k) {
    M[k][0] = 1.0f * R[k][0] + 0.0f * R[k][1] + 0.0f * R[k][2];
    M[k][1] = 0.0f * R[k][0] + 1.0f * R[k][1] + 0.0f * R[k][2];
    M[k][2] = 0.0f * R[k][0] + 0.0f * R[k][1] + 1.0f * R[k][2];
  }
  return {M[0][0] * v.x + M[1][0] * v.y + M[2][0] * v.z, M[0][1] * v.x + M[1][1] * v.y + M[2][1] * v.z,
          M[0][2] * v.x + M[1][2] * v.y + M[2][2] * v.z};
}

void Camera::move(float fwd, float right, float up, float dt, const float boundaries[2], float max_height) {
  const float factor = dt > 1 ? 1 : dt; /* main.cpp:755 */
  const Vec3 f = normalize({forward.x, 0, forward.z});
  const Vec3 r = normalize(cross(forward, {0, 1, 0}));
  position.x += factor * (f.x * fwd + r.x * right);
  position.y += factor * (up + f.y * fwd + r.y * right);
  position.z += factor * (f.z * fwd + r.z * right);
  if (position.x < 0) position.x = 0; /* main.cpp:758-771 */
  if (position.x >= boundaries[0]) position.x = boundaries[0] - 0.00001f;
  if (position.z < 0) position.z = 0;
  if (position.z >= boundaries[1]) position.z = boundaries[1] - 0.00001f;
  if (position.y < 0) position.y = 0;
  if (position.y >= max_height * 4) position.y = max_height * 4;
}

void Camera::rotate(float yaw, float pitch, float dt) { /* main.cpp:779-780 */
  forward = hmrt_host::rotate(forward, yaw * dt, {0, 1, 0});
  forward = hmrt_host::rotate(forward, pitch * dt, normalize(cross(forward, {0, 1, 0})));
}

hmrt_camera Camera::abi() const {
  hmrt_camera c;
  c.frame_dim[0] = frame_dimension.x, c.frame_dim[1] = frame_dimension.y, c.frame_dim[2] = frame_dimension.z;
  c.forward[0] = forward.x, c.forward[1] = forward.y, c.forward[2] = forward.z;
  c.position[0] = position.x, c.position[1] = position.y, c.position[2] = position.z;
  return c;
}

/* ---------------------------------------------------------------------------------------------- */
Renderer::Renderer(hmrt_ctx* ctx, int width, int height) : ctx_(ctx), w_(width), h_(height) {
  cudaMalloc(&d_rgb_, (size_t)width * height * 3); /* the PBO of main.cpp:642-645 */
}
Renderer::~Renderer() { cudaFree(d_rgb_); }

int Renderer::set_heightmap(const Heightmap& hm, float max_height) {
  return hmrt_set_heightmap(ctx_, hm.d_pyramid(), hm.d_color_map(), hm.layout().coarse_res, hm.layout().levels, max_height);
}

int Renderer::render(const Camera& cam, const hmrt_trace_opts& opts) {
  if (!d_rgb_) return HMRT_E_NOMEM;
  const hmrt_camera c = cam.abi();
  int rc = hmrt_trace(ctx_, w_, h_, &c, 1, &opts, d_rgb_, nullptr);
  if (rc == 0) rc = hmrt_synchronize(ctx_);
  return rc;
}

int Renderer::render_to_host(const std::vector<Camera>& cams, const hmrt_trace_opts& opts, uint8_t* h_rgb) {
  std::vector<hmrt_camera> c(cams.size());
  for (size_t i = 0; i < cams.size(); ++i) c[i] = cams[i].abi();
  return hmrt_trace_host(ctx_, w_, h_, c.data(), (int)c.size(), &opts, h_rgb);
}

int Renderer::download(std::vector<uint8_t>& rgb) const {
  rgb.resize((size_t)w_ * h_ * 3);
  return (int)cudaMemcpy(rgb.data(), d_rgb_, rgb.size(), cudaMemcpyDeviceToHost);
}

bool write_ppm(const std::string& path, const uint8_t* rgb, int width, int height, bool flip) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) return false;
  std::fprintf(f, "P6\n%d %d\n255\n", width, height);
  for (int y = 0; y < height; ++y) {
    const int row = flip ? height - 1 - y : y;
    std::fwrite(rgb + (size_t)row * width * 3, 1, (size_t)width * 3, f);
  }
  return std::fclose(f) == 0;
}

}  // namespace hmrt_host

/* ---- plain-C views of the host helpers (used by the Python tests through ctypes) -------------- */
extern "C" {

int hmrt_host_pdg_generate(int n, uint64_t seed, float* out_xyz) {
  const std::vector<float> v = hmrt_host::pdg_generate(n, seed);
  if (v.empty()) return HMRT_E_ARG;
  std::memcpy(out_xyz, v.data(), v.size() * sizeof(float));
  return 0;
}

/* out[0..11] = scale xyz, offset xyz, min xyz, max xyz; meta = format, record_len, n_points, offset_to_points */
int hmrt_host_las_info(const char* path, double* out, uint64_t* meta) {
  hmrt_host::LasFile las;
  if (!hmrt_host::LasFile::open(path, las)) return HMRT_E_ARG;
  for (int i = 0; i < 3; ++i) out[i] = las.scale[i], out[3 + i] = las.offset[i], out[6 + i] = las.min[i], out[9 + i] = las.max[i];
  meta[0] = (uint64_t)las.point_format, meta[1] = (uint64_t)las.record_len, meta[2] = las.n_points, meta[3] = las.offset_to_points;
  return 0;
}

/* scene = cell_size xyz, boundaries xy, camera xyz, max_height (readLASHeader, main.cpp:153-164) */
int hmrt_host_las_scene(const char* path, float* scene) {
  hmrt_host::LasFile las;
  if (!hmrt_host::LasFile::open(path, las)) return HMRT_E_ARG;
  const hmrt_host::SceneInfo s = hmrt_host::read_las_header(las);
  scene[0] = s.cell_size[0], scene[1] = s.cell_size[1], scene[2] = s.cell_size[2];
  scene[3] = s.boundaries[0], scene[4] = s.boundaries[1];
  scene[5] = s.camera_position.x, scene[6] = s.camera_position.y, scene[7] = s.camera_position.z;
  scene[8] = s.max_height;
  return 0;
}

/* camera motion rules: state = position xyz, forward xyz */
void hmrt_host_camera_step(float* state, float fwd, float right, float up, float yaw, float pitch, float dt, const float* boundaries,
                           float max_height) {
  hmrt_host::Camera c;
  c.position = {state[0], state[1], state[2]};
  c.forward = {state[3], state[4], state[5]};
  c.move(fwd, right, up, dt, boundaries, max_height);
  c.rotate(yaw, pitch, dt);
  state[0] = c.position.x, state[1] = c.position.y, state[2] = c.position.z;
  state[3] = c.forward.x, state[4] = c.forward.y, state[5] = c.forward.z;
}

/* SectionLayout from Python: origins [grid][grid][2] and tags [grid][grid] in / out; loads (i, j, tag) triples out.
 * init != 0: initializeSections, else manageSections.  Returns the number of loads. */
int hmrt_host_section_layout_step(int init, int grid, int coarse_res, int levels, const float* cam, float* origins, int* tags,
                                  int* loads_ijt, float* load_origins, int max_loads) {
  hmrt_host::SectionLayout L;
  std::vector<hmrt_host::SectionLayout::Load> loads;
  const hmrt_host::Vec3 c = {cam[0], cam[1], cam[2]};
  if (init) {
    loads = L.initialize(grid, coarse_res, levels, c);
  } else {
    L.grid = grid, L.coarse_res = coarse_res, L.levels = levels;
    L.origin.assign(origins, origins + (size_t)grid * grid * 2);
    L.tag.assign(tags, tags + (size_t)grid * grid);
    loads = L.manage(c);
  }
  std::memcpy(origins, L.origin.data(), L.origin.size() * sizeof(float));
  std::memcpy(tags, L.tag.data(), L.tag.size() * sizeof(int));
  if ((int)loads.size() > max_loads) return HMRT_E_ARG;
  for (size_t k = 0; k < loads.size(); ++k) {
    loads_ijt[3 * k] = loads[k].i, loads_ijt[3 * k + 1] = loads[k].j, loads_ijt[3 * k + 2] = loads[k].tag;
    load_origins[2 * k] = loads[k].origin[0], load_origins[2 * k + 1] = loads[k].origin[1];
  }
  return (int)loads.size();
}

/* End-to-end driver of the section flow for the tests (needs a GPU): a SectionGrid whose loader rasterises one point per
 * cell with the exactly reproducible height ((wx * 73856093) ^ (wz * 19349663)) & 1023) / 8 at world cell (wx, wz) follows
 * the camera path cams[0..n) (initializeSections at cams[0], manageSections at every later position), then the window of the
 * last position is composed.  Out: the window pyramid, camera_point_buffer, the final origins and tags, total sections loaded. */
int hmrt_host_section_grid_run(int coarse_res, int levels, int grid, const float* cams, int n, float* out_window, float* out_cam_pb,
                               float* out_origins, int* out_tags, int* out_loaded) {
  using namespace hmrt_host;
  if (n < 1) return HMRT_E_ARG;
  hmrt_ctx* ctx = nullptr;
  int rc = hmrt_create(0, &ctx);
  if (rc) return rc;
  {
    const PyramidLayout lay(coarse_res, levels);
    const int r0 = lay.finest();
    auto loader = [&](Heightmap& sec, const float origin[2]) {
      std::vector<float> xyz((size_t)r0 * r0 * 3);
      for (int z = 0; z < r0; ++z)
        for (int x = 0; x < r0; ++x) {
          const long long wx = (long long)std::floor(origin[0]) + x, wz = (long long)std::floor(origin[1]) + z;
          float* p = &xyz[((size_t)z * r0 + x) * 3];
          p[0] = origin[0] + (float)x + 0.5f, p[1] = origin[1] + (float)z + 0.5f;
          p[2] = (float)((((unsigned long long)wx * 73856093ull) ^ ((unsigned long long)wz * 19349663ull)) & 1023ull) / 8.0f;
        }
      const float cell[3] = {1.f, 1.f, 1.f};
      return sec.rasterise_xyz(xyz, cell, origin);
    };
    SectionGrid sg(ctx, coarse_res, levels, grid, false, loader);
    Heightmap window(ctx, coarse_res, levels, false);
    rc = sg.status() ? sg.status() : window.status();
    int loaded = 0;
    for (int k = 0; k < n && rc == 0; ++k) {
      const Vec3 c = {cams[3 * k], cams[3 * k + 1], cams[3 * k + 2]};
      const int r = k == 0 ? sg.initialize(c) : sg.manage(c);
      if (r < 0) rc = r; else loaded += r;
    }
    Vec3 pb{0, 0, 0};
    if (rc == 0) rc = sg.prepare_window({cams[3 * (n - 1)], cams[3 * (n - 1) + 1], cams[3 * (n - 1) + 2]}, window, &pb);
    if (rc == 0) rc = hmrt_synchronize(ctx);
    if (rc == 0) rc = (int)cudaMemcpy(out_window, window.d_pyramid(), sizeof(float) * (size_t)lay.total, cudaMemcpyDeviceToHost);
    if (rc == 0) {
      out_cam_pb[0] = pb.x, out_cam_pb[1] = pb.y, out_cam_pb[2] = pb.z;
      std::memcpy(out_origins, sg.layout().origin.data(), sg.layout().origin.size() * sizeof(float));
      std::memcpy(out_tags, sg.layout().tag.data(), sg.layout().tag.size() * sizeof(int));
      *out_loaded = loaded;
    }
  }
  hmrt_destroy(ctx);
  return rc;
}

}  // extern "C"
