/*
 * hmrt_render -- headless counterpart of the reference application's main loop
 * (GPUHeightmapRaytracer/src/main.cpp:947-972, :1043-1087): load or generate point data, rasterise
 * it into the max-mipmap pyramid on the GPU, fly the camera, write PPM frames.
 *
 *   hmrt_render --generate 1024 --seed 7 --width 640 --height 480 --out frame        (BASELINE config 1 input)
 *   hmrt_render --las ../Data/autzen.las --width 1920 --height 1080 --frames 8 --shadows --out fly
 *   hmrt_render --pdg ../Data/data --grid 1024 --out terrain
 *   hmrt_render --las cloud.las --sections 4 --coarse 32 --frames 8 --out win       (the reference's section grid + window)
 */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "hmrt_host.hpp"

using namespace hmrt_host;

static int die(const char* what, int code) {
  std::fprintf(stderr, "hmrt_render: %s: %d (%s)\n", what, code, hmrt_error_string(code));
  return 1;
}

int main(int argc, char** argv) {
  std::string las_path, pdg_path, out = "frame";
  int generate = 0, grid = 0, levels = 8, width = 640, height = 480, frames = 1, device = 0, sections = 0, coarse_arg = 32;
  uint64_t seed = 1;
  bool shadows = false, colors = false;
  for (int i = 1; i < argc; ++i) {
    auto arg = [&](const char* name) { return std::strcmp(argv[i], name) == 0 && i + 1 < argc; };
    if (arg("--las")) las_path = argv[++i];
    else if (arg("--pdg")) pdg_path = argv[++i];
    else if (arg("--generate")) generate = std::atoi(argv[++i]);
    else if (arg("--grid")) grid = std::atoi(argv[++i]);
    else if (arg("--seed")) seed = std::strtoull(argv[++i], nullptr, 0);
    else if (arg("--levels")) levels = std::atoi(argv[++i]);
    else if (arg("--width")) width = std::atoi(argv[++i]);
    else if (arg("--height")) height = std::atoi(argv[++i]);
    else if (arg("--frames")) frames = std::atoi(argv[++i]);
    else if (arg("--device")) device = std::atoi(argv[++i]);
    else if (arg("--sections")) sections = std::atoi(argv[++i]);
    else if (arg("--coarse")) coarse_arg = std::atoi(argv[++i]);
    else if (arg("--out")) out = argv[++i];
    else if (std::strcmp(argv[i], "--shadows") == 0) shadows = true;
    else if (std::strcmp(argv[i], "--colors") == 0) colors = true;
    else {
      std::fprintf(stderr, "usage: hmrt_render (--las F | --pdg F --grid N | --generate N [--seed S]) [--levels L] [--width W] [--height H] "
                           "[--frames N] [--shadows] [--colors] [--device D] [--out PREFIX] [--sections G --coarse C  (LAS only: G x G resident sections of "
                           "C * 2^(L-1) cells around the camera, one window per frame, main.cpp:276-625)]\n");
      return 2;
    }
  }

  hmrt_ctx* ctx = nullptr;
  int rc = hmrt_create(device, &ctx);
  if (rc) return die("hmrt_create", rc);

  /* ---- point data -> grid dimensions ---- */
  LasFile las;
  std::vector<float> xyz;
  SceneInfo scene;
  float cell[3] = {2.f, 2.f, 2.f}, origin[2] = {0.f, 0.f};
  int r0 = 0;
  if (!las_path.empty()) {
    std::string err;
    if (!LasFile::open(las_path, las, &err)) {
      std::fprintf(stderr, "Error opening %s: %s\n", las_path.c_str(), err.c_str()); /* main.cpp:129-133 */
      return 1;
    }
    scene = read_las_header(las);
    std::printf("LAS File Loaded.\nPoints count: %llu\nScale: %g %g %g\n", (unsigned long long)las.n_points, las.scale[0], las.scale[1], las.scale[2]);
    const float extent = std::max(scene.boundaries[0], scene.boundaries[1]);
    r0 = 1 << (levels - 1);
    while (r0 < extent) r0 <<= 1; /* the reference keeps 4x4 sections of 4096^2; here the whole cloud is resident */
  } else {
    if (!pdg_path.empty()) {
      if (!read_pdg_text(pdg_path, xyz)) return die("read_pdg_text", HMRT_E_ARG);
    } else {
      if (generate <= 0) generate = 1024;
      xyz = pdg_generate(generate, seed);
      if (xyz.empty()) return die("pdg_generate (N must be a power of two)", HMRT_E_ARG);
      grid = generate;
    }
    if (grid <= 0) return die("--grid N is required with --pdg", HMRT_E_ARG);
    r0 = grid;
    cell[0] = cell[1] = cell[2] = 1.f;
    scene.boundaries[0] = scene.boundaries[1] = (float)r0;
  }
  /* ---- the reference's own flow: a grid of sections around the camera, one composed window per frame (main.cpp:947-966) ---- */
  if (sections > 0) {
    if (las_path.empty()) return die("--sections needs --las", HMRT_E_ARG);
    SectionGrid sg(ctx, coarse_arg, levels, sections, colors, [&](Heightmap& sec, const float org[2]) {
      return sec.rasterise_las(las, scene.cell_size, org); /* loadLASToSection(file, origin, ...), main.cpp:174 */
    });
    if (!sg.ok()) return die("SectionGrid", sg.status());
    Heightmap window(ctx, coarse_arg, levels, colors); /* d_point_buffer + d_color_map, main.cpp:1012-1013 */
    if (!window.ok()) return die("window", window.status());
    Camera cam;
    cam.position = scene.camera_position;
    rc = sg.initialize(cam.position);
    if (rc < 0) return die("initializeSections", rc);
    std::printf("%d x %d sections of %d^2 cells loaded\n", sections, sections, window.layout().finest());
    Renderer renderer(ctx, width, height);
    hmrt_trace_opts opts;
    hmrt_trace_opts_default(&opts, scene.max_height);
    opts.use_color_map = colors ? 1 : 0;
    opts.shadows = shadows ? 1 : 0;
    opts.light_dir[0] = 0.3244f, opts.light_dir[1] = 0.8111f, opts.light_dir[2] = 0.4867f;
    std::vector<uint8_t> rgb;
    for (int f = 0; f < frames; ++f) {
      const auto f0 = std::chrono::steady_clock::now();
      rc = sg.manage(cam.position); /* manageSections */
      if (rc < 0) return die("manageSections", rc);
      const int reloaded = rc;
      Camera wc = cam; /* camera_point_buffer: the camera in window coordinates */
      rc = sg.prepare_window(cam.position, window, &wc.position); /* preparePointBuffer + copyPointBuffer */
      if (rc) return die("preparePointBuffer", rc);
      rc = renderer.set_heightmap(window, scene.max_height);
      if (rc == 0) rc = renderer.render(wc, opts);
      if (rc) return die("rayTrace", rc);
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - f0).count();
      rc = renderer.download(rgb);
      if (rc) return die("download", rc);
      char name[512];
      std::snprintf(name, sizeof name, "%s_%04d.ppm", out.c_str(), f);
      if (!write_ppm(name, rgb.data(), width, height)) return die("write_ppm", HMRT_E_ARG);
      std::printf("FPS: %.1f  Pos: %.1f %.1f %.1f  window %.1f %.1f  sections reloaded %d -> %s\n", 1.0 / dt, cam.position.x, cam.position.y,
                  cam.position.z, wc.position.x, wc.position.z, reloaded, name);
      cam.move(250.f, 0.f, 0.f, 0.1f, scene.boundaries, scene.max_height);
      cam.rotate(1.f, 0.f, 0.1f);
    }
    hmrt_destroy(ctx);
    return 0;
  }

  if (r0 % (1 << (levels - 1))) return die("grid size must be a multiple of 2^(levels-1)", HMRT_E_SHAPE);
  const int coarse = r0 >> (levels - 1);

  /* ---- rasterise (GPU scatter + max-mipmap build) ---- */
  const auto t0 = std::chrono::steady_clock::now();
  Heightmap hm(ctx, coarse, levels, colors && !las_path.empty());
  if (!hm.ok()) return die("Heightmap", hm.status());
  rc = las_path.empty() ? hm.rasterise_xyz(xyz, cell, origin) : hm.rasterise_las(las, cell, origin);
  if (rc) return die("rasterise", rc);
  rc = hm.finish();
  if (rc) return die("build_mips", rc);
  float max_height = 0;
  rc = hm.max_height(&max_height);
  if (rc) return die("max_height", rc);
  const double raster_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  const unsigned long long n_points = las_path.empty() ? xyz.size() / 3 : las.n_points;
  std::printf("grid %d^2, %d levels, %llu points rasterised in %.3f s, max height %.3f cells\n", r0, levels, n_points, raster_s, max_height);
  if (las_path.empty()) {
    scene.max_height = max_height;
    scene.camera_position = {r0 * 0.5f, 1.5f * max_height, r0 * 0.5f};
  }

  /* ---- render a short fly-through ---- */
  Renderer renderer(ctx, width, height);
  rc = renderer.set_heightmap(hm, scene.max_height);
  if (rc) return die("set_heightmap", rc);
  Camera cam;
  cam.position = scene.camera_position;
  hmrt_trace_opts opts;
  hmrt_trace_opts_default(&opts, scene.max_height);
  opts.use_color_map = (colors && hm.d_color_map()) ? 1 : 0;
  opts.shadows = shadows ? 1 : 0;
  opts.light_dir[0] = 0.3244f, opts.light_dir[1] = 0.8111f, opts.light_dir[2] = 0.4867f;
  std::vector<uint8_t> rgb;
  for (int f = 0; f < frames; ++f) {
    const auto f0 = std::chrono::steady_clock::now();
    rc = renderer.render(cam, opts);
    if (rc) return die("rayTrace", rc);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - f0).count();
    rc = renderer.download(rgb);
    if (rc) return die("download", rc);
    char name[512];
    std::snprintf(name, sizeof name, "%s_%04d.ppm", out.c_str(), f);
    if (!write_ppm(name, rgb.data(), width, height)) return die("write_ppm", HMRT_E_ARG);
    std::printf("FPS: %.1f  Pos: %.1f %.1f %.1f  -> %s\n", 1.0 / dt, cam.position.x, cam.position.y, cam.position.z, name); /* drawFPS, main.cpp:803-806 */
    cam.move(250.f, 0.f, 0.f, 0.1f, scene.boundaries, scene.max_height); /* wasd_movement_distance, main.cpp:98 */
    cam.rotate(1.f, 0.f, 0.1f);                                          /* jl_rotation_angle, main.cpp:104-105 */
  }
  hmrt_destroy(ctx);
  return 0;
}
