/*
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load these
 * libraries, and only as the checker or the reported CPU baseline.
 *
 * Two CPU libraries share this interface (struct types come from include/hmrt.h):
 *
 *   oracle/_ref/libhmrt_ref.so      the REFERENCE'S OWN device code, lines 1-286 of
 *       (hmrt_ref_*)                GPUHeightmapRaytracer/src/CudaKernel.cu compiled for the host
 *                                   from where it lies under /root/reference by build_ref.sh
 *                                   (float-pow = original MSVC meaning; *_dpow.so = g++ meaning).
 *   oracle/_build/libhmrt_oracle.so plain-C restatement (hmrt_oracle.c), each function citing
 *       (hmrt_oracle_*)             the reference lines it follows; validated against _ref by
 *                                   tests/test_oracle_vs_ref.py and the golden fixtures.
 */
#ifndef HMRT_ORACLE_H_
#define HMRT_ORACLE_H_

#include "../include/hmrt.h"

#ifdef __cplusplus
extern "C" {
#endif

/*
 * Render rows [row_begin,row_end) of one W x H frame on n_threads host threads (rows
 * interleaved).  rgb is indexed like the full frame ((px + py*W)*3); hits (optional) likewise.
 * opts->tile_* are ignored.  Returns 0, or <0 on bad arguments; the _ref build returns -100 if
 * its instrumented path (glue around the reference's castRay) ever disagrees with the
 * reference's verbatim cuda_rayTrace colours.
 */
int hmrt_ref_trace(const float* pyramid, const hmrt_color* color_map, int coarse_res, int levels,
                   int W, int H, const hmrt_camera* cam, const hmrt_trace_opts* opts,
                   int n_threads, int row_begin, int row_end, uint8_t* rgb, hmrt_hit* hits);
/* 1 when built with the fp32 pow overload (MSVC meaning), 0 for the g++ double meaning. */
int hmrt_ref_float_math(void);

int hmrt_oracle_trace(const float* pyramid, const hmrt_color* color_map, int coarse_res,
                      int levels, int W, int H, const hmrt_camera* cam,
                      const hmrt_trace_opts* opts, int n_threads, int row_begin, int row_end,
                      uint8_t* rgb, hmrt_hit* hits);

/* == main.cpp:995-1003 / CudaKernel.cu:250-258 */
int hmrt_oracle_pyramid_layout(int coarse_res, int levels, int* res, int64_t* idx,
                               int64_t* total);

/*
 * Restatement of the loadLASToSection inner loop (main.cpp:193-234) over n LAS records
 * (formats 0-3), applied in file order to a caller-zeroed pyramid + colour map -- i.e. it
 * updates ALL levels with the reference's early-break max propagation, so its output is the
 * oracle for scatter + mip build together.  color_map may be NULL.
 */
int hmrt_oracle_rasterise_las(const uint8_t* records, int64_t n, int record_len,
                              int point_format, const hmrt_las_transform* xf, float* pyramid,
                              int coarse_res, int levels, hmrt_color* color_map);
int hmrt_oracle_rasterise_xyz(const float* xyz, int64_t n, const hmrt_las_transform* xf,
                              float* pyramid, int coarse_res, int levels);
/* "level i+1 = max of 2x2 children" from the finest level (property form of main.cpp:227-233) */
int hmrt_oracle_build_mips(float* pyramid, int coarse_res, int levels);

/*
 * Seeded restatement of PointdataGenerator (PointdataGenerator/main.cpp:72-184): diamond-square
 * on a (n+1)^2 grid, then z *= 10 except the last row and column.  out = (n+1)^2 (x,y,z) float
 * triples in the generator's row order.  The PRNG is a fixed 64-bit LCG (the reference's
 * std::default_random_engine(time(NULL)) is neither seeded nor portable).
 */
int hmrt_oracle_pdg_generate(int n, uint64_t seed, float* out_xyz);

/*
 * preparePointBuffer (main.cpp:459-618) restated: the host arithmetic (:461-516) and the four memcpy loops per level
 * (:519-567) + the colour loops (:570-618), on HOST buffers.  sections[x][y]: x 0 = minX, 1 = maxX; y 0 = minY, 1 = maxY.
 */
int hmrt_oracle_window_place(const float camera_position[3], const float* section_origins, int grid, int coarse_res, int levels,
                             hmrt_window_placement* out);
int hmrt_oracle_compose_window(const float* const sections[2][2], const hmrt_color* const colors[2][2], int coarse_res, int levels,
                               int cell_x, int cell_y, float* h_point_buffer, hmrt_color* h_color_map);

#ifdef __cplusplus
}
#endif
#endif
