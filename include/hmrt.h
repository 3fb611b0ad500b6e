/*
 * hmrt.h -- C ABI of the B200-native heightfield ray traversal + point rasterisation path.
 *
 * This is the drop-in boundary.  Every entry point names the
 * reference interface it replaces; paths are relative to the reference tree
 * (GPUHeightmapRaytracer/src/...).  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions
 *   - every function returns 0 on success, a positive cudaError_t value when the CUDA
 *     runtime failed, or a negative HMRT_E_* code for argument errors.  Nothing exits the
 *     process (the reference's checkCudaErrors logs and continues, inc/helper_cuda.h:985-995).
 *   - "d_" pointers are device pointers on the context's device, "h_" pointers are host
 *     pointers.  Device pointers are BORROWED exactly as in the reference: the caller
 *     allocates, fills and frees them (main.cpp:1012-1013, :623-624, :1022-1023).
 *   - a context is bound to one device and one stream and is not re-entrant (the reference
 *     calls its three entry points from one thread only, main.cpp:947-966,1079).
 *   - pyramid layout = the reference's: `levels` square grids, coarsest level at float
 *     offset 0, finest level last; level i has resolution coarse_res * 2^(levels-1-i) and
 *     starts at idx[i] = idx[i+1] + res[i+1]^2 (main.cpp:995-1003, CudaKernel.cu:250-258).
 *   - framebuffer = packed RGB8, row-major, pixel (px,py) at (px + py*W)*3, py = 0 is the
 *     first row written (GL bottom row in the reference, CudaKernel.cu:202,219-221).
 */
#ifndef HMRT_H_
#define HMRT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HMRT_VERSION 201 /* 200: hmrt_trace_opts grew by full_frame_output; hmrt_rx_*, hmrt_ipc_*, hmrt_trace_stats, hmrt_get_stream added; 201: hmrt_trace_host_begin / _wait */
#define HMRT_MAX_LEVELS 16

/* argument errors (negative so they never collide with cudaError_t) */
#define HMRT_E_ARG (-1)      /* null / out-of-range argument */
#define HMRT_E_STATE (-2)    /* call order: e.g. trace before set_heightmap */
#define HMRT_E_SHAPE (-3)    /* grid does not tile / too large for 32-bit cell indices */
#define HMRT_E_NOMEM (-4)    /* host allocation failed */
#define HMRT_E_NCCL (-5)     /* NCCL missing (libnccl.so.2 not loadable) or a NCCL call failed */

typedef struct hmrt_ctx hmrt_ctx;

/* == struct CudaSpace::Color (CudaKernel.cuh:37-48): 3 bytes, no padding. */
typedef struct hmrt_color {
  uint8_t r, g, b;
} hmrt_color;

/* The per-frame arguments of CudaSpace::rayTrace (CudaKernel.cuh:49, call site main.cpp:686). */
typedef struct hmrt_camera {
  float frame_dim[3]; /* frame_dimensions: image-plane width, height, distance (main.cpp:57) */
  float forward[3];   /* camera_forward: unit length, not parallel to +y (CudaKernel.cu:237) */
  float position[3];  /* grid_camera_position, finest-cell units (main.cpp:514-517) */
} hmrt_camera;

/*
 * Optional per-pixel traversal record (parity instrumentation; the reference has no such
 * output -- it is what castRay leaves in its by-reference `ray_position`, CudaKernel.cu:121).
 * x,y,z: ray position when castRay returned, in castRay's MIRRORED space (CudaKernel.cu:130-150).
 * flags: HMRT_HIT_* bits | (loop iterations, primary + shadow segments) << HMRT_HIT_STEPS_SHIFT.
 */
typedef struct hmrt_hit {
  float x, y, z;
  uint32_t flags;
} hmrt_hit;
#define HMRT_HIT_HIT 1u       /* castRay returned from the finest level (CudaKernel.cu:161-167) */
#define HMRT_HIT_MIRROR_X 2u  /* ray_direction.x was negative (CudaKernel.cu:130-135) */
#define HMRT_HIT_MIRROR_Z 4u  /* ray_direction.z was negative (CudaKernel.cu:141-146) */
#define HMRT_HIT_SHADOWED 8u  /* shadow segment hit the terrain (extension, see hmrt_trace_opts) */
#define HMRT_HIT_STEPS_SHIFT 8

/*
 * Options of a trace call.  Zero-initialise, then set what you need; hmrt_trace_opts_default()
 * fills the reference's behaviour (no shadows, whole frame).
 */
typedef struct hmrt_trace_opts {
  int use_color_map; /* rayTrace's `use_color` (CudaKernel.cu:163-166) */
  float max_height;  /* rayTrace's `max_height` (CudaKernel.cu:41,153) */
  /* Shadow rays (north star; the reference has none, semantics defined in DESIGN.md section 5):
   * from the primary hit point, stepped back by shadow_bias along the primary ray, a second
   * castRay toward light_dir; if it hits, every colour channel is halved (c >> 1). */
  int shadows;
  float light_dir[3]; /* unit vector TOWARD the light */
  float shadow_bias;  /* cells; 0 selects the default 1/16 */
  /* Row-tile sharding (multi-GPU): the frame is cut into tiles of
   * HMRT_ROW_TILE rows; this call renders tiles tile_first, tile_first + tile_stride, ...
   * and stores local tile j at rows [j*HMRT_ROW_TILE, ...) of the output.  first=0, stride=1
   * (or stride 0) = whole frame in natural order. */
  int tile_first;
  int tile_stride;
  /* 0: the output holds only the rows this call renders (compact, see above).  1: the output is a WHOLE frame (H rows per
   * frame) and every rendered tile is stored at its place in it -- the buffer may be another GPU's memory (a peer mapping,
   * see hmrt_ipc_*): the ranks of a row-tile-sharded render then assemble the frame the reference delivers per call
   * (main.cpp:675-703) with their own stores, over NVLink, inside the traversal kernel.  Not for hmrt_trace_host. */
  int full_frame_output;
} hmrt_trace_opts;
#define HMRT_ROW_TILE 8

/* LAS public-header fields the rasteriser needs (liblas::Header accessors used at
 * main.cpp:137-164,187-202): X = raw*scale + offset in double (libLAS 1.8.0 Point::GetX). */
typedef struct hmrt_las_transform {
  double scale[3];
  double offset[3];
  double min[3];        /* header.GetMinX/Y/Z (main.cpp:200-202) */
  float cell_size[3];   /* cell_size (main.cpp:154-155: 2.0 in the reference) */
  float origin[2];      /* section origin in cell units (main.cpp:174, :205-206) */
} hmrt_las_transform;

/* ---- context ---------------------------------------------------------------------------- */

/* Bind a context to CUDA device `device` (replaces cudaGLSetGLDevice(gpuGetMaxGflopsDeviceId()),
 * main.cpp:1008).  Creates no stream: work runs on the stream set by hmrt_set_stream
 * (default: the legacy default stream, like the reference). */
int hmrt_create(int device, hmrt_ctx** out);
int hmrt_destroy(hmrt_ctx* ctx);
/* `cuda_stream` is a cudaStream_t passed as void* (0 = default stream). */
int hmrt_set_stream(hmrt_ctx* ctx, void* cuda_stream);
/* The stream set by hmrt_set_stream (hosts that stage their own copies in front of a call order them on it). */
void* hmrt_get_stream(const hmrt_ctx* ctx);
/* Block until all work queued by this context has finished (the reference synchronises inside
 * every call, CudaKernel.cu:301,307,316,325; here calls are asynchronous unless stated). */
int hmrt_synchronize(hmrt_ctx* ctx);
const char* hmrt_error_string(int code);
int hmrt_version(void);

/* ---- pyramid layout (main.cpp:995-1003 == CudaKernel.cu:250-258) ----------------------- */

/* Fills res[i], idx[i] for i in [0,levels) (level 0 = finest) and the total float count
 * (= stride_x * coarse_res^2).  Any output pointer may be NULL. */
int hmrt_pyramid_layout(int coarse_res, int levels, int* res, int64_t* idx, int64_t* total);

/* ---- ray traversal ---------------------------------------------------------------------- */

/* == CudaSpace::initializeDeviceVariables (CudaKernel.cuh:50, CudaKernel.cu:313-317,245-272).
 * d_pyramid / d_color_map are borrowed; d_color_map may be NULL when use_color_map is never
 * set.  The reference also takes texture_res and stride_x; the resolution is a per-call
 * argument here and stride_x is implied by (coarse_res, levels). */
int hmrt_set_heightmap(hmrt_ctx* ctx, const float* d_pyramid, const hmrt_color* d_color_map,
                       int coarse_res, int levels, float max_height);

void hmrt_trace_opts_default(hmrt_trace_opts* opts, float max_height);

/* Up to HMRT_MAX_CALLS_IN_FLIGHT hmrt_trace calls may be in flight at once when the caller switches streams between them
 * (hmrt_set_stream): every call takes its own set of work counters, so a renderer that double-buffers its frames overlaps
 * the drain of one call with the head of the next (a call issued on another stream than the previous one is assumed to
 * overlap with it and takes the launch shape that is best for that).  Calls on ONE stream simply serialise. */
#define HMRT_MAX_CALLS_IN_FLIGHT 4

/* == CudaSpace::rayTrace (CudaKernel.cuh:49, CudaKernel.cu:291-308) + cuda_setParameters
 * (CudaKernel.cu:227-240) + cuda_rayTrace (CudaKernel.cu:195-222) for n_frames cameras in ONE
 * launch.  Frame f is written at d_rgb + f * rows_local * W * 3 (rows_local = rows selected by
 * the tile options; = H for a whole frame).  d_hits (optional, same indexing, one record per
 * pixel) selects the instrumented kernel.  Asynchronous on the context's stream. */
int hmrt_trace(hmrt_ctx* ctx, int W, int H, const hmrt_camera* h_cameras, int n_frames,
               const hmrt_trace_opts* opts, uint8_t* d_rgb, hmrt_hit* d_hits);

/* Same call with HOST output: renders into a context-owned device framebuffer and copies the
 * result to h_rgb (pinned memory recommended); synchronous, like the reference's rayTrace +
 * the GL read-back it feeds (main.cpp:675-703). */
int hmrt_trace_host(hmrt_ctx* ctx, int W, int H, const hmrt_camera* h_cameras, int n_frames,
                    const hmrt_trace_opts* opts, uint8_t* h_rgb);

/* The same call in two halves, for a double-buffered renderer: _begin returns once the work is enqueued (the cameras have
 * been consumed; h_rgb must stay valid and untouched), _wait returns when the OLDEST call begun and not yet waited for has
 * delivered all its frames.  Up to HMRT_MAX_HOST_CALLS_IN_FLIGHT calls may be in flight, each rendering into its own
 * context-owned device framebuffer: the traversal of call k + 1 runs under the device->host copies of call k (the
 * reference serialises trace and read-back every frame, main.cpp:947-966).  HMRT_E_STATE: _begin with the maximum already
 * in flight, _wait with nothing in flight.  hmrt_trace_host == _begin + collect everything. */
#define HMRT_MAX_HOST_CALLS_IN_FLIGHT 2
int hmrt_trace_host_begin(hmrt_ctx* ctx, int W, int H, const hmrt_camera* h_cameras, int n_frames,
                          const hmrt_trace_opts* opts, uint8_t* h_rgb);
int hmrt_trace_host_wait(hmrt_ctx* ctx);

/* Number of local rows selected by (H, tile_first, tile_stride). */
int hmrt_rows_local(int H, int tile_first, int tile_stride);

/* == CudaSpace::freeDeviceVariables (CudaKernel.cuh:51): forget the borrowed pointers. */
int hmrt_clear_heightmap(hmrt_ctx* ctx);

/* ---- point rasterisation (loadLASToSection inner loop, main.cpp:193-234) ---------------- */

/* Zero a pyramid (the reference's `new float[n]()`, main.cpp:259) and, if given, the colour
 * keys / colour map (`new Color[n]` default-constructs to 0,0,0, main.cpp:260). */
int hmrt_clear_section(hmrt_ctx* ctx, float* d_pyramid, int coarse_res, int levels,
                       uint64_t* d_color_keys, hmrt_color* d_color_map);

/* Bin LAS point records into the FINEST level of d_pyramid with an order-independent
 * float-as-int atomic max (heights are >= +0, main.cpp:227-233 + :259).
 *   d_records  n records of record_len bytes each, LAS 1.2 point data formats 0-3
 *   point_format  0..3 (selects where classification / RGB live)
 *   first_index  global index of record 0 (file order), used for the colour keys (first_index + n < 2^39)
 *   d_color_keys  optional, res0^2 uint64: keeps (index+1)<<24 | rgb of the LAST point in file
 *                 order per cell (the reference's last-writer-wins, main.cpp:223-224)
 * Points outside [0,res0)^2 or with classification 7 are skipped (main.cpp:209).
 * Asynchronous on the context's stream, except that in scatter mode 0 the FIRST call of an input (first_index == 0) with
 * >= 4 M points on a grid larger than L2 reads back a 4 KB locality probe (one stream synchronisation) to choose its path;
 * later chunks of the same input (first_index > 0) reuse the verdict. */
int hmrt_scatter_las(hmrt_ctx* ctx, const uint8_t* d_records, int64_t n, int record_len,
                     int point_format, const hmrt_las_transform* xf, int64_t first_index,
                     float* d_pyramid, int coarse_res, int levels, uint64_t* d_color_keys);

/* Rasterisation strategy of hmrt_scatter_las: 0 (default) = decide per call from a locality probe of the
 * input, 1 = one atomic max per point straight into the grid (best for survey-ordered files and small grids),
 * 2 = bin the points by grid tile (tiles of up to 2048 x 2048 cells, 8 to 16 per axis) first (best for unordered clouds on grids larger than L2; needs 16-byte
 *     aligned records of at most 64 bytes, else the call takes path 1; colour keys travel with the points).
 * Results are bit-identical in every mode. */
int hmrt_set_scatter_mode(hmrt_ctx* ctx, int mode);

/* Same for PointdataGenerator output (PointdataGenerator/main.cpp:186-205): n float32
 * (x, y, z) triples, treated as LAS coordinates with scale 1 and offset 0. */
int hmrt_scatter_xyz(hmrt_ctx* ctx, const float* d_xyz, int64_t n, const hmrt_las_transform* xf,
                     float* d_pyramid, int coarse_res, int levels);

/* Build every coarser level from the finest one: level i+1 = max of its 2x2 children
 * (the fixed point of main.cpp:227-233). */
int hmrt_build_mips(hmrt_ctx* ctx, float* d_pyramid, int coarse_res, int levels);

/* Turn colour keys into the reference's colour map (cells never written stay 0,0,0). */
int hmrt_resolve_colors(hmrt_ctx* ctx, const uint64_t* d_color_keys, hmrt_color* d_color_map,
                        int64_t n_cells);

/* ---- camera window over resident sections (preparePointBuffer + copyPointBuffer, main.cpp:459-625) ------ */

/*
 * The reference keeps a grid of sections (each one a pyramid + colour map of its own, main.cpp:256-269) in HOST memory and,
 * every frame, memcpy's the camera-centred window out of the up-to-four sections it straddles into a host point buffer,
 * level by level (preparePointBuffer, main.cpp:459-618), then uploads 139.8 MB over PCIe (copyPointBuffer, :620-625).
 * Here the sections are resident in device memory and the window is composed on the device.
 *
 * hmrt_window_place is the host arithmetic of preparePointBuffer (:461-514): which sections the window straddles, the
 * coarsest-level cell of the lower-left one where the window starts (cell_position, :508) and the camera in window
 * coordinates (camera_point_buffer, :511-514).  `section_origins` is the reference's point_sections_origins[i][j]
 * flattened as [(i * grid + j) * 2 + {x, y}].  HMRT_E_ARG when the window leaves the grid (the reference would index out
 * of bounds; its manageSections keeps the camera inside the inner sections).
 */
typedef struct hmrt_window_placement {
  int min_x, min_y, max_x, max_y; /* section indices (main.cpp:472-502) */
  int cell_x, cell_y;             /* cell_position at the COARSEST level (main.cpp:508) */
  float camera[3];                /* camera_point_buffer (main.cpp:511-514) */
} hmrt_window_placement;
int hmrt_window_place(const float camera_position[3], const float* section_origins, int grid, int coarse_res, int levels,
                      hmrt_window_placement* out);

/* The four sections under the window: [0][*] = left (min_x), [1][*] = right (max_x), [*][0] = bottom (min_y),
 * [*][1] = top (max_y); entries may alias.  Colour maps: all NULL, or one per section. */
typedef struct hmrt_window_sections {
  const float* d_pyramid[2][2];
  const hmrt_color* d_color_map[2][2];
} hmrt_window_sections;

/* == the copy loops of preparePointBuffer (main.cpp:516-618) + copyPointBuffer (:620-625), on the device: at every level i,
 * with c = cell << (levels-1-i), window(x, y) = section[x + c.x >= res_i][y + c.y >= res_i]((x + c.x) % res_i, (y + c.y) % res_i);
 * the colour map likewise at the finest resolution.  d_window_color_map may be NULL.  Asynchronous on the context's stream. */
int hmrt_compose_window(hmrt_ctx* ctx, const hmrt_window_sections* sections, int coarse_res, int levels, int cell_x, int cell_y,
                        float* d_window_pyramid, hmrt_color* d_window_color_map);

/* ---- device buffers other processes on the same box can map (cudaIpc; NVLink peer access) -----------------------------
 * hmrt_ipc_alloc: a device buffer on the context's device + its 64-byte handle; hmrt_ipc_open (in ANOTHER process): map
 * it on this context's device; hmrt_ipc_close / hmrt_ipc_free undo them.  Used for the frame of a row-tile-sharded render
 * (hmrt_trace_opts.full_frame_output): rank 0 allocates the frame, every other rank opens it and renders straight into it. */
int hmrt_ipc_alloc(hmrt_ctx* ctx, size_t bytes, void** d_ptr, void* handle64);
/* The other way to assemble a sharded render: copy a call's COMPACT output (d_tiles, the layout of hmrt_trace without
 * full_frame_output) to its place in whole frames (d_frames: H rows per frame; may be a peer mapping) with 2-D device
 * copies on the context's stream -- copy engines instead of in-kernel stores (DESIGN.md section 6 compares the two). */
int hmrt_copy_tiles_to_frames(hmrt_ctx* ctx, const uint8_t* d_tiles, uint8_t* d_frames, int W, int H, int n_frames, int tile_first,
                              int tile_stride);
int hmrt_ipc_free(hmrt_ctx* ctx, void* d_ptr);
int hmrt_ipc_open(hmrt_ctx* ctx, const void* handle64, void** d_ptr);
int hmrt_ipc_close(hmrt_ctx* ctx, void* d_ptr);

/* ---- multi-GPU exchange steps --------------------------------------------------------- */

/*
 * For C / C++ hosts that run one context per GPU (one process per GPU, or one process driving several): the two collective
 * steps of the path, on the CALLER'S NCCL communicator (`nccl_comm` is this rank's ncclComm_t passed as void*).  The
 * library has no link-time NCCL dependency: libnccl.so.2 is looked up with dlopen on first use (the copy already loaded in
 * the process, if any), HMRT_E_NCCL if there is none.  Both calls are asynchronous on the context's stream; a process
 * that drives several ranks from one thread wraps the per-rank calls in ncclGroupStart / ncclGroupEnd as usual.
 * (Python hosts use torch.distributed for the same two steps, hmrt/dist.py.)
 *
 * hmrt_broadcast_heightmap: replicate the pyramid (and colour map, if given) of rank `root` -- the heightmap is replicated,
 * the image is sharded by row tiles (hmrt_trace_opts.tile_first / tile_stride).
 * hmrt_allreduce_max_heights: combine per-rank partial rasterisations: MAX over the finest level viewed as int32 (heights
 * are >= +0, so integer order == float order, NaN-free and exact) and, if given, over the colour keys as int64 (the last
 * writer in file order wins on every rank); follow with hmrt_build_mips / hmrt_resolve_colors on every rank.
 */
int hmrt_broadcast_heightmap(hmrt_ctx* ctx, void* nccl_comm, float* d_pyramid, hmrt_color* d_color_map, int coarse_res,
                             int levels, int root);
int hmrt_allreduce_max_heights(hmrt_ctx* ctx, void* nccl_comm, float* d_pyramid, uint64_t* d_color_keys, int coarse_res,
                               int levels);

/*
 * Multi-GPU rasterisation as an owner-computes exchange over PEER MEMORY (NVLink P2P through cudaIpc; csrc/rasterx.cu) -- the
 * faster equivalent of "scatter into a private grid, hmrt_allreduce_max_heights, hmrt_build_mips": points are sharded by
 * contiguous range (main.cpp:193's file order), every rank OWNS a band of rows of the finest level, and what travels is
 * the points' (cell, height) pairs to the owner plus one all-gather of the finished bands that is fused with the mip
 * build.  Bit-identical to one GPU (max is associative, commutative, idempotent).  One hmrt_rx per rank; all ranks pass the
 * same (coarse_res, levels, world, max_points_per_rank) and make the same sequence of calls:
 *
 *     hmrt_rx_create; hmrt_rx_export -> exchange the 64-byte handles (any transport) -> hmrt_rx_connect          (once)
 *     hmrt_rx_begin; hmrt_rx_bin (1..n times); hmrt_rx_barrier; hmrt_rx_apply; hmrt_rx_barrier; hmrt_rx_gather_mips   (per cloud)
 *     hmrt_rx_status -> overflow must be 0 on every rank (else: rerun with a larger max_points_per_rank or use the all-reduce path)
 *
 * Everything is asynchronous on the context's stream except hmrt_rx_status; the barriers are flag exchanges in peer memory
 * with a ~10 s time-out (reported by hmrt_rx_status, never a hang).  Needs res0 % 128 == 0, res0 >= 2048, an even coarse_res, 2 <= levels <= 8
 * (HMRT_E_SHAPE otherwise) and 16-byte aligned records of at most 64 bytes.  world == 1 works without any peer (tests).
 * Colour keys are not carried by this path (use hmrt_scatter_las + hmrt_allreduce_max_heights for coloured clouds).
 */
typedef struct hmrt_rx hmrt_rx;
int hmrt_rx_create(hmrt_ctx* ctx, int coarse_res, int levels, int rank, int world, int64_t max_points_per_rank, hmrt_rx** out);
int hmrt_rx_destroy(hmrt_rx* rx);
size_t hmrt_rx_region_bytes(const hmrt_rx* rx);
int hmrt_rx_export(hmrt_rx* rx, void* handle64);
int hmrt_rx_connect(hmrt_rx* rx, const void* handles /* world x 64 bytes, rank order */);
int hmrt_rx_begin(hmrt_rx* rx);
int hmrt_rx_bin(hmrt_rx* rx, const uint8_t* d_records, int64_t n, int record_len, int point_format, const hmrt_las_transform* xf);
int hmrt_rx_barrier(hmrt_rx* rx);
int hmrt_rx_apply(hmrt_rx* rx);
int hmrt_rx_gather_mips(hmrt_rx* rx, float* d_pyramid);
int hmrt_rx_status(hmrt_rx* rx, uint32_t* overflow, uint32_t* error);
/* rank r owns finest rows [rows[r], rows[r + 1]); rows has world + 1 entries */
int hmrt_rx_bands(const hmrt_rx* rx, int* rows);

/* ---- instrumentation -------------------------------------------------------------------- */

/* 0 (default) = production traversal kernel, bit-identical to the reference's castRay (CudaKernel.cu:121-177);
 * 1 = diagnostic: the operation-by-operation walk that mirrors CudaKernel.cu:121-177 line by line (slower; 0 and 1 must
 *     agree bit for bit);
 * 2 = EXPERIMENTAL, opt-in: the march through the empty air above the terrain is replaced by one closed-form step to the
 *     top-level cell in which the ray comes down to the terrain's maximum height; the descent from there is the exact one.
 *     2.3x the rays per second on high-altitude views, but not bit-identical: the entry point carries one rounding instead
 *     of the ~50 accumulated ones of the reference's walk.  Measured against the reference's own code on the benchmark
 *     poses (tests/test_gpu_fullsize.py, bench.py extra.tolerance_mode): colour within 1/255 on 99.99 % and pixel-exact on
 *     99.9 % of the pixels, hit cell equal on 99.8 % -- BELOW the 99.9 % hit-cell bar BASELINE.json states, which is why
 *     no default path and no reported number uses it.  Grids whose coarse_res is not a power of two fall back to variant 0. */
int hmrt_set_trace_variant(hmrt_ctx* ctx, int variant);

/* hmrt_trace_host schedule.  1 = one launch per group of frames (>= 4 M rays), each followed by its device->host copy on a
 * copy stream, the last single-frame launch cut into four tile ranges.  2 = streamed: ONE persistent launch over all frames
 * whose quarter-frame row segments signal their completion; the copy of a segment starts the moment its last tile is stored
 * (cuStreamWaitValue32 on the copy stream; falls back to 1 when the driver entry point is missing).  0 (default) = streamed
 * for single-frame calls (latency-bound: -6 %), per-group for batches (bound by the device->host link either way).
 * Identical results. */
int hmrt_set_host_variant(hmrt_ctx* ctx, int variant);

/* hmrt_compose_window formulation: 0 (default) = one launch of TMA bulk copies (cp.async.bulk through shared-memory stages,
 * no per-thread data movement) wherever rows and shifts are multiples of 16 bytes, 1 = the per-thread 128-bit gather.
 * Identical results. */
int hmrt_set_window_variant(hmrt_ctx* ctx, int variant);

/* Experiment ("L2 residency"): fetch the pyramid levels >= first_level with a persisting-L2 access
 * policy window (cudaAccessPropertyPersisting for `hit_ratio` of the window, streaming for the rest) on every trace launch;
 * first_level < 1 switches it off and resets the persisting lines.  Measured neutral-to-negative on B200 (126 MB of L2 keep
 * the coarse levels resident by themselves), hence off by default. */
int hmrt_set_l2_persist(hmrt_ctx* ctx, int first_level, float hit_ratio);

/* Counters of the instrumented kernels (calls with d_hits != NULL) since the last reset: out[0] = rays, out[1] = loop
 * iterations (== height fetches of the reference algorithm, CudaKernel.cu:153-176; the S of SURVEY.md section 8(d)),
 * out[2] = the share of them taken in the production walk's air phase (arithmetic only, no height fetch), out[3] = 0.
 * Synchronises the context's stream. */
int hmrt_trace_stats(hmrt_ctx* ctx, uint64_t out[4], int reset);

/* Kernel launches issued by this context since creation (bench.py's gpu_launches). */
int64_t hmrt_launch_count(const hmrt_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* HMRT_H_ */
