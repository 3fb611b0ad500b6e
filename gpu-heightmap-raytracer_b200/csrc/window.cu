/*
 * window.cu -- the camera window over resident sections.
 *
 * Replaces preparePointBuffer + copyPointBuffer (GPUHeightmapRaytracer/src/main.cpp:459-625): the reference assembles the
 * window pyramid (and colour map) on the CPU with four memcpy loops per level out of the 2 x 2 sections the window
 * straddles, then uploads it (139.8 MB at its default size) every frame.  With the sections resident in HBM the window is
 * one gather pass on the device: 2 x 89.5 MB + 2 x 50.3 MB of traffic at the default size, no PCIe.
 */
#include <float.h>
#include <math.h>
#include <string.h>

#include "hmrt_internal.cuh"

namespace hmrt {

struct WindowParams {
  const float* sec[2][2];
  const uint8_t* col[2][2];
  float* out;
  uint8_t* out_col;
  int levels;
  int res[HMRT_MAX_LEVELS];       /* level resolutions, 0 = finest */
  uint32_t idx[HMRT_MAX_LEVELS];  /* level offsets in floats */
  uint32_t cx[HMRT_MAX_LEVELS], cy[HMRT_MAX_LEVELS]; /* cell_position at each level (main.cpp:573-574 doubles it per level) */
  uint32_t quad_end[HMRT_MAX_LEVELS]; /* exclusive prefix ends of the per-level work items (groups of 4 cells), finest first */
  uint32_t vec[HMRT_MAX_LEVELS];      /* level may use 128-bit accesses: res, cx, level offset multiples of 4 cells, buffers 16-byte aligned */
  uint32_t col_vec;                   /* the same for the colour rows */
  uint32_t total_quads;
};

/* One work item = 4 consecutive cells of one row of one level (one cell where the resolution is not a multiple of 4).
 * The source of a window cell is ((x + cx) mod res, (y + cy) mod res) of the section selected by the two carries
 * (the four loops of main.cpp:519-571).  cx is a multiple of 4 on every level but the two coarsest, so the four sources of
 * an item are contiguous, 16-byte aligned and in one section there: one 128-bit load, one 128-bit store. */
__global__ void __launch_bounds__(256) compose_window_kernel(const __grid_constant__ WindowParams p) {
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < p.total_quads; q += gridDim.x * blockDim.x) {
    int l = 0;
    while (q >= p.quad_end[l]) ++l;
    const uint32_t local = q - (l ? p.quad_end[l - 1] : 0u);
    const uint32_t res = (uint32_t)p.res[l];
    const uint32_t per = (res & 3u) == 0 ? 4u : 1u, row_items = res / per;
    const uint32_t y = local / row_items, x = (local - y * row_items) * per;
    const uint32_t sy = y + p.cy[l], wy = sy >= res, yy = wy ? sy - res : sy;
    HMRT_DCHECK(y < res && yy < res && x + per <= res && p.cx[l] < res && p.cy[l] < res);
    float* dst = p.out + p.idx[l] + (size_t)y * res + x;
    const uint32_t sx = x + p.cx[l];
    if (p.vec[l]) {
      const uint32_t wx = sx >= res, xx = wx ? sx - res : sx;
      *reinterpret_cast<float4*>(dst) = __ldcs(reinterpret_cast<const float4*>(p.sec[wx][wy] + p.idx[l] + (size_t)yy * res + xx));
    } else {
      for (uint32_t k = 0; k < per; ++k) {
        const uint32_t s = sx + k, wx = s >= res, xx = wx ? s - res : s;
        dst[k] = __ldcs(p.sec[wx][wy] + p.idx[l] + (size_t)yy * res + xx);
      }
    }
  }
}

/* Colour map (main.cpp:576-618): rows of res0 * 3 bytes, shifted by cx0 * 3 bytes.  One work item = 16 output bytes; when
 * the row length and the shift are multiples of 16 bytes (every power-of-two resolution >= 16 with levels >= 3) the source
 * is one aligned 128-bit load inside one section. */
__global__ void __launch_bounds__(256) compose_colors_kernel(const __grid_constant__ WindowParams p) {
  const uint32_t res = (uint32_t)p.res[0];
  const size_t row_bytes = (size_t)res * 3, shift = (size_t)p.cx[0] * 3;
  const bool vec = p.col_vec != 0;
  const size_t items_per_row = (row_bytes + 15) / 16, total = items_per_row * res;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
    const uint32_t y = (uint32_t)(q / items_per_row);
    const size_t b0 = (q - (size_t)y * items_per_row) * 16;
    const uint32_t sy = y + p.cy[0], wy = sy >= res, yy = wy ? sy - res : sy;
    uint8_t* dst = p.out_col + (size_t)y * row_bytes + b0;
    if (vec) {
      const size_t s = b0 + shift;
      const uint32_t wx = s >= row_bytes;
      const size_t sb = wx ? s - row_bytes : s;
      *reinterpret_cast<uint4*>(dst) = __ldcs(reinterpret_cast<const uint4*>(p.col[wx][wy] + (size_t)yy * row_bytes + sb));
    } else {
      const size_t n = row_bytes - b0 < 16 ? row_bytes - b0 : 16;
      for (size_t k = 0; k < n; ++k) {
        const size_t s = b0 + k + shift;
        const uint32_t wx = s >= row_bytes;
        dst[k] = __ldcs(p.col[wx][wy] + (size_t)yy * row_bytes + (wx ? s - row_bytes : s));
      }
    }
  }
}

/* ---------------------------------------------------------------------------------------------------------------------
 * The same gather as ONE launch of TMA bulk copies.  A window row is, per level, at most two contiguous runs of a section
 * row (the x wrap): a 2-D tile move with no arithmetic at all, so no thread needs to touch the data.  Every CTA is a single
 * warp whose first lane drives a ring of kWinStages x 16 KB shared-memory stages: cp.async.bulk global -> shared
 * (completion on the stage's mbarrier), then cp.async.bulk shared -> global (bulk groups); loads of the next stages are in
 * flight while a stage is written out.  Planes = the pyramid levels + the colour map, one launch for all of them (the two
 * launches of the per-thread formulation above left a gap between them).  Planes whose rows / shifts are not multiples of
 * 16 bytes (the two coarsest levels, odd resolutions) are copied by the warp's lanes directly.
 */
constexpr int kWinStages = 7; /* 7 x 16 KB per CTA, two CTAs per SM: up to 12 loads in flight per SM */
constexpr uint32_t kWinPiece = 16384;
constexpr int kWinMaxPlanes = HMRT_MAX_LEVELS + 1;

struct WinPlane {
  const uint8_t* src[2][2];
  uint8_t* dst;
  uint32_t row_bytes, rows, shift_bytes, cy, pieces_per_row, tma_ok;
  uint32_t job_end; /* exclusive prefix end of this plane's jobs */
};
struct WinTmaParams {
  WinPlane plane[kWinMaxPlanes];
  int n_planes;
  uint32_t total_jobs;
};

__device__ __forceinline__ void win_mbar_init(uint32_t bar) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void win_mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void win_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void win_load(uint32_t dst_s, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_s), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void win_store(void* dst, uint32_t src_s, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_s), "r"(bytes) : "memory");
}

struct WinJob {
  const uint8_t *a, *b; /* up to two source runs */
  uint8_t* dst;
  uint32_t len_a, len_b, tma_ok;
};

__device__ __forceinline__ WinJob win_job(const WinTmaParams& p, uint32_t j) {
  int l = 0;
  while (j >= p.plane[l].job_end) ++l;
  const WinPlane& pl = p.plane[l];
  const uint32_t local = j - (l ? p.plane[l - 1].job_end : 0u);
  const uint32_t y = local / pl.pieces_per_row, piece = local - y * pl.pieces_per_row;
  const uint32_t b0 = piece * kWinPiece;
  const uint32_t len = min(kWinPiece, pl.row_bytes - b0);
  const uint32_t sy = y + pl.cy, wy = sy >= pl.rows, yy = wy ? sy - pl.rows : sy;
  const uint32_t s0 = b0 + pl.shift_bytes;
  WinJob w;
  w.dst = pl.dst + (size_t)y * pl.row_bytes + b0;
  w.tma_ok = pl.tma_ok;
  if (s0 >= pl.row_bytes) { /* wholly in the right-hand section */
    w.a = pl.src[1][wy] + (size_t)yy * pl.row_bytes + (s0 - pl.row_bytes);
    w.len_a = len, w.b = nullptr, w.len_b = 0;
  } else {
    w.a = pl.src[0][wy] + (size_t)yy * pl.row_bytes + s0;
    w.len_a = min(len, pl.row_bytes - s0);
    w.len_b = len - w.len_a;
    w.b = pl.src[1][wy] + (size_t)yy * pl.row_bytes;
  }
  return w;
}

__global__ void __launch_bounds__(32) compose_window_tma_kernel(const __grid_constant__ WinTmaParams p) {
  extern __shared__ __align__(128) uint8_t win_smem[];
  __shared__ __align__(8) unsigned long long bars[kWinStages];
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
  const uint32_t stage0 = (uint32_t)__cvta_generic_to_shared(win_smem);
  const int lane = threadIdx.x;
  if (lane == 0) {
    for (int s = 0; s < kWinStages; ++s) win_mbar_init(bar0 + 8 * s);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const uint32_t n_local = p.total_jobs > blockIdx.x ? (p.total_jobs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
  uint32_t phase = 0; /* bit s = parity of the next completion of stage s's mbarrier (only TMA jobs complete a phase) */
  /* what the write-out of a stage needs, noted when its load is issued (the job arithmetic runs once per job) */
  __shared__ uint8_t* ring_dst[kWinStages];
  __shared__ uint32_t ring_len[kWinStages];
  for (uint32_t k = 0; k < n_local + (kWinStages - 1); ++k) {
    if (k >= (uint32_t)(kWinStages - 1)) { /* write out job kk */
      const uint32_t kk = k - (kWinStages - 1);
      const uint32_t s = kk % kWinStages;
      if (ring_len[s]) {
        if (lane == 0) {
          win_mbar_wait(bar0 + 8 * s, (phase >> s) & 1u);
          phase ^= 1u << s;
          win_store(ring_dst[s], stage0 + s * kWinPiece, ring_len[s]);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else { /* unaligned plane: the lanes copy the piece themselves (small coarse levels, odd resolutions) */
        const WinJob w = win_job(p, blockIdx.x + kk * gridDim.x);
        for (uint32_t i = lane; i < w.len_a; i += 32) w.dst[i] = __ldcs(w.a + i);
        for (uint32_t i = lane; i < w.len_b; i += 32) w.dst[w.len_a + i] = __ldcs(w.b + i);
        if (lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory"); /* keeps the group count in step */
      }
      __syncwarp();
    }
    if (k < n_local) { /* load job k into the stage job k - kWinStages used: its write-out was committed one trip ago */
      const uint32_t s = k % kWinStages;
      if (lane == 0) {
        const WinJob w = win_job(p, blockIdx.x + k * gridDim.x);
        HMRT_DCHECK(w.len_a + w.len_b <= kWinPiece && (!w.tma_ok || ((w.len_a | w.len_b) & 15u) == 0));
        ring_len[s] = w.tma_ok ? w.len_a + w.len_b : 0u;
        ring_dst[s] = w.dst;
        if (w.tma_ok) {
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          win_mbar_expect(bar0 + 8 * s, w.len_a + w.len_b);
          win_load(stage0 + s * kWinPiece, w.a, w.len_a, bar0 + 8 * s);
          if (w.len_b) win_load(stage0 + s * kWinPiece + w.len_a, w.b, w.len_b, bar0 + 8 * s);
        }
      }
      __syncwarp();
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace hmrt

extern "C" {

int hmrt_window_place(const float camera_position[3], const float* section_origins, int grid, int coarse_res, int levels,
                      hmrt_window_placement* out) {
  if (!camera_position || !section_origins || !out || grid < 1) return HMRT_E_ARG;
  int rc = hmrt::pyramid_layout(coarse_res, levels, nullptr, nullptr, nullptr);
  if (rc) return rc;
  auto origin = [&](int i, int j, int a) { return section_origins[((size_t)i * grid + j) * 2 + a]; };
  const float top = ldexpf(1.0f, levels - 1);                  /* glm::pow(2.0f, LOD_levels - 1) */
  const float off = top * (float)coarse_res / 2.0f;            /* main.cpp:465-467 */
  const float blx = camera_position[0] - off, bly = camera_position[2] - off; /* :469 */
  const float trx = (camera_position[0] + off) - FLT_MIN, try_ = (camera_position[2] + off) - FLT_MIN; /* :470 */
  /* :472-502.  The reference tests the origin before the index bound (an out-of-bounds read when the window lies
   * beyond the last section); the bound is tested first here, the result is the same whenever the reference's is defined. */
  int min_x = 0, min_y = 0, max_x = 0, max_y = 0;
  while (min_x < grid && blx > origin(min_x, 0, 0)) min_x++;
  min_x--;
  while (min_y < grid && bly > origin(0, min_y, 1)) min_y++;
  min_y--;
  while (max_x < grid && trx > origin(max_x, 0, 0)) max_x++;
  max_x--;
  while (max_y < grid && try_ > origin(0, max_y, 1)) max_y++;
  max_y--;
  if (min_x < 0 || min_y < 0 || max_x < 0 || max_y < 0) return HMRT_E_ARG;
  const float spx = blx - origin(min_x, min_y, 0), spy = bly - origin(min_x, min_y, 1); /* :509 */
  const int cx = (int)floorf(spx / top), cy = (int)floorf(spy / top);                   /* :510 */
  if (cx < 0 || cy < 0 || cx >= coarse_res || cy >= coarse_res) return HMRT_E_ARG;
  out->min_x = min_x, out->min_y = min_y, out->max_x = max_x, out->max_y = max_y;
  out->cell_x = cx, out->cell_y = cy;
  /* :513-516; glm::pow(2.0f, LOD_levels - 2) is 0.5 for a single level */
  const float half_top = ldexpf(1.0f, levels - 2);
  out->camera[0] = (spx - (float)cx * top) + (float)(coarse_res - 1) * half_top;
  out->camera[1] = camera_position[1];
  out->camera[2] = (spy - (float)cy * top) + (float)(coarse_res - 1) * half_top;
  return 0;
}

int hmrt_compose_window(hmrt_ctx* ctx, const hmrt_window_sections* sections, int coarse_res, int levels, int cell_x, int cell_y,
                        float* d_window_pyramid, hmrt_color* d_window_color_map) {
  if (!ctx || !sections || !d_window_pyramid) return HMRT_E_ARG;
  if (cell_x < 0 || cell_y < 0 || cell_x >= coarse_res || cell_y >= coarse_res) return HMRT_E_ARG;
  hmrt::WindowParams p;
  memset(&p, 0, sizeof(p));
  int64_t idx[HMRT_MAX_LEVELS], total = 0;
  int rc = hmrt::pyramid_layout(coarse_res, levels, p.res, idx, &total);
  if (rc) return rc;
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      if (!sections->d_pyramid[a][b]) return HMRT_E_ARG;
      if (sections->d_pyramid[a][b] == d_window_pyramid) return HMRT_E_ARG; /* not in place */
      if (d_window_color_map && !sections->d_color_map[a][b]) return HMRT_E_ARG;
      p.sec[a][b] = sections->d_pyramid[a][b];
      p.col[a][b] = reinterpret_cast<const uint8_t*>(sections->d_color_map[a][b]);
    }
  p.out = d_window_pyramid;
  p.out_col = reinterpret_cast<uint8_t*>(d_window_color_map);
  p.levels = levels;
  uint64_t quads = 0;
  uintptr_t align = reinterpret_cast<uintptr_t>(d_window_pyramid), col_align = reinterpret_cast<uintptr_t>(d_window_color_map);
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) align |= reinterpret_cast<uintptr_t>(p.sec[a][b]), col_align |= reinterpret_cast<uintptr_t>(p.col[a][b]);
  for (int l = 0; l < levels; ++l) {
    p.idx[l] = (uint32_t)idx[l];
    p.cx[l] = (uint32_t)cell_x << (levels - 1 - l); /* main.cpp:573-574 */
    p.cy[l] = (uint32_t)cell_y << (levels - 1 - l);
    p.vec[l] = (p.res[l] % 4 == 0) && (p.cx[l] % 4 == 0) && (idx[l] % 4 == 0) && (align % 16 == 0);
    quads += (uint64_t)p.res[l] * (uint64_t)p.res[l] / (p.res[l] % 4 == 0 ? 4u : 1u);
    p.quad_end[l] = (uint32_t)quads;
  }
  p.col_vec = ((uint64_t)p.res[0] * 3 % 16 == 0) && ((uint64_t)p.cx[0] * 3 % 16 == 0) && (col_align % 16 == 0);
  if (quads >= (1ull << 32)) return HMRT_E_SHAPE;
  p.total_quads = (uint32_t)quads;
  hmrt::DeviceGuard guard(ctx->device);
  /* one launch of TMA bulk copies when the big planes qualify (rows and shifts multiples of 16 bytes) */
  if (p.vec[0] && (!d_window_color_map || p.col_vec) && ctx->window_variant != 1) {
    hmrt::WinTmaParams t;
    memset(&t, 0, sizeof(t));
    uint64_t jobs = 0;
    int n = 0;
    for (int l = 0; l < levels; ++l, ++n) {
      hmrt::WinPlane& pl = t.plane[n];
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) pl.src[a][b] = reinterpret_cast<const uint8_t*>(p.sec[a][b] + p.idx[l]);
      pl.dst = reinterpret_cast<uint8_t*>(p.out + p.idx[l]);
      pl.row_bytes = (uint32_t)p.res[l] * 4u;
      pl.rows = (uint32_t)p.res[l];
      pl.shift_bytes = p.cx[l] * 4u;
      pl.cy = p.cy[l];
      pl.pieces_per_row = (pl.row_bytes + hmrt::kWinPiece - 1) / hmrt::kWinPiece;
      pl.tma_ok = p.vec[l];
      jobs += (uint64_t)pl.rows * pl.pieces_per_row;
      pl.job_end = (uint32_t)jobs;
    }
    if (d_window_color_map) {
      hmrt::WinPlane& pl = t.plane[n++];
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) pl.src[a][b] = p.col[a][b];
      pl.dst = p.out_col;
      pl.row_bytes = (uint32_t)p.res[0] * 3u;
      pl.rows = (uint32_t)p.res[0];
      pl.shift_bytes = p.cx[0] * 3u;
      pl.cy = p.cy[0];
      pl.pieces_per_row = (pl.row_bytes + hmrt::kWinPiece - 1) / hmrt::kWinPiece;
      pl.tma_ok = 1;
      jobs += (uint64_t)pl.rows * pl.pieces_per_row;
      pl.job_end = (uint32_t)jobs;
    }
    if (jobs < (1ull << 32)) {
      t.n_planes = n;
      t.total_jobs = (uint32_t)jobs;
      const size_t smem = (size_t)hmrt::kWinStages * hmrt::kWinPiece;
      static bool configured = false;
      if (!configured) {
        HMRT_CUDA(cudaFuncSetAttribute(hmrt::compose_window_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
      }
      const unsigned cap_tma = (unsigned)ctx->sm_count * 2u; /* 2 x 112 KB of stages per SM */
      const unsigned grid = jobs < cap_tma ? (unsigned)jobs : cap_tma;
      hmrt::compose_window_tma_kernel<<<grid, 32, smem, ctx->stream>>>(t);
      HMRT_LAUNCHED(ctx);
      return 0;
    }
  }
  const unsigned cap = (unsigned)ctx->sm_count * 8u; /* 8 x 256 threads per SM, grid-stride */
  unsigned blocks = (unsigned)((quads + 255) / 256);
  hmrt::compose_window_kernel<<<blocks < cap ? blocks : cap, 256, 0, ctx->stream>>>(p);
  HMRT_LAUNCHED(ctx);
  if (d_window_color_map) {
    const uint64_t items = ((uint64_t)p.res[0] * 3 + 15) / 16 * (uint64_t)p.res[0];
    blocks = (unsigned)((items + 255) / 256 < cap ? (items + 255) / 256 : cap);
    hmrt::compose_colors_kernel<<<blocks, 256, 0, ctx->stream>>>(p);
    HMRT_LAUNCHED(ctx);
  }
  return 0;
}

}  // extern "C"
