/*
 * rasterx.cu -- tile-binned rasterisation of point clouds WITHOUT spatial order, on one GPU and -- as an owner-computes
 * exchange over peer memory -- on the GPUs of one NVLink box.
 *
 * CPU semantics being replaced: the loadLASToSection inner loop (GPUHeightmapRaytracer/src/main.cpp:193-234); its fixed
 * point is `finest = max over points (floored at +0), level i+1 = max of 2x2 children`, independent of point order
 * (raster.cu).  A direct RED.MAX per point (raster.cu) is the right kernel for survey-ordered files; for uniformly random
 * points every atomic touches a different 32-byte DRAM sector of a 1 GiB grid (BASELINE config 4: 22 G points/s).  Here:
 *
 *   pass 1  rx_bin_kernel     streams the records through shared memory with TMA bulk copies (cp.async.bulk + mbarrier,
 *                             two stages: chunk k+1 is in flight while chunk k is decoded), decodes each point once,
 *                             counting-sorts the chunk's (cell, height bits) pairs by grid tile (2048^2 cells: 8 x 8 tiles at
 *                             16384^2) in shared memory and appends every tile's run to the CTA's private slice of that
 *                             tile's bucket: no global atomics, lanes store to a few contiguous runs.
 *   pass 2  rx_apply_kernel   walks the buckets tile by tile, so the part of the grid under a tile (16 MB) stays L2-resident
 *                             while its RED.MAXes land.
 *
 * Multi-GPU (BASELINE config 4: points sharded by contiguous range).  The north star's exchange is a max all-reduce of the
 * dense finest level (1 GiB per rank at 16384^2, 93 % of it untouched by a rank's own points).  Because pass 1 already
 * sorts by tile, the exchange can move POINTS instead of grids: every rank owns a band of tile rows,
 *
 *   bin (local)  ->  barrier  ->  apply: each rank pulls the pairs of ITS tiles from every rank's buckets over NVLink
 *   (P2P loads from peer memory, cudaIpc) and reduces them into its band  ->  barrier  ->  gather + mip build in ONE
 *   kernel: every 128 x 128 tile is read from its owner's band (P2P), stored into the local finest level and reduced to
 *   all coarser levels on the way.
 *
 * Exchange volume per rank: (N-1)/N of the pairs it owns (8 B per point) + (N-1)/N of the finest level, instead of
 * 2 (N-1)/N of the finest level for the all-reduce, and no separate mip pass.  The barriers are flag exchanges in peer
 * memory (one warp per GPU, system-scope stores / polls with a timeout).  Results are bit-identical to one GPU: max is
 * associative, commutative and idempotent.
 */
#include <new>
#include <stdio.h>

#include "raster_common.cuh"

namespace hmrt {

constexpr int kBinThreads = 256; /* CTA size of the apply pass */
constexpr int kMaxTiles = 256; /* binned_tile_shift() gives at most 16 x 16 tiles (usually 8 x 8) */
constexpr int kMaxPeers = 16;
constexpr size_t kHeaderBytes = 4096;

/* Tile side = 2^shift.  The apply pass wants the grid under a tile to sit in L2 (16 MB = 2048^2 cells measured fine, 64 MB
 * twice as slow), the bin pass wants few tiles (64 instead of 256: 3.12 -> 2.72 ms for 500 M points,
 * profiles/raw_r02/raster_tiles_per_axis.json): ceil(res0 / 2048) tiles per axis, at least 8, at most 16 (kMaxTiles).  The
 * exchange deals out whole tile rows, so a world size that does not divide 8 keeps 16 rows. */
int binned_tile_shift(int res0, int world) {
  int per_axis = (res0 + 2047) / 2048;
  per_axis = per_axis < 8 ? 8 : per_axis > 16 ? 16 : per_axis;
  if (8 % world != 0) per_axis = 16;
  int s = 0;
  while (((res0 + (1 << s) - 1) >> s) > per_axis) ++s;
  return s;
}

/* region header, at offset 0 of every rank's exchange region */
struct RxHeader {
  uint32_t flags[kMaxPeers]; /* flags[r]: last barrier epoch rank r has reached (written by rank r over NVLink) */
  uint32_t overflow;         /* points the bin pass could not store because a slice was full (distributed mode) */
  uint32_t error;            /* 1 = a barrier timed out */
};

template <bool KEYS> struct PairOf { typedef uint2 type; };
template <> struct PairOf<true> { typedef uint4 type; };

struct BinParams {
  ScatterParams sp;
  const uint8_t* records;
  int64_t n;
  int record_len;
  int per;                /* records per thread and step: chunk = kBinCtaThreads * per (a template parameter of the kernel launched) */
  void* pairs;            /* [n_tiles][n_slices][slice_cap] (cell, height bits) pairs, or -- with colour keys -- (cell, height
                           * bits, key lo, key hi) quads */
  uint32_t* counts;       /* [n_tiles][n_slices] */
  uint32_t slice_cap;
  int tile_shift, tiles_x, n_tiles;
  int* finest;            /* single GPU: overflowing points go straight to the grid; NULL: they are counted in *overflow */
  unsigned long long* keys; /* colour keys next to `finest` (single GPU, KEYS instantiation) */
  uint32_t* overflow;
  int accumulate;         /* continue the slices of an earlier call of the same rasterisation */
  uint32_t cta_rot;       /* chunk c goes to CTA (c + cta_rot) % grid: successive small calls fill different slices */
  int64_t first_index;    /* colour keys: file index of record 0 (main.cpp:223-224: the last writer in file order wins) */
};

/* ---- mbarrier / TMA bulk copy (sm_90+) -------------------------------------------------------------------------- */
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
  } while (!ok);
}
/* global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion is counted on `bar` */
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ void store_pair(uint2* dst, uint32_t cell, uint32_t hb, uint32_t, int64_t) { *dst = make_uint2(cell, hb); }
__device__ __forceinline__ void store_pair(uint4* dst, uint32_t cell, uint32_t hb, uint32_t rgb, int64_t file_index) {
  const unsigned long long key = ((unsigned long long)(file_index + 1) << 24) | rgb; /* +1: key 0 means "never written" */
  *dst = make_uint4(cell, hb, (uint32_t)key, (uint32_t)(key >> 32));
}
__device__ __forceinline__ void apply_pair(int* finest, unsigned long long*, const uint2& v) { atomicMax(finest + v.x, (int)v.y); }
__device__ __forceinline__ void apply_pair(int* finest, unsigned long long* keys, const uint4& v) {
  atomicMax(keys + v.x, ((unsigned long long)v.w << 32) | v.z); /* main.cpp:223-224 */
  if (v.y) atomicMax(finest + v.x, (int)v.y);                    /* :227-233 */
}

/*
 * Pass 1.  256 threads x PER records per step (PER = 8 -> steps of 2048 points; 4 or 2 when the records are so long that two
 * CTAs would no longer fit an SM), two stages of TMA bulk copies.  The record layout (ALIGNED: 20 / 28-byte records whose
 * words are aligned; else five words and funnel shifts), the cell-size form (POW2: the reference's 2.0 -> one multiplication;
 * else the IEEE division) and the step size are compile-time facts, so the decodes of a thread are one straight-line block,
 * run as three passes over its records -- loads, arithmetic, ranks -- whose latencies overlap.  The step:
 *   A  decode; rank of every point inside its tile = value returned by a shared-memory atomicAdd on the tile's counter
 *   B  one thread per tile: its warp derives the base of its group of 32 tiles itself (sum of the earlier groups' counts,
 *      butterfly) and scans its own 32 counts -- no single-warp prefix while the others wait; slice bookkeeping
 *   C  the pairs go to their sorted place in shared memory
 *   D  write-out: consecutive lanes hit consecutive addresses inside a tile's run
 * with three CTA barriers (the histogram is double buffered and cleared in B; what D reads is next written behind two
 * barriers of the following step).  Every shared-memory access is an LDS / STS on a compile-time offset of the dynamic window.
 * Shared memory: [2 mbarriers][fill 256][hist 2 x 256][offs, dest0 256 x 2][sdest chunk][spair chunk][stage 0][stage 1], a
 * stage = chunk * record_len bytes (+ 16 so that the 5-word head load of the last record stays inside).
 */
constexpr int kBinCtaThreads = 256;
constexpr uint32_t kFastFill = 16, kFastHist = kFastFill + 4 * kMaxTiles, kFastOffd = kFastHist + 8 * kMaxTiles, kFastSdest = kFastOffd + 8 * kMaxTiles;
template <bool KEYS, int CHUNK> __host__ __device__ constexpr uint32_t fast_stage0() {
  return (kFastSdest + 4u * CHUNK + (uint32_t)sizeof(typename PairOf<KEYS>::type) * CHUNK + 127u) & ~127u;
}

template <bool KEYS, bool ALIGNED, bool POW2, bool FULL, int kFastThreads, int kFastPer>
__device__ __forceinline__ void bin_decode(const BinParams& p, const uint8_t* stage, uint32_t stage_bytes, int count, uint32_t* hist, uint32_t (&cell)[kFastPer],
                                            uint32_t (&hb)[kFastPer], uint32_t (&slot)[kFastPer], uint32_t (&rgb)[kFastPer]) {
  const float r0f = (float)p.sp.res0;
  const double bias = 4503601774854144.0; /* 2^52 + 2^31 */
  /* three passes over the thread's four records -- loads, arithmetic, ranks -- so that their latencies overlap */
  int32_t X[kFastPer], Y[kFastPer], Z[kFastPer];
  uint32_t tail[kFastPer];
#pragma unroll
  for (int q = 0; q < kFastPer; ++q) {
    const int i = q * kFastThreads + (int)threadIdx.x;
    X[q] = Y[q] = Z[q] = 0;
    tail[q] = 7u << 24; /* past the end of the input: rejected like a class-7 point */
    if (FULL || i < count) {
      const uint32_t off = (uint32_t)i * (uint32_t)p.record_len;
      HMRT_DCHECK(off + 20u <= stage_bytes);
      if (ALIGNED) { /* 20 / 28-byte records: X, Y, Z and the flags word are aligned words */
        const uint32_t* w = reinterpret_cast<const uint32_t*>(stage + off);
        X[q] = (int32_t)w[0], Y[q] = (int32_t)w[1], Z[q] = (int32_t)w[2], tail[q] = w[3];
      } else { /* 26 / 34-byte records alternate between 0 and 2 mod 4: five words and funnel shifts */
        const uint32_t* w = reinterpret_cast<const uint32_t*>(stage + (off & ~3u));
        const uint32_t sh = (off & 3u) * 8u;
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
        X[q] = (int32_t)__funnelshift_r(w0, w1, sh), Y[q] = (int32_t)__funnelshift_r(w1, w2, sh), Z[q] = (int32_t)__funnelshift_r(w2, w3, sh);
        tail[q] = __funnelshift_r(w3, w4, sh);
      }
    }
  }
  uint32_t tile[kFastPer];
  bool ok[kFastPer];
#pragma unroll
  for (int q = 0; q < kFastPer; ++q) {
    /* libLAS 1.8.0 Point::GetX(): raw * scale + offset, two roundings in double, then main.cpp:200-209 with the conversions
     * taken off the XU pipe: (double)int32 by the 2^52 + 2^31 bias trick (exact), and -- since floor(t) >= 0 <=> t >= 0 and
     * floor(t) < res0 <=> t < res0 for an integer res0 -- the range test on t itself */
    const double rx_ = __dsub_rn(__hiloint2double(0x43300000, (int)((uint32_t)X[q] ^ 0x80000000u)), bias);
    const double ry_ = __dsub_rn(__hiloint2double(0x43300000, (int)((uint32_t)Y[q] ^ 0x80000000u)), bias);
    const double rz_ = __dsub_rn(__hiloint2double(0x43300000, (int)((uint32_t)Z[q] ^ 0x80000000u)), bias);
    const double gx = __dadd_rn(__dmul_rn(rx_, p.sp.scale[0]), p.sp.offset[0]);
    const double gy = __dadd_rn(__dmul_rn(ry_, p.sp.scale[1]), p.sp.offset[1]);
    const double gz = __dadd_rn(__dmul_rn(rz_, p.sp.scale[2]), p.sp.offset[2]);
    const float vx = __double2float_rn(__dsub_rn(gx, p.sp.mn[0])), vy = __double2float_rn(__dsub_rn(gy, p.sp.mn[1]));
    const float vz = __double2float_rn(__dsub_rn(gz, p.sp.mn[2]));
    const float fX = POW2 ? __fmul_rn(vx, p.sp.rcell[0]) : div_cell(vx, p.sp.cell[0], p.sp.rcell[0]); /* :200, x / 2^k == x * 2^-k */
    const float fY = POW2 ? __fmul_rn(vy, p.sp.rcell[1]) : div_cell(vy, p.sp.cell[1], p.sp.rcell[1]); /* :201 */
    const float fZ = POW2 ? __fmul_rn(vz, p.sp.rcell[2]) : div_cell(vz, p.sp.cell[2], p.sp.rcell[2]); /* :202 */
    const float tx = __fsub_rn(fX, p.sp.origin[0]), ty = __fsub_rn(fY, p.sp.origin[1]);      /* :205-206 */
    const bool inside = tx >= 0.0f && tx < r0f && ty >= 0.0f && ty < r0f;                    /* :209 */
    /* the colour is written before the height test (main.cpp:223-224 precede :229): with keys, points below the floor still
     * travel, carrying +0.0f (a no-op for the height) */
    ok[q] = inside && ((tail[q] >> 24) & 0x1fu) != 7u && (KEYS || fZ >= 0.0f);
    /* floor(t) from the significand of t + 2^23 rounded down; garbage (never used) when the point is rejected */
    const uint32_t cx = __float_as_uint(__fadd_rd(tx, 8388608.0f)) & 0x7fffffu;
    const uint32_t cy = __float_as_uint(__fadd_rd(ty, 8388608.0f)) & 0x7fffffu;
    cell[q] = cx + cy * (uint32_t)p.sp.res0;
    hb[q] = fZ >= 0.0f ? __float_as_uint(fZ) : 0u;
    tile[q] = ok[q] ? (cy >> p.tile_shift) * (uint32_t)p.tiles_x + (cx >> p.tile_shift) : 0u;
    if (KEYS) {
      rgb[q] = 0;
      if (ok[q] && p.sp.rgb_off >= 0) { /* three u16, 2-byte aligned in every LAS 1.2 format */
        const uint32_t off = (uint32_t)(q * kFastThreads + (int)threadIdx.x) * (uint32_t)p.record_len;
        HMRT_DCHECK(off + (uint32_t)p.sp.rgb_off + 6u <= stage_bytes);
        const uint16_t* c16 = reinterpret_cast<const uint16_t*>(stage + off + (uint32_t)p.sp.rgb_off);
        rgb[q] = (color16(c16[0]) << 16) | (color16(c16[1]) << 8) | color16(c16[2]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < kFastPer; ++q) {
    uint32_t rank = 0;
    HMRT_DCHECK(!ok[q] || (tile[q] < (uint32_t)p.n_tiles && cell[q] < (uint32_t)p.sp.res0 * (uint32_t)p.sp.res0));
    if (ok[q]) rank = atomicAdd(&hist[tile[q]], 1u); /* rank < 2048 */
    slot[q] = ok[q] ? (tile[q] << 16) | rank : 0xffffffffu;
  }
}

template <bool KEYS, bool ALIGNED, bool POW2, int kFastPer>
__global__ void __launch_bounds__(kBinCtaThreads, 2) rx_bin_kernel(const __grid_constant__ BinParams p) {
  typedef typename PairOf<KEYS>::type Pair;
  constexpr int kFastThreads = kBinCtaThreads;
  constexpr int kFastChunk = kFastThreads * kFastPer;
  constexpr uint32_t kFastSpair = kFastSdest + 4 * kFastChunk;
  static_assert(kFastThreads >= kMaxTiles, "phase B: one thread per tile");
  extern __shared__ __align__(128) uint8_t rx_smem[];
  uint32_t* fill = reinterpret_cast<uint32_t*>(rx_smem + kFastFill);  /* entries used in this CTA's slice of each bucket (persistent) */
  uint32_t* hist2 = reinterpret_cast<uint32_t*>(rx_smem + kFastHist); /* points of this step per tile, buffer k & 1 */
  uint2* offd = reinterpret_cast<uint2*>(rx_smem + kFastOffd);        /* x: first sorted slot of the tile, y: first destination or ~0u */
  uint32_t* sdest = reinterpret_cast<uint32_t*>(rx_smem + kFastSdest);
  Pair* spair = reinterpret_cast<Pair*>(rx_smem + kFastSpair);
  const uint32_t stage_bytes = ((uint32_t)kFastChunk * (uint32_t)p.record_len + 16u + 127u) & ~127u;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(rx_smem);
  const uint32_t stage0_s = bar0 + fast_stage0<KEYS, kFastChunk>();
  __shared__ uint32_t total_s;

  const uint32_t n_slices = gridDim.x;
  if (threadIdx.x < kMaxTiles) {
    const int t = threadIdx.x;
    fill[t] = (p.accumulate && t < p.n_tiles) ? p.counts[(size_t)t * n_slices + blockIdx.x] : 0u;
    hist2[t] = 0;
    hist2[kMaxTiles + t] = 0;
  }
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int64_t n_chunks = (p.n + kFastChunk - 1) / kFastChunk;
  const int tiles_padded = (p.n_tiles + 31) & ~31; /* whole warps of phase B */
  /* issue the load of chunk c into stage s (thread 0 issues the bulk copy, the < 16 tail bytes of the very last chunk are
   * copied by the first lanes; every use is separated from this point by at least one __syncthreads) */
  auto issue = [&](int64_t c, int s) {
    const int64_t first = c * kFastChunk;
    const int64_t remaining = p.n - first;
    const uint32_t count = remaining < kFastChunk ? (uint32_t)remaining : (uint32_t)kFastChunk;
    const uint32_t bytes = count * (uint32_t)p.record_len, bulk = bytes & ~15u;
    const uint8_t* src = p.records + first * p.record_len;
    if (threadIdx.x == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (bulk) {
        mbar_expect_tx(bar0 + 8 * s, bulk);
        tma_load_1d(stage0_s + (uint32_t)s * stage_bytes, src, bulk, bar0 + 8 * s);
      } else {
        mbar_arrive(bar0 + 8 * s);
      }
    }
    if (threadIdx.x < bytes - bulk) rx_smem[fast_stage0<KEYS, kFastChunk>() + (size_t)s * stage_bytes + bulk + threadIdx.x] = __ldg(src + bulk + threadIdx.x);
  };

  int64_t c = (blockIdx.x + gridDim.x - p.cta_rot % gridDim.x) % gridDim.x; /* slices stay indexed by blockIdx.x */
  if (c < n_chunks) issue(c, 0);
  __syncthreads();
  for (int k = 0; c < n_chunks; ++k, c += gridDim.x) {
    const int s = k & 1;
    if (c + gridDim.x < n_chunks) issue(c + gridDim.x, s ^ 1);
    mbar_wait(bar0 + 8 * s, (uint32_t)(k >> 1) & 1u);
    const uint8_t* stage = rx_smem + fast_stage0<KEYS, kFastChunk>() + (size_t)s * stage_bytes;
    const int64_t remaining = p.n - c * kFastChunk;
    const int count = remaining < kFastChunk ? (int)remaining : kFastChunk;
    uint32_t* hist = hist2 + s * kMaxTiles;

    /* A: decode, rank inside the tile */
    uint32_t cell[kFastPer], hb[kFastPer], slot[kFastPer], rgb[kFastPer];
    if (count == kFastChunk)
      bin_decode<KEYS, ALIGNED, POW2, true, kFastThreads, kFastPer>(p, stage, stage_bytes, count, hist, cell, hb, slot, rgb);
    else
      bin_decode<KEYS, ALIGNED, POW2, false, kFastThreads, kFastPer>(p, stage, stage_bytes, count, hist, cell, hb, slot, rgb);
    __syncthreads();
    /* B: thread t of the first eight warps owns tile t.  Its warp sums, lane by lane, the counts of the tiles of all earlier
     * groups of 32 (a butterfly turns that into the group's base) and scans its own 32 counts: no second barrier, no
     * single-warp prefix while fifteen warps wait. */
    if ((int)threadIdx.x < tiles_padded) {
      const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5, t = threadIdx.x;
      const uint32_t mine = hist[t];
      uint32_t before = 0;
#pragma unroll
      for (int g = 0; g < kMaxTiles / 32 - 1; ++g)
        if ((uint32_t)g < w) before += hist[g * 32 + lane];
      uint32_t incl = mine;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if ((int)lane >= d) incl += v;
        before += __shfl_xor_sync(0xffffffffu, before, d);
      }
      uint32_t d0 = 0xffffffffu; /* slice full: this step's points of the tile take the overflow route */
      if ((int)t < p.n_tiles) {
        const uint32_t f = fill[t];
        if (f + mine <= p.slice_cap) {
          d0 = (uint32_t)(((size_t)t * n_slices + blockIdx.x) * p.slice_cap + f); /* < 2^32: checked by the launcher */
          fill[t] = f + mine;
        }
      }
      offd[t] = make_uint2(before + incl - mine, d0);
      hist2[(s ^ 1) * kMaxTiles + t] = 0; /* the other buffer: last read in the previous step, next used in the next */
      if ((int)t == tiles_padded - 1) total_s = before + incl;
    }
    __syncthreads();
    /* C: scatter into sorted order (shared memory); the four table loads first, then the stores */
    uint2 od[kFastPer];
#pragma unroll
    for (int q = 0; q < kFastPer; ++q) od[q] = offd[slot[q] == 0xffffffffu ? 0u : slot[q] >> 16];
#pragma unroll
    for (int q = 0; q < kFastPer; ++q) {
      if (slot[q] == 0xffffffffu) continue;
      const uint32_t rank = slot[q] & 0xffffu;
      const uint32_t j = od[q].x + rank;
      HMRT_DCHECK(j < (uint32_t)kFastChunk && rank < (uint32_t)kFastChunk);
      store_pair(spair + j, cell[q], hb[q], KEYS ? rgb[q] : 0u, p.first_index + c * kFastChunk + (int64_t)(q * kFastThreads + (int)threadIdx.x));
      sdest[j] = od[q].y == 0xffffffffu ? od[q].y : od[q].y + rank;
    }
    __syncthreads();
    /* D: write out; consecutive lanes hit consecutive addresses inside a tile's run */
    const uint32_t total = total_s;
    uint32_t dropped = 0;
    for (uint32_t j = threadIdx.x; j < total; j += kFastThreads) {
      const uint32_t d = sdest[j];
      const Pair v = spair[j];
      HMRT_DCHECK(total <= (uint32_t)kFastChunk);
      HMRT_DCHECK(d == 0xffffffffu || (unsigned long long)d < (unsigned long long)p.n_tiles * n_slices * p.slice_cap);
      HMRT_DCHECK(v.x < (uint32_t)p.sp.res0 * (uint32_t)p.sp.res0);
      if (d != 0xffffffffu)
        static_cast<Pair*>(p.pairs)[d] = v;
      else if (p.finest)
        apply_pair(p.finest, p.keys, v);
      else
        ++dropped;
    }
    if (dropped) atomicAdd(p.overflow, dropped);
    /* no barrier here: what D reads (sdest, spair, total_s) is next written behind the two barriers of the following step,
     * and the stage the next iteration refills was last read in A, two barriers ago */
  }
  for (int t = threadIdx.x; t < p.n_tiles; t += kFastThreads) p.counts[(size_t)t * n_slices + blockIdx.x] = fill[t];
}

/* Pass 2.  The buckets of the owned tiles, read from every rank's region (own entry = local memory, the others = peer
 * memory over NVLink), reduced into dst[cell - cell_base].  Every thread keeps four 16-byte loads (eight pairs) in flight
 * before it issues their RED.MAXes.  Slices start 32-byte aligned (slice_cap % 4 == 0). */
struct ApplyParams {
  const uint8_t* peer[kMaxPeers];
  size_t counts_off, pairs_off;
  uint32_t slice_cap, n_slices, world, rank;
  uint32_t tile_first;      /* owned tiles are [tile_first, tile_first + gridDim.x / groups_per_tile) */
  uint32_t groups_per_tile; /* CTAs per tile: ceil(world * n_slices * pieces_per_slice / warps per CTA) */
  uint32_t pieces_per_slice; /* ceil(slice_cap / kApplyPiece) */
  int* dst;
  uint32_t cell_base;
  uint32_t dst_cells;       /* cells behind dst (bounds checks of the checked build) */
  unsigned long long* keys; /* KEYS instantiation (single GPU): colour keys, indexed like dst */
};

/* A WARP's unit of work is one piece of kApplyPiece pairs of one slice (4 trips of 32 lanes x 16 bytes x 2 pairs): slices
 * shrink with the number of ranks (825 pairs each at 8 GPUs), and a whole CTA per slice left most lanes idle there.  The
 * pieces of a tile are numbered slice-major, so the warps of a CTA read neighbouring memory; many CTAs per tile = few
 * tiles in flight = the grid under them stays in L2. */
constexpr uint32_t kApplyPiece = 1024;

template <bool KEYS>
__global__ void __launch_bounds__(kBinThreads) rx_apply_kernel(const __grid_constant__ ApplyParams p) {
  const uint32_t tile = p.tile_first + blockIdx.x / p.groups_per_tile, g = blockIdx.x % p.groups_per_tile;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t all = p.world * p.n_slices;
  const uint32_t item = g * (kBinThreads / 32) + warp; /* piece number inside the tile */
  const uint32_t w = item / p.pieces_per_slice, piece = item - w * p.pieces_per_slice;
  if (w >= all) return;
  int* __restrict__ dst = p.dst - p.cell_base;
  /* ring order over the sources: at any moment the ranks pull from different peers */
  const uint32_t src_rank = (w / p.n_slices + p.rank) % p.world, s = w % p.n_slices;
  const uint8_t* base = p.peer[src_rank];
  const uint32_t count = __ldcv(reinterpret_cast<const uint32_t*>(base + p.counts_off) + (size_t)tile * p.n_slices + s);
  const uint32_t lo = piece * kApplyPiece;
  if (lo >= count) return;
  const uint32_t hi = min(count, lo + kApplyPiece);
  HMRT_DCHECK(count <= p.slice_cap && tile < 256u);
  if (KEYS) { /* one 16-byte quad per point */
    const uint4* src4 = reinterpret_cast<const uint4*>(base + p.pairs_off) + ((size_t)tile * p.n_slices + s) * p.slice_cap;
    for (uint32_t i = lo + lane; i < hi; i += 32 * 4) {
      uint4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t j = i + (uint32_t)k * 32;
        v[k] = j < hi ? __ldcs(src4 + j) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (i + (uint32_t)k * 32 < hi) {
          HMRT_DCHECK(v[k].x - p.cell_base < p.dst_cells);
          apply_pair(dst, p.keys, v[k]);
        }
    }
    return;
  }
  const uint2* src = reinterpret_cast<const uint2*>(base + p.pairs_off) + ((size_t)tile * p.n_slices + s) * p.slice_cap;
  const uint4* src4 = reinterpret_cast<const uint4*>(src); /* lo is even: pieces start 16-byte aligned */
  const uint32_t q_lo = lo >> 1, q_hi = hi >> 1; /* whole 16-byte pieces of two pairs */
  for (uint32_t i = q_lo + lane; i < q_hi; i += 32 * 4) {
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t j = i + (uint32_t)k * 32;
      v[k] = j < q_hi ? __ldcs(src4 + j) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i + (uint32_t)k * 32 < q_hi) {
        HMRT_DCHECK(v[k].x - p.cell_base < p.dst_cells && v[k].z - p.cell_base < p.dst_cells);
        atomicMax(dst + v[k].x, (int)v[k].y);
        atomicMax(dst + v[k].z, (int)v[k].w);
      }
    }
  }
  if ((hi & 1u) && lane == 0) { /* only the last piece of a slice can end on an odd count */
    const uint2 v = __ldcs(src + (hi - 1));
    HMRT_DCHECK(v.x - p.cell_base < p.dst_cells);
    atomicMax(dst + v.x, (int)v.y);
  }
}

/* All-gather of the finest level fused with the mip build: tile (x, z) comes from its owner's band. */
struct GatherParams {
  MipParams mp;
  const float* band[kMaxPeers];  /* first cell of every rank's band (peer memory) */
  int band_row0[kMaxPeers + 1];  /* rank r owns finest rows [band_row0[r], band_row0[r + 1]) */
  int world;
  int first_tile_row;            /* rotation of the tile rows: every rank starts with its own band, then its successor's (ring) */
};

__global__ void __launch_bounds__(512) rx_gather_mips_kernel(const __grid_constant__ GatherParams g) {
  const int tile_z = (int)((blockIdx.y + (unsigned)g.first_tile_row) % gridDim.y);
  const int z = tile_z * 128;
  int o = 0;
  while (o + 1 < g.world && z >= g.band_row0[o + 1]) ++o;
  const float* src = g.band[o] + (size_t)(z - g.band_row0[o]) * g.mp.res0 + (size_t)blockIdx.x * 128;
  mips_tile(g.mp, src, (size_t)g.mp.res0, true, blockIdx.x, tile_z);
}

/* Cross-GPU barrier: rank -> every peer "I have reached `epoch`", then wait until every peer has told me the same.
 * One warp; lane r talks to rank r.  A bounded wait: a rank that never arrives sets header.error instead of hanging the GPU. */
struct BarrierParams {
  uint8_t* peer[kMaxPeers];
  int rank, world;
  uint32_t epoch;
  long long timeout_cycles;
};

__global__ void __launch_bounds__(32) rx_barrier_kernel(const __grid_constant__ BarrierParams p) {
  const int lane = threadIdx.x;
  if (lane < p.world) {
    __threadfence_system();
    volatile uint32_t* remote = &reinterpret_cast<RxHeader*>(p.peer[lane])->flags[p.rank];
    *remote = p.epoch;
    volatile uint32_t* mine = &reinterpret_cast<RxHeader*>(p.peer[p.rank])->flags[lane];
    const long long t0 = clock64();
    while ((int32_t)(*mine - p.epoch) < 0) {
      if (clock64() - t0 > p.timeout_cycles) {
        reinterpret_cast<RxHeader*>(p.peer[p.rank])->error = 1u;
        break;
      }
      __nanosleep(200);
    }
    __threadfence_system();
  }
}

/* ---- host side ---------------------------------------------------------------------------------------------------- */
struct BinGeometry {
  int threads, per, chunk;
  size_t smem;
  int ctas_per_sm;
  const void* fn;
};

/* development knobs (benchmarks/raster_probe.py); 0 = the built-in choice */
static int g_knob_bin_per = 0;

template <int PER>
static const void* bin_kernel_of(bool keys, bool aligned, bool pow2) {
#define HMRT_BIN_K(K, A, P) reinterpret_cast<const void*>(&rx_bin_kernel<K, A, P, PER>)
  return keys ? (aligned ? (pow2 ? HMRT_BIN_K(true, true, true) : HMRT_BIN_K(true, true, false))
                         : (pow2 ? HMRT_BIN_K(true, false, true) : HMRT_BIN_K(true, false, false)))
              : (aligned ? (pow2 ? HMRT_BIN_K(false, true, true) : HMRT_BIN_K(false, true, false))
                         : (pow2 ? HMRT_BIN_K(false, false, true) : HMRT_BIN_K(false, false, false)));
#undef HMRT_BIN_K
}

static int bin_geometry(int record_len, bool keys, const ScatterParams& sp, BinGeometry& g) {
  /* 256 threads x 8 records = steps of 2048 points, two CTAs per SM: measured best of 512 x 4, 512 x 2, 256 x 4, 256 x 8 (2.69 /
   * 3.25 / 2.98 / 2.58 ms for 500 M points, profiles/raw_r02/raster_bin_step_geometry.json).  Longer records take the largest
   * step with which two CTAs still fit an SM (record_len <= 64: 256 x 2 always does). */
  const bool aligned = (record_len & 3) == 0;
  const bool pow2 = sp.rcell[0] != 0.0f && sp.rcell[1] != 0.0f && sp.rcell[2] != 0.0f;
  g.threads = kBinCtaThreads;
  for (int per = g_knob_bin_per ? g_knob_bin_per : 8; per >= 2; per >>= 1) {
    const int chunk = kBinCtaThreads * per;
    const size_t stage = ((size_t)chunk * record_len + 16 + 127) & ~(size_t)127;
    g.per = per, g.chunk = chunk;
    g.fn = per == 8 ? bin_kernel_of<8>(keys, aligned, pow2) : per == 4 ? bin_kernel_of<4>(keys, aligned, pow2) : bin_kernel_of<2>(keys, aligned, pow2);
    g.smem = (per == 8 ? (keys ? fast_stage0<true, 2048>() : fast_stage0<false, 2048>())
              : per == 4 ? (keys ? fast_stage0<true, 1024>() : fast_stage0<false, 1024>())
                         : (keys ? fast_stage0<true, 512>() : fast_stage0<false, 512>())) + 2 * stage;
    if (g.smem > 220 * 1024) continue;
    HMRT_CUDA(cudaFuncSetAttribute(g.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    HMRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g.ctas_per_sm, g.fn, g.threads, g.smem));
    if (g.ctas_per_sm >= 2 || (per == 2 && g.ctas_per_sm >= 1)) return 0;
  }
  return HMRT_E_ARG;
}

static int launch_bin(const BinGeometry& g, const BinParams& bp, unsigned grid, cudaStream_t stream) {
  void* args[] = {const_cast<BinParams*>(&bp)};
  return (int)cudaLaunchKernel(g.fn, dim3(grid), dim3((unsigned)g.threads), args, g.smem, stream);
}

static int64_t slice_capacity(int64_t points, int n_tiles, int n_slices) {
  int64_t cap = 2 * ((points + (int64_t)n_tiles * n_slices - 1) / ((int64_t)n_tiles * n_slices));
  const int64_t floor_cap = 16384 / n_tiles < 64 ? 64 : 16384 / n_tiles; /* small inputs: a step can put 2048 / n_tiles points and more into one slice */
  if (cap < floor_cap) cap = floor_cap;
  return (cap + 3) / 4 * 4; /* keep slices 32-byte aligned */
}

int scatter_binned_single(hmrt_ctx* ctx, const uint8_t* d_records, int64_t n, int record_len, const ScatterParams& sp, int* finest,
                          unsigned long long* keys, int64_t first_index) {
  BinGeometry bg;
  int rc = bin_geometry(record_len, keys != nullptr, sp, bg);
  const size_t pair_size = keys ? sizeof(uint4) : sizeof(uint2);
  if (rc) return rc;
  const int tile_shift = binned_tile_shift(sp.res0, 1);
  const int tiles_x = (sp.res0 + (1 << tile_shift) - 1) >> tile_shift;
  const int n_tiles = tiles_x * tiles_x;
  /* sub-batches bound the workspace (~16 B per point at 2x mean slice capacity) to ~16 GB */
  const int64_t max_batch = (int64_t)1 << 30;
  for (int64_t first = 0; first < n; first += max_batch) {
    const int64_t nb = n - first < max_batch ? n - first : max_batch;
    const int64_t chunks = (nb + bg.chunk - 1) / bg.chunk;
    const int64_t resident = (int64_t)ctx->sm_count * bg.ctas_per_sm;
    const int n_slices = (int)(chunks < resident ? chunks : resident);
    const int64_t slice_cap = slice_capacity(nb, n_tiles, n_slices);
    if ((unsigned long long)n_tiles * (unsigned long long)n_slices * (unsigned long long)slice_cap >= (1ull << 32)) return HMRT_E_SHAPE;
    const size_t counts_bytes = ((size_t)n_tiles * (size_t)n_slices * sizeof(uint32_t) + 255) & ~(size_t)255;
    const size_t need = counts_bytes + (size_t)n_tiles * (size_t)n_slices * (size_t)slice_cap * pair_size;
    if (ctx->ws_cap < need) {
      if (ctx->d_ws) HMRT_CUDA(cudaFree(ctx->d_ws));
      ctx->d_ws = nullptr;
      ctx->ws_cap = 0;
      HMRT_CUDA(cudaMalloc(&ctx->d_ws, need));
      ctx->ws_cap = need;
    }
    uint8_t* ws = static_cast<uint8_t*>(ctx->d_ws);
    BinParams bp;
    bp.sp = sp;
    bp.records = d_records + first * record_len;
    bp.n = nb;
    bp.record_len = record_len;
    bp.per = bg.per;
    bp.counts = reinterpret_cast<uint32_t*>(ws);
    bp.pairs = ws + counts_bytes;
    bp.keys = keys;
    bp.first_index = first_index + first;
    bp.slice_cap = (uint32_t)slice_cap;
    bp.tile_shift = tile_shift;
    bp.tiles_x = tiles_x;
    bp.n_tiles = n_tiles;
    bp.finest = finest;
    bp.overflow = nullptr;
    bp.accumulate = 0;
    bp.cta_rot = 0;
    HMRT_CUDA((cudaError_t)launch_bin(bg, bp, (unsigned)n_slices, ctx->stream));
    HMRT_LAUNCHED(ctx);
    ApplyParams ap;
    memset(&ap, 0, sizeof(ap));
    ap.peer[0] = ws;
    ap.counts_off = 0;
    ap.pairs_off = counts_bytes;
    ap.slice_cap = (uint32_t)slice_cap;
    ap.n_slices = (uint32_t)n_slices;
    ap.world = 1;
    ap.rank = 0;
    ap.tile_first = 0;
    ap.pieces_per_slice = ((uint32_t)slice_cap + kApplyPiece - 1) / kApplyPiece;
    ap.groups_per_tile = ((uint32_t)n_slices * ap.pieces_per_slice + (kBinThreads / 32) - 1) / (kBinThreads / 32);
    ap.dst = finest;
    ap.cell_base = 0;
    ap.dst_cells = (uint32_t)sp.res0 * (uint32_t)sp.res0;
    ap.keys = keys;
    if (keys)
      rx_apply_kernel<true><<<(unsigned)n_tiles * ap.groups_per_tile, kBinThreads, 0, ctx->stream>>>(ap);
    else
      rx_apply_kernel<false><<<(unsigned)n_tiles * ap.groups_per_tile, kBinThreads, 0, ctx->stream>>>(ap);
    HMRT_LAUNCHED(ctx);
  }
  return 0;
}

}  // namespace hmrt

/* ================================================== distributed form ================================================ */

struct hmrt_rx {
  hmrt_ctx* ctx;
  int coarse_res, levels, res0;
  int64_t idx[HMRT_MAX_LEVELS];
  int rank, world;
  int tile_shift, tile, tiles_x, n_tiles;
  int n_slices;
  uint32_t slice_cap;
  size_t counts_off, pairs_off, band_off, region_bytes;
  uint8_t* region;
  uint8_t* peer[hmrt::kMaxPeers];
  bool connected;
  bool binned_any;
  uint32_t chunk_cursor; /* chunks binned since hmrt_rx_begin, modulo n_slices */
  uint32_t epoch;
  int band_row0[hmrt::kMaxPeers + 1];
  int tile_row0[hmrt::kMaxPeers + 1];
};

extern "C" {

int hmrt_rx_barrier(hmrt_rx* rx);

/* Development knob of the binned rasterisation (not part of include/hmrt.h): key 1 = records per thread and step of the bin
 * pass (8, 4 or 2; 0 restores the built-in choice). */
int hmrt_debug_raster_knob(int key, int value) {
  if (key == 1 && (value == 0 || value == 2 || value == 4 || value == 8)) hmrt::g_knob_bin_per = value;
  else return HMRT_E_ARG;
  return 0;
}

int hmrt_rx_create(hmrt_ctx* ctx, int coarse_res, int levels, int rank, int world, int64_t max_points_per_rank, hmrt_rx** out) {
  if (!ctx || !out || world < 1 || world > hmrt::kMaxPeers || rank < 0 || rank >= world || max_points_per_rank < 0) return HMRT_E_ARG;
  *out = nullptr;
  int res[HMRT_MAX_LEVELS];
  int64_t idx[HMRT_MAX_LEVELS];
  int rc = hmrt::pyramid_layout(coarse_res, levels, res, idx, nullptr);
  if (rc) return rc;
  const int shift = hmrt::binned_tile_shift(res[0], world);
  /* bands are whole tile rows and must be whole 128-row mip tiles; the fused mip kernel covers 8 levels */
  if (res[0] % 128 != 0 || shift < 7 || levels > 8 || levels < 2) return HMRT_E_SHAPE;
  /* the fused gather + mip kernel moves 16-byte pieces of level 0 and 8-byte pieces of level 1: odd coarse resolutions put
   * those levels at odd float offsets */
  if (idx[0] % 4 != 0 || idx[1] % 2 != 0) return HMRT_E_SHAPE;
  hmrt::DeviceGuard guard(ctx->device);
  hmrt_rx* rx = new (std::nothrow) hmrt_rx();
  if (!rx) return HMRT_E_NOMEM;
  memset(rx, 0, sizeof(*rx));
  rx->ctx = ctx;
  rx->coarse_res = coarse_res, rx->levels = levels, rx->res0 = res[0];
  for (int i = 0; i < levels; ++i) rx->idx[i] = idx[i];
  rx->rank = rank, rx->world = world;
  rx->tile_shift = shift, rx->tile = 1 << shift;
  rx->tiles_x = (res[0] + rx->tile - 1) >> shift;
  rx->n_tiles = rx->tiles_x * rx->tiles_x;
  /* the same geometry on every rank: all ranks pass the same (coarse_res, levels, world, max_points_per_rank) */
  rx->n_slices = ctx->sm_count * 2;
  rx->slice_cap = (uint32_t)hmrt::slice_capacity(max_points_per_rank, rx->n_tiles, rx->n_slices);
  if ((unsigned long long)rx->n_tiles * (unsigned long long)rx->n_slices * (unsigned long long)rx->slice_cap >= (1ull << 32)) {
    delete rx;
    return HMRT_E_SHAPE;
  }
  /* tile rows are dealt out as evenly as possible, in rank order (contiguous bands) */
  const int base = rx->tiles_x / world, rem = rx->tiles_x % world;
  rx->tile_row0[0] = 0;
  for (int r = 0; r < world; ++r) rx->tile_row0[r + 1] = rx->tile_row0[r] + base + (r < rem ? 1 : 0);
  for (int r = 0; r <= world; ++r) {
    const long long row = (long long)rx->tile_row0[r] * rx->tile;
    rx->band_row0[r] = (int)(row < res[0] ? row : res[0]);
  }
  const size_t counts_bytes = ((size_t)rx->n_tiles * rx->n_slices * sizeof(uint32_t) + 255) & ~(size_t)255;
  const size_t pairs_bytes = ((size_t)rx->n_tiles * rx->n_slices * rx->slice_cap * sizeof(uint2) + 255) & ~(size_t)255;
  /* every rank allocates the LARGEST band so that offsets agree everywhere */
  const size_t band_rows_max = (size_t)(base + (rem ? 1 : 0)) * rx->tile;
  const size_t band_bytes = band_rows_max * (size_t)res[0] * sizeof(float);
  rx->counts_off = hmrt::kHeaderBytes;
  rx->pairs_off = rx->counts_off + counts_bytes;
  rx->band_off = rx->pairs_off + pairs_bytes;
  rx->region_bytes = rx->band_off + band_bytes;
  cudaError_t e = cudaMalloc(&rx->region, rx->region_bytes);
  if (e == cudaSuccess) e = cudaMemsetAsync(rx->region, 0, hmrt::kHeaderBytes + counts_bytes, ctx->stream);
  if (e != cudaSuccess) {
    if (rx->region) cudaFree(rx->region);
    delete rx;
    return (int)e;
  }
  rx->peer[rank] = rx->region;
  rx->connected = world == 1;
  *out = rx;
  return 0;
}

int hmrt_rx_destroy(hmrt_rx* rx) {
  if (!rx) return HMRT_E_ARG;
  hmrt::DeviceGuard guard(rx->ctx->device);
  cudaStreamSynchronize(rx->ctx->stream);
  for (int r = 0; r < rx->world; ++r)
    if (r != rx->rank && rx->peer[r]) cudaIpcCloseMemHandle(rx->peer[r]);
  if (rx->region) cudaFree(rx->region);
  delete rx;
  return 0;
}

size_t hmrt_rx_region_bytes(const hmrt_rx* rx) { return rx ? rx->region_bytes : 0; }

/* 64-byte cudaIpcMemHandle_t of this rank's exchange region */
int hmrt_rx_export(hmrt_rx* rx, void* handle64) {
  if (!rx || !handle64) return HMRT_E_ARG;
  hmrt::DeviceGuard guard(rx->ctx->device);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaIpcMemHandle_t h;
  HMRT_CUDA(cudaIpcGetMemHandle(&h, rx->region));
  memcpy(handle64, &h, 64);
  return 0;
}

/* handles: world x 64 bytes, rank order (the own entry is ignored).  Collective in spirit: every rank calls it. */
int hmrt_rx_connect(hmrt_rx* rx, const void* handles) {
  if (!rx || !handles) return HMRT_E_ARG;
  if (rx->connected) return 0;
  hmrt::DeviceGuard guard(rx->ctx->device);
  for (int r = 0; r < rx->world; ++r) {
    if (r == rx->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const uint8_t*>(handles) + (size_t)r * 64, 64);
    void* p = nullptr;
    HMRT_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    rx->peer[r] = static_cast<uint8_t*>(p);
  }
  rx->connected = true;
  return 0;
}

/* Start a rasterisation: empty slices, cleared band (the reference's `new float[n]()`, main.cpp:259), no overflow. */
int hmrt_rx_begin(hmrt_rx* rx) {
  if (!rx) return HMRT_E_ARG;
  hmrt::DeviceGuard guard(rx->ctx->device);
  cudaStream_t st = rx->ctx->stream;
  const size_t band_rows = (size_t)(rx->band_row0[rx->rank + 1] - rx->band_row0[rx->rank]);
  HMRT_CUDA(cudaMemsetAsync(rx->region + rx->counts_off, 0, (size_t)rx->n_tiles * rx->n_slices * sizeof(uint32_t), st));
  HMRT_CUDA(cudaMemsetAsync(rx->region + offsetof(hmrt::RxHeader, overflow), 0, sizeof(uint32_t), st));
  if (band_rows) HMRT_CUDA(cudaMemsetAsync(rx->region + rx->band_off, 0, band_rows * (size_t)rx->res0 * sizeof(float), st));
  rx->binned_any = false;
  rx->chunk_cursor = 0;
  return 0;
}

/* Pass 1 on this rank's records (any number of calls between begin and apply). */
int hmrt_rx_bin(hmrt_rx* rx, const uint8_t* d_records, int64_t n, int record_len, int point_format, const hmrt_las_transform* xf) {
  if (!rx || !xf || n < 0 || point_format < 0 || point_format > 3) return HMRT_E_ARG;
  if (record_len < hmrt::kLasMinLen[point_format] || record_len > 64) return HMRT_E_ARG;
  if (n == 0) return 0;
  if (!d_records || (reinterpret_cast<uintptr_t>(d_records) & 15)) return HMRT_E_ARG;
  hmrt::DeviceGuard guard(rx->ctx->device);
  hmrt::BinGeometry bg;
  hmrt::BinParams bp;
  int rc = hmrt::fill_scatter_params(xf, rx->res0, bp.sp);
  if (rc) return rc;
  rc = hmrt::bin_geometry(record_len, false, bp.sp, bg);
  if (rc) return rc;
  bp.sp.cls_off = 15;
  bp.sp.rgb_off = hmrt::kLasRgbOff[point_format];
  bp.records = d_records;
  bp.n = n;
  bp.record_len = record_len;
  bp.per = bg.per;
  bp.counts = reinterpret_cast<uint32_t*>(rx->region + rx->counts_off);
  bp.pairs = rx->region + rx->pairs_off;
  bp.keys = nullptr;
  bp.first_index = 0;
  bp.slice_cap = rx->slice_cap;
  bp.tile_shift = rx->tile_shift;
  bp.tiles_x = rx->tiles_x;
  bp.n_tiles = rx->n_tiles;
  bp.finest = nullptr;
  bp.overflow = reinterpret_cast<uint32_t*>(rx->region + offsetof(hmrt::RxHeader, overflow));
  bp.accumulate = rx->binned_any ? 1 : 0;
  bp.cta_rot = rx->chunk_cursor;
  rx->chunk_cursor = (uint32_t)((rx->chunk_cursor + (n + bg.chunk - 1) / bg.chunk) % rx->n_slices);
  /* always n_slices CTAs: the slice layout is part of the exchange geometry */
  HMRT_CUDA((cudaError_t)hmrt::launch_bin(bg, bp, (unsigned)rx->n_slices, rx->ctx->stream));
  HMRT_LAUNCHED(rx->ctx);
  rx->binned_any = true;
  return 0;
}

/* All ranks have reached this point of their streams (flag exchange in peer memory; no host synchronisation). */
int hmrt_rx_barrier(hmrt_rx* rx) {
  if (!rx) return HMRT_E_ARG;
  if (!rx->connected) return HMRT_E_STATE;
  if (rx->world == 1) return 0;
  hmrt::DeviceGuard guard(rx->ctx->device);
  hmrt::BarrierParams bp;
  memset(&bp, 0, sizeof(bp));
  for (int r = 0; r < rx->world; ++r) bp.peer[r] = rx->peer[r];
  bp.rank = rx->rank, bp.world = rx->world;
  bp.epoch = ++rx->epoch;
  bp.timeout_cycles = 20000000000LL; /* ~10 s at 1.9 GHz */
  hmrt::rx_barrier_kernel<<<1, 32, 0, rx->ctx->stream>>>(bp);
  HMRT_LAUNCHED(rx->ctx);
  return 0;
}

/* Pass 2 of the exchange: pull the pairs of the owned tiles from every rank and reduce them into the own band. */
int hmrt_rx_apply(hmrt_rx* rx) {
  if (!rx) return HMRT_E_ARG;
  if (!rx->connected) return HMRT_E_STATE;
  hmrt::DeviceGuard guard(rx->ctx->device);
  const int rows_owned = rx->tile_row0[rx->rank + 1] - rx->tile_row0[rx->rank];
  if (rows_owned <= 0) return 0;
  hmrt::ApplyParams ap;
  memset(&ap, 0, sizeof(ap));
  for (int r = 0; r < rx->world; ++r) ap.peer[r] = rx->peer[r];
  ap.counts_off = rx->counts_off;
  ap.pairs_off = rx->pairs_off;
  ap.slice_cap = rx->slice_cap;
  ap.n_slices = (uint32_t)rx->n_slices;
  ap.world = (uint32_t)rx->world;
  ap.rank = (uint32_t)rx->rank;
  ap.tile_first = (uint32_t)(rx->tile_row0[rx->rank] * rx->tiles_x);
  ap.pieces_per_slice = (rx->slice_cap + hmrt::kApplyPiece - 1) / hmrt::kApplyPiece;
  ap.groups_per_tile = ((uint32_t)rx->world * (uint32_t)rx->n_slices * ap.pieces_per_slice + (hmrt::kBinThreads / 32) - 1) / (hmrt::kBinThreads / 32);
  ap.dst = reinterpret_cast<int*>(rx->region + rx->band_off);
  ap.cell_base = (uint32_t)rx->band_row0[rx->rank] * (uint32_t)rx->res0;
  ap.dst_cells = (uint32_t)(rx->band_row0[rx->rank + 1] - rx->band_row0[rx->rank]) * (uint32_t)rx->res0;
  const unsigned grid = (unsigned)(rows_owned * rx->tiles_x) * ap.groups_per_tile;
  hmrt::rx_apply_kernel<false><<<grid, hmrt::kBinThreads, 0, rx->ctx->stream>>>(ap);
  HMRT_LAUNCHED(rx->ctx);
  return 0;
}

/* All-gather of the finest level from the owners' bands + every coarser level, into the caller's pyramid. */
int hmrt_rx_gather_mips(hmrt_rx* rx, float* d_pyramid) {
  if (!rx || !d_pyramid) return HMRT_E_ARG;
  if (!rx->connected) return HMRT_E_STATE;
  if ((reinterpret_cast<uintptr_t>(d_pyramid + rx->idx[0]) & 15) || (reinterpret_cast<uintptr_t>(d_pyramid + rx->idx[1]) & 7)) return HMRT_E_ARG;
  hmrt::DeviceGuard guard(rx->ctx->device);
  hmrt::GatherParams g;
  memset(&g, 0, sizeof(g));
  g.mp.pyramid = d_pyramid;
  g.mp.res0 = rx->res0;
  g.mp.out_levels = rx->levels - 1;
  for (int i = 0; i < rx->levels; ++i) g.mp.idx[i] = rx->idx[i];
  for (int r = 0; r < rx->world; ++r) g.band[r] = reinterpret_cast<const float*>(rx->peer[r] + rx->band_off);
  for (int r = 0; r <= rx->world; ++r) g.band_row0[r] = rx->band_row0[r];
  g.world = rx->world;
  g.first_tile_row = rx->band_row0[rx->rank] / 128;
  const dim3 grid(rx->res0 / 128, rx->res0 / 128);
  hmrt::rx_gather_mips_kernel<<<grid, 512, 0, rx->ctx->stream>>>(g);
  HMRT_LAUNCHED(rx->ctx);
  /* nobody may reuse (clear, refill) its region before every rank has finished reading it */
  return hmrt_rx_barrier(rx);
}

/* After a rasterisation (synchronises the stream): points lost to full slices on THIS rank, barrier time-outs. */
int hmrt_rx_status(hmrt_rx* rx, uint32_t* overflow, uint32_t* error) {
  if (!rx) return HMRT_E_ARG;
  hmrt::DeviceGuard guard(rx->ctx->device);
  hmrt::RxHeader h;
  HMRT_CUDA(cudaMemcpyAsync(&h, rx->region, sizeof(h), cudaMemcpyDeviceToHost, rx->ctx->stream));
  HMRT_CUDA(cudaStreamSynchronize(rx->ctx->stream));
  if (overflow) *overflow = h.overflow;
  if (error) *error = h.error;
  return 0;
}

/* rank r owns finest rows [rows[r], rows[r + 1]) (rows has world + 1 entries) */
int hmrt_rx_bands(const hmrt_rx* rx, int* rows) {
  if (!rx || !rows) return HMRT_E_ARG;
  for (int r = 0; r <= rx->world; ++r) rows[r] = rx->band_row0[r];
  return 0;
}

}  // extern "C"
