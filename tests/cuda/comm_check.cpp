/*
 * comm_check.cpp -- the C-ABI collective steps (hmrt_allreduce_max_heights, hmrt_broadcast_heightmap) driven the way a
 * C++ host would: one process, one hmrt context + one NCCL communicator per visible GPU (ncclCommInitAll), per-rank calls
 * inside an NCCL group.  Checks the results bit for bit against the element-wise maximum / the root's buffers computed on
 * the host.  Works with a single GPU too (communicator of size 1).
 *   g++ -std=c++17 comm_check.cpp -I../../include -I/usr/local/cuda/include -L../../gpu-heightmap-raytracer_b200/csrc -lhmrt
 *       -L/usr/local/cuda/lib64 -lcudart -lnccl -o comm_check
 */
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "hmrt.h"

#define CK(x)                                                                  \
  do {                                                                         \
    const long long rc_ = (long long)(x);                                      \
    if (rc_ != 0) {                                                            \
      std::printf("comm_check: %s failed: %lld (line %d)\n", #x, rc_, __LINE__); \
      return 1;                                                                \
    }                                                                          \
  } while (0)

static uint64_t mix(uint64_t x) {
  x ^= x >> 33, x *= 0xff51afd7ed558ccdull, x ^= x >> 33, x *= 0xc4ceb9fe1a85ec53ull, x ^= x >> 33;
  return x;
}

int main() {
  int n = 0;
  CK(cudaGetDeviceCount(&n));
  if (n < 1) return 2;
  if (n > 8) n = 8;
  const int coarse = 16, levels = 5;
  int res[HMRT_MAX_LEVELS];
  int64_t idx[HMRT_MAX_LEVELS], total = 0;
  CK(hmrt_pyramid_layout(coarse, levels, res, idx, &total));
  const size_t cells = (size_t)res[0] * res[0];
  std::vector<ncclComm_t> comms(n);
  CK(ncclCommInitAll(comms.data(), n, nullptr));
  std::vector<hmrt_ctx*> ctx(n);
  std::vector<float*> d_pyr(n);
  std::vector<uint64_t*> d_keys(n);
  std::vector<hmrt_color*> d_col(n);
  std::vector<std::vector<float>> h_pyr(n, std::vector<float>((size_t)total));
  std::vector<std::vector<uint64_t>> h_keys(n, std::vector<uint64_t>(cells));
  std::vector<std::vector<uint8_t>> h_col(n, std::vector<uint8_t>(cells * 3));
  for (int r = 0; r < n; ++r) {
    CK(cudaSetDevice(r));
    CK(hmrt_create(r, &ctx[r]));
    CK(cudaMalloc(&d_pyr[r], sizeof(float) * (size_t)total));
    CK(cudaMalloc(&d_keys[r], sizeof(uint64_t) * cells));
    CK(cudaMalloc(&d_col[r], cells * 3));
    for (int64_t i = 0; i < total; ++i) { /* heights >= +0, many exact zeros (cells no point of this rank fell into) */
      const uint64_t h = mix((uint64_t)i * 8 + r);
      h_pyr[r][(size_t)i] = (h & 3) == 0 ? 0.0f : (float)(h >> 40) / 1024.0f;
    }
    for (size_t i = 0; i < cells; ++i) {
      const uint64_t h = mix(i * 16 + r + 1000);
      h_keys[r][i] = (h & 1) ? 0 : (((h >> 8) % 500000000ull + 1) << 24) | (h >> 40 & 0xffffff);
      for (int c = 0; c < 3; ++c) h_col[r][i * 3 + c] = (uint8_t)mix(i * 3 + c + 77 * r);
    }
    CK(cudaMemcpy(d_pyr[r], h_pyr[r].data(), sizeof(float) * (size_t)total, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_keys[r], h_keys[r].data(), sizeof(uint64_t) * cells, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_col[r], h_col[r].data(), cells * 3, cudaMemcpyHostToDevice));
  }
  /* ---- max all-reduce of the finest level + colour keys ---- */
  CK(ncclGroupStart());
  for (int r = 0; r < n; ++r) CK(hmrt_allreduce_max_heights(ctx[r], comms[r], d_pyr[r], d_keys[r], coarse, levels));
  CK(ncclGroupEnd());
  std::vector<float> want(h_pyr[0]);
  std::vector<uint64_t> want_keys(h_keys[0]);
  for (int r = 1; r < n; ++r) {
    for (size_t i = 0; i < cells; ++i) {
      float& w = want[(size_t)idx[0] + i];
      if (h_pyr[r][(size_t)idx[0] + i] > w) w = h_pyr[r][(size_t)idx[0] + i];
      if (h_keys[r][i] > want_keys[i]) want_keys[i] = h_keys[r][i];
    }
  }
  std::vector<float> got((size_t)total);
  std::vector<uint64_t> got_keys(cells);
  for (int r = 0; r < n; ++r) {
    CK(cudaSetDevice(r));
    CK(hmrt_synchronize(ctx[r]));
    CK(cudaMemcpy(got.data(), d_pyr[r], sizeof(float) * (size_t)total, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(got_keys.data(), d_keys[r], sizeof(uint64_t) * cells, cudaMemcpyDeviceToHost));
    /* finest level = element-wise max; the coarser levels (offsets below idx[0]) are this rank's own, untouched */
    CK(std::memcmp(got.data() + idx[0], want.data() + idx[0], sizeof(float) * cells));
    CK(std::memcmp(got.data(), h_pyr[r].data(), sizeof(float) * (size_t)idx[0]));
    CK(std::memcmp(got_keys.data(), want_keys.data(), sizeof(uint64_t) * cells));
  }
  /* ---- broadcast of pyramid + colour map from the last rank ---- */
  const int root = n - 1;
  CK(ncclGroupStart());
  for (int r = 0; r < n; ++r) CK(hmrt_broadcast_heightmap(ctx[r], comms[r], d_pyr[r], d_col[r], coarse, levels, root));
  CK(ncclGroupEnd());
  std::vector<float> root_pyr((size_t)total);
  std::vector<uint8_t> root_col(cells * 3), got_col(cells * 3);
  CK(cudaSetDevice(root));
  CK(hmrt_synchronize(ctx[root]));
  CK(cudaMemcpy(root_pyr.data(), d_pyr[root], sizeof(float) * (size_t)total, cudaMemcpyDeviceToHost));
  for (int r = 0; r < n; ++r) {
    CK(cudaSetDevice(r));
    CK(hmrt_synchronize(ctx[r]));
    CK(cudaMemcpy(got.data(), d_pyr[r], sizeof(float) * (size_t)total, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(got_col.data(), d_col[r], cells * 3, cudaMemcpyDeviceToHost));
    CK(std::memcmp(got.data(), root_pyr.data(), sizeof(float) * (size_t)total));
    CK(std::memcmp(got_col.data(), h_col[root].data(), cells * 3));
  }
  /* argument errors do not reach NCCL */
  if (hmrt_allreduce_max_heights(ctx[0], nullptr, d_pyr[0], nullptr, coarse, levels) != HMRT_E_ARG) return 3;
  if (hmrt_broadcast_heightmap(ctx[0], comms[0], nullptr, nullptr, coarse, levels, 0) != HMRT_E_ARG) return 3;
  for (int r = 0; r < n; ++r) {
    cudaSetDevice(r);
    cudaFree(d_pyr[r]), cudaFree(d_keys[r]), cudaFree(d_col[r]);
    hmrt_destroy(ctx[r]);
    ncclCommDestroy(comms[r]);
  }
  std::printf("comm_check: ok on %d GPU(s): max all-reduce (finest level + colour keys) and heightmap broadcast are exact\n", n);
  return 0;
}
