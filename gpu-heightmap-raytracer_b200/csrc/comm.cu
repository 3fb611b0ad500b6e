/*
 * comm.cu -- the two collective steps of the path for C / C++ hosts, on a caller-provided NCCL
 * communicator.  NCCL is bound at run time (dlopen "libnccl.so.2"), so libhmrt.so itself neither links nor ships NCCL and
 * shares whatever copy the process already uses.
 */
#include <dlfcn.h>
#include <stdio.h>

#include "hmrt_internal.cuh"

namespace hmrt {

/* the stable part of nccl.h (NCCL 2.x): result code 0 = success; data types and reduction operators by value */
enum { kNcclUint8 = 1, kNcclInt32 = 2, kNcclInt64 = 4, kNcclFloat32 = 7, kNcclMax = 2 };
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_broadcast_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*nccl_errstr_fn)(int);

struct Nccl {
  nccl_allreduce_fn all_reduce = nullptr;
  nccl_broadcast_fn broadcast = nullptr;
  nccl_errstr_fn err = nullptr;
  bool ok = false;
};

static const Nccl& nccl() {
  static const Nccl lib = [] {
    Nccl n;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return n;
    n.all_reduce = reinterpret_cast<nccl_allreduce_fn>(dlsym(h, "ncclAllReduce"));
    n.broadcast = reinterpret_cast<nccl_broadcast_fn>(dlsym(h, "ncclBroadcast"));
    n.err = reinterpret_cast<nccl_errstr_fn>(dlsym(h, "ncclGetErrorString"));
    n.ok = n.all_reduce && n.broadcast;
    return n;
  }();
  return lib;
}

static int nccl_check(int rc, const char* what) {
  if (rc == 0) return 0;
  fprintf(stderr, "hmrt: %s failed: NCCL error %d (%s)\n", what, rc, nccl().err ? nccl().err(rc) : "?");
  return HMRT_E_NCCL;
}

}  // namespace hmrt

extern "C" {

int hmrt_broadcast_heightmap(hmrt_ctx* ctx, void* nccl_comm, float* d_pyramid, hmrt_color* d_color_map, int coarse_res, int levels,
                             int root) {
  if (!ctx || !nccl_comm || !d_pyramid || root < 0) return HMRT_E_ARG;
  int res[HMRT_MAX_LEVELS];
  int64_t total = 0;
  int rc = hmrt::pyramid_layout(coarse_res, levels, res, nullptr, &total);
  if (rc) return rc;
  if (!hmrt::nccl().ok) return HMRT_E_NCCL;
  hmrt::DeviceGuard guard(ctx->device);
  rc = hmrt::nccl_check(hmrt::nccl().broadcast(d_pyramid, d_pyramid, (size_t)total, hmrt::kNcclFloat32, root, nccl_comm, ctx->stream),
                        "ncclBroadcast(pyramid)");
  if (rc == 0 && d_color_map)
    rc = hmrt::nccl_check(hmrt::nccl().broadcast(d_color_map, d_color_map, (size_t)res[0] * res[0] * 3, hmrt::kNcclUint8, root, nccl_comm, ctx->stream),
                          "ncclBroadcast(colour map)");
  return rc;
}

int hmrt_allreduce_max_heights(hmrt_ctx* ctx, void* nccl_comm, float* d_pyramid, uint64_t* d_color_keys, int coarse_res, int levels) {
  if (!ctx || !nccl_comm || !d_pyramid) return HMRT_E_ARG;
  int res[HMRT_MAX_LEVELS];
  int64_t idx[HMRT_MAX_LEVELS];
  int rc = hmrt::pyramid_layout(coarse_res, levels, res, idx, nullptr);
  if (rc) return rc;
  if (!hmrt::nccl().ok) return HMRT_E_NCCL;
  hmrt::DeviceGuard guard(ctx->device);
  const size_t cells = (size_t)res[0] * res[0];
  float* finest = d_pyramid + idx[0];
  /* int32 view: non-negative floats order like their bit patterns, so the integer MAX is the float max, exactly */
  rc = hmrt::nccl_check(hmrt::nccl().all_reduce(finest, finest, cells, hmrt::kNcclInt32, hmrt::kNcclMax, nccl_comm, ctx->stream),
                        "ncclAllReduce(finest level, max)");
  if (rc == 0 && d_color_keys) /* keys = (file index + 1) << 24 | rgb < 2^63: int64 MAX = last writer in file order */
    rc = hmrt::nccl_check(hmrt::nccl().all_reduce(d_color_keys, d_color_keys, cells, hmrt::kNcclInt64, hmrt::kNcclMax, nccl_comm, ctx->stream),
                          "ncclAllReduce(colour keys, max)");
  return rc;
}

}  // extern "C"
