/* Internal declarations shared by the translation units of libhmrt.so (not installed). */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ray_core.cuh"

#define HMRT_CUDA(call)                          \
  do {                                           \
    cudaError_t e_ = (call);                     \
    if (e_ != cudaSuccess) return (int)e_;       \
  } while (0)

/* kernel launches are checked with cudaGetLastError (cudaPeekAtLastError keeps sticky errors) */
/* HMRT_DCHECK (device-side bounds checks of the checked build) is defined in ray_core.cuh, the header every kernel source includes. */
#define HMRT_LAUNCHED(ctx)                       \
  do {                                           \
    cudaError_t e_ = cudaGetLastError();         \
    if (e_ != cudaSuccess) return (int)e_;       \
    (ctx)->launches++;                           \
  } while (0)

struct hmrt_ctx {
  int device;
  int sm_count;
  cudaStream_t stream;
  int64_t launches;
  int trace_variant; /* 0 = production walk, 1 = operation-by-operation walk (diagnostic), 2 = tolerance mode (air phase in one step) */
  /* rasterisation: 0 = choose per call (locality probe), 1 = direct atomics, 2 = tile-binned */
  int scatter_mode;
  void* d_ws; /* bucket workspace of the binned scatter */
  size_t ws_cap;
  uint32_t* d_probe;
  int probe_verdict; /* -1 = no locality probe yet, 0 = spatially ordered input (direct atomics), 1 = unordered (tile-binned) */
  /* borrowed heightmap (hmrt_set_heightmap) */
  bool have_grid;
  hmrt::Grid grid;
  float init_max_height;
  void* d_scratch;    /* TraceScratch[kCallSets][scratch_cap]: max(top level) key + one work counter per launch of a call */
  int scratch_cap;
  cudaStream_t last_trace_stream; /* stream of the previous hmrt_trace call: a caller that alternates streams overlaps its calls */
  bool have_last_trace_stream;
  int call_set;       /* scratch / frame-constant set of the latest trace call (round robin over kCallSets) */
  unsigned long long* d_stats; /* counters of the instrumented kernels (hmrt_trace_stats) */
  void* d_seg_done;   /* SegDone[seg_cap]: per-(frame, row segment) completion of hmrt_trace_host's streaming form */
  int seg_cap;
  int host_mid_groups; /* development knob: frame groups per middle launch of the per-group host schedule (0 = built-in 2) */
  int host_segments;  /* development knob: row segments per frame of the streamed host schedule (0 = built-in) */
  int host_variant;   /* hmrt_trace_host: 0 = auto, 1 = one launch per frame group, 2 = streamed (segment flags + stream memory operations) */
  int window_variant; /* 0 = TMA bulk-copy gather where the planes qualify, 1 = per-thread 128-bit gather (hmrt_set_window_variant) */
  int l2_first_level; /* experiment (hmrt_set_l2_persist): levels >= this get a persisting-L2 access policy window; -1 = off */
  float l2_hit_ratio;
  int ctas_per_sm[32]; /* resident CTAs per SM of each trace kernel instantiation (0 = not queried yet) */
  /* per-frame constants for multi-frame launches */
  hmrt::FrameConsts* d_frames;
  int frames_cap;
  /* context-owned framebuffers for hmrt_trace_host: kHostCalls halves, one per call in flight */
  uint8_t* d_fb;
  size_t fb_cap;
  size_t fb_half;                /* bytes per half */
  unsigned host_begun, host_waited; /* hmrt_trace_host_begin / _wait: calls enqueued / collected */
  cudaEvent_t host_done[2];      /* recorded on the copy stream behind the last copy of a call */
  /* hmrt_trace_host: frames alternate between two streams, copies run on a third */
  cudaStream_t copy_stream;
  cudaStream_t frame_stream[2];
  cudaEvent_t frame_event[2];
  cudaEvent_t prep_event;
};

namespace hmrt {

struct DeviceGuard {
  int prev;
  bool switched;
  explicit DeviceGuard(int dev) : prev(-1), switched(false) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) {
      cudaSetDevice(dev);
      switched = true;
    }
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

int pyramid_layout(int coarse_res, int levels, int* res, int64_t* idx, int64_t* total);

}  // namespace hmrt
