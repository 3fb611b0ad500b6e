/*
 * rcpcheck.cu -- enumerates every fp32 value d with 2^-40 <= |d| <= 2 (the range fast_div_ok admits) and checks that
 * hmrt::rcp_rn_inrange(d), the reciprocal the production walk uses, has the bits of the IEEE reciprocal __frcp_rn(d).
 * The function under test is the shipped one (included from csrc/ray_fast.cuh).
 *   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I../../include rcpcheck.cu -o rcpcheck && ./rcpcheck
 */
#include <cstdio>
#include <cuda_runtime.h>

#include "../../gpu-heightmap-raytracer_b200/csrc/ray_fast.cuh"

__global__ void check(unsigned long long* bad, unsigned* first_bad) {
  /* exponent fields 127-40 .. 127+1 (2.0 itself is the single value with field 128) */
  const unsigned lo = (127u - 40u) << 23, hi = (128u << 23);
  unsigned long long local = 0;
  for (unsigned long long b = lo + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; b <= hi;
       b += (unsigned long long)gridDim.x * blockDim.x) {
    for (unsigned sign = 0; sign < 2; ++sign) {
      const float d = __uint_as_float((unsigned)b | (sign << 31));
      if (!hmrt::fast_div_ok(d)) {
        ++local; /* the enumeration must stay inside the guarded range */
        continue;
      }
      const unsigned got = __float_as_uint(hmrt::rcp_rn_inrange(d)), want = __float_as_uint(__frcp_rn(d));
      if (got != want) {
        if (local == 0) first_bad[0] = __float_as_uint(d);
        ++local;
      }
    }
  }
  if (local) atomicAdd(bad, local);
}

int main() {
  unsigned long long* bad;
  unsigned* first_bad;
  cudaMallocManaged(&bad, sizeof(*bad));
  cudaMallocManaged(&first_bad, sizeof(unsigned));
  *bad = 0;
  *first_bad = 0;
  check<<<148 * 8, 256>>>(bad, first_bad);
  const cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) {
    printf("CUDA error: %s\n", cudaGetErrorString(err));
    return 2;
  }
  const unsigned long long n = 2ull * (((128ull - 87ull) << 23) + 1ull);
  printf("rcpcheck: %llu operands with 2^-40 <= |d| <= 2, mismatches vs __frcp_rn: %llu", n, *bad);
  if (*bad) printf("  (one of them: d=0x%08x)", *first_bad);
  printf("\n");
  return *bad ? 1 : 0;
}
