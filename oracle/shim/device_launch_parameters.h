/*
 * TEST INFRASTRUCTURE ONLY (oracle/_ref build) -- never part of the product path.
 *
 * Host stand-in for CUDA's <device_launch_parameters.h>, found first on the include path when
 * the reference's device code (GPUHeightmapRaytracer/src/CudaKernel.cu:1-286, included through
 * CudaKernel.cuh:12) is compiled by plain g++.  It gives the reference's `blockIdx/threadIdx`
 * (CudaKernel.cu:200-201) a per-host-thread meaning so the harness can "launch" one pixel at a
 * time, and supplies the two things the original MSVC/CUDA-8 toolchain provided implicitly:
 *   - <climits> for USHRT_MAX (CudaKernel.cuh:43-45),
 *   - the fp32 overloads `float pow(float,int)` and `float floor(float)` (SURVEY.md section 8(c),
 *     "pow fork"): under MSVC 2015 / CUDA 8 the calls `pow(2.f, LOD)` and `floor(x / pow(..))`
 *     (CudaKernel.cu:77-89,101,134,145) are evaluated entirely in fp32.  Plain g++ resolves both
 *     to the double versions (checked with static_asserts: decltype(floor(1.0f)) is double here),
 *     nvcc-on-Linux resolves floor to float but pow to double -- either way tX/tZ
 *     (CudaKernel.cu:77-78) would be computed in double and rounded once.
 *     HMRT_REF_FLOAT_MATH selects the original all-fp32 meaning (canonical for this repo);
 *     without it the build keeps the Linux double meaning (reported as a variant).
 */
#pragma once
#include <vector_types.h>
#include <climits>
#include <cmath>

extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;

#ifdef HMRT_REF_FLOAT_MATH
namespace CudaSpace {
inline float pow(float a, int b) { return ::powf(a, static_cast<float>(b)); }
inline float floor(float a) { return ::floorf(a); }
}
#endif
