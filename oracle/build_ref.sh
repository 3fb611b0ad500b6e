#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.
# Builds oracle/_ref/libhmrt_ref.so (+ libhmrt_ref_dpow.so): the reference's OWN device code
# (GPUHeightmapRaytracer/src/CudaKernel.cu lines 1-286, i.e. everything before the <<<>>> host
# wrappers) compiled for the host by plain g++ from where it lies under /root/reference.
# Nothing from the reference is written into the repository: the two mechanical shims below
# live in a mktemp directory that is removed on exit; only the .so files land in oracle/_ref/
# (git-ignored, shipped to the GPU box with the snapshot).
#
#   shim 1: `head -n 286` + closing brace  -> drops the host wrappers g++ cannot parse (<<<>>>)
#   shim 2: sed on a temp copy of CudaKernel.cuh:43-45 -> the MSVC-only functional cast
#           `unsigned char (expr)` becomes `(unsigned char)(expr)` (never executed on this path)
#   oracle/shim/: device_launch_parameters.h (thread-local blockIdx/threadIdx, fp32 pow
#           overload = original MSVC meaning) and an empty math_functions.hpp.
#
# No FMA contraction (x86-64 baseline has no FMA; -ffp-contract=off makes that explicit).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ref="${HMRT_REFERENCE_ROOT:-/root/reference}/GPUHeightmapRaytracer"
out="$here/_ref"
cuda_inc="${CUDA_HOME:-/usr/local/cuda}/include"

if [ ! -f "$ref/src/CudaKernel.cu" ]; then
  echo "build_ref.sh: reference tree not present at $ref (expected on the GPU box): keeping prebuilt $out" >&2
  exit 0
fi

mkdir -p "$out"
tmp="$(mktemp -d)"
trap 'rm -rf "$tmp"' EXIT

sed -e 's/unsigned char (glm::floor/(unsigned char)(glm::floor/' "$ref/src/CudaKernel.cuh" > "$tmp/CudaKernel.cuh"
{ head -n 286 "$ref/src/CudaKernel.cu"; echo "}"; } > "$tmp/ref_device.cpp"

cxxflags=(-std=c++14 -O2 -fPIC -ffp-contract=off -w
          -I"$here/shim" -I"$tmp" -I"$ref/inc" -I"$cuda_inc" -I"$here")

build_variant() {  # $1 = output name, $2... = extra defines
  local name="$1"; shift
  g++ "${cxxflags[@]}" "$@" -c "$tmp/ref_device.cpp" -o "$tmp/ref_device_$name.o"
  g++ "${cxxflags[@]}" "$@" -c "$here/ref_harness.cpp" -o "$tmp/ref_harness_$name.o"
  g++ -shared -o "$out/$name.so" "$tmp/ref_device_$name.o" "$tmp/ref_harness_$name.o" -lpthread
}

build_variant libhmrt_ref -DHMRT_REF_FLOAT_MATH   # canonical: pow(float,int) in fp32 (MSVC/CUDA 8)
build_variant libhmrt_ref_dpow                   # variant: g++/nvcc-Linux promotion to double

# The reference's own HOST-side hot-path code (rasteriser loop, section walk, window composition, camera), cut verbatim out
# of main.cpp by line range and compiled against stub liblas / thread / ifstream types: see refhost_harness.cpp.
build_refhost() {  # $1 = output name, $2 = main.cpp to cut from
  sed -n '44,618p;745,781p' "$2" > "$tmp/$1_cut_main.inc"
  sed -n '995,1003p' "$2" > "$tmp/$1_cut_tables.inc"
  # -fno-aggressive-loop-optimizations: preparePointBuffer's index searches (main.cpp:474,482,490,498) test
  # `origins[maxX][0].x` BEFORE `maxX < point_sections_size`, i.e. they read one element past the inner array on their
  # last trip.  MSVC compiles that as written (a harmless read of the neighbouring row); g++ -O2 treats the out-of-bounds
  # read as proof that the index test can never fail and drops it, and the window then comes from the wrong sections.
  g++ "${cxxflags[@]}" -fno-aggressive-loop-optimizations -DHMRT_REF_FLOAT_MATH -D__device__= -D__global__= -D__host__= \
      -DREFHOST_CUT_MAIN="\"$tmp/$1_cut_main.inc\"" -DREFHOST_CUT_TABLES="\"$tmp/$1_cut_tables.inc\"" \
      -shared -o "$out/$1.so" "$here/refhost_harness.cpp" -lpthread
}
build_refhost libhmrt_refhost "$ref/src/main.cpp"      # verbatim: 4x4 sections, 8 levels (main.cpp:74,83)
# variant with ONLY the two compile-time constants changed (3x3 sections, 4 levels), for small fast cases
sed -e 's/^const int point_sections_size = 4;/const int point_sections_size = 3;/' \
    -e 's/^const int LOD_levels = 8;/const int LOD_levels = 4;/' "$ref/src/main.cpp" > "$tmp/main_g3l4.cpp"
[ "$(diff "$ref/src/main.cpp" "$tmp/main_g3l4.cpp" | grep -c '^>')" = 2 ] || { echo "build_ref.sh: constant patch did not apply" >&2; exit 1; }
build_refhost libhmrt_refhost_g3l4 "$tmp/main_g3l4.cpp"
echo "built $out/libhmrt_refhost.so $out/libhmrt_refhost_g3l4.so (reference main.cpp:44-618,745-781,995-1003 for the host)"

# The reference's CUDA kernel itself, recompiled for sm_100a (baseline "reference kernel on B200").
if command -v nvcc >/dev/null 2>&1; then
  cp "$ref/src/CudaKernel.cu" "$tmp/CudaKernel_ref.cu"   # temp copy so that its #include "CudaKernel.cuh" finds the patched header
  mkdir -p "$tmp/empty" && : > "$tmp/empty/math_functions.hpp"
  sed -i '1s/^\xEF\xBB\xBF//' "$tmp/CudaKernel_ref.cu"
  nvcc -std=c++14 -O3 -gencode arch=compute_100a,code=sm_100a -w -Xcompiler -fPIC -shared \
       -DREF_CU="\"$tmp/CudaKernel_ref.cu\"" -I"$tmp" -I"$tmp/empty" -I"$ref/inc" -I"$here" \
       -include climits -o "$out/libhmrt_ref_gpu.so" "$here/refgpu_harness.cu"
  echo "built $out/libhmrt_ref_gpu.so (reference CUDA kernel for sm_100a)"
fi
echo "built $out/libhmrt_ref.so $out/libhmrt_ref_dpow.so"
