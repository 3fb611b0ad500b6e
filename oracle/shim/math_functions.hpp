/* TEST INFRASTRUCTURE ONLY. Empty stand-in for the CUDA-8 header the reference includes at
 * CudaKernel.cu:4 (`<math_functions.hpp>`); CUDA 12.9 only ships crt/math_functions.hpp. */
#pragma once
