// Host emulation shim (TEST INFRASTRUCTURE): lets g++ compile the element-wise CUDA kernels of csrc/sepconv_elem.cu as plain
// C++ so that tests/test_sepconv_elem_host.py can run them on the CPU (one "thread", grid-stride loops cover everything)
// against numpy / torch restatements.  The tensor-core kernels are not emulated.  Never part of the product build.
#pragma once
#define VFI_HOST_EMU 1
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#undef __global__
#undef __device__
#undef __host__
#undef __forceinline__
#undef __shared__
#undef __restrict__
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __restrict__
struct EmuDim {
  unsigned x = 0, y = 0, z = 0;
};
static EmuDim blockIdx, threadIdx;
static EmuDim blockDim{1, 1, 1}, gridDim{1, 1, 1};
static inline void __syncthreads() {}
static inline double atomicAdd(double* p, double v) {
  const double o = *p;
  *p += v;
  return o;
}
using std::max;
using std::min;
