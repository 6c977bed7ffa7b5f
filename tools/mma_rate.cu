// Micro-benchmark (profiling tool, not part of the library): issue rate of tcgen05.mma.cta_group::1.kind::f16
// M=128 x N x K=16 with both operands in shared memory (SWIZZLE_128B, K-major), as tapconv issues them, for
// N = 16..256, with (a) aligned operand start addresses and SBO = 1024 B and (b) the tap-shifted form tapconv
// uses (start address shifted by whole 128-byte rows, SBO = halo_w * 128 = 1280 B, K step +32 B).
// One CTA per SM, one thread issues `iters` x 36 MMAs into one accumulator, commit, wait; cycles per MMA printed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I comfyui-frame-interpolation_b200/csrc \
//        tools/mma_rate.cu -o gpurun_out/mma_rate && gpurun_out/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace vfi;

struct Res { long long cycles; };

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int shifted, int iters, int pattern, int commit_every, Res* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = smem_u32(smem);
  // zero the operands (values do not matter for timing; zeros keep the accumulator finite)
  for (int i = threadIdx.x; i < (64 * 1024) / 2; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + 12345u * (blockIdx.x + 1);
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    const float f = pattern ? ((float)(h & 0xffff) / 32768.f - 1.f) * 0.5f : 0.f;
    reinterpret_cast<__half*>(smem)[i] = __float2half(f);
  }
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_fence_init();
  }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_slot), 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t a_base = base;               // A window: 24 KB (18 x 10 rows of 128 B)
  const uint32_t b_base = base + 32 * 1024;   // B: up to 256 rows x 128 B = 32 KB
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t a_hi = ((shifted ? 1280u : 1024u) >> 4) | (1u << 14) | (2u << 29);
  const uint32_t b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t tm = tmem + ((commit_every && (it & 1)) ? 128u : 0u);  // alternate two accumulators
#pragma unroll
      for (int j = 0; j < 36; ++j) {
        const int tap = j >> 2, k = j & 3;
        const uint32_t a_off = shifted ? (uint32_t)((tap / 3) * 10 + (tap % 3)) * 128u + k * 32u : k * 32u;
        const uint32_t a_lo = (1u << 16) | ((a_base + a_off) >> 4);
        const uint32_t b_lo = (1u << 16) | ((b_base + k * 32u) >> 4);
        umma_f16_split(tm, a_lo, a_hi, b_lo, b_hi, idesc, (commit_every ? j : (it | j)) ? 1u : 0u);
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0, 99);
    t1 = clock64();
    out[blockIdx.x].cycles = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

int main() {
  int dev = 0, sms = 0;
  cudaSetDevice(dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  Res* d;
  cudaMalloc(&d, sizeof(Res) * sms);
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 200;
  printf("-- N = 64, tap-shifted: zeros vs random fp16 operands, one accumulator vs two alternating (fresh each 36)\n");
  for (int pattern = 0; pattern < 2; ++pattern)
    for (int ce = 0; ce < 2; ++ce) {
      for (int rep = 0; rep < 2; ++rep) mma_rate_kernel<<<sms, 128, 64 * 1024>>>(64, 1, iters, pattern, ce, d);
      cudaDeviceSynchronize();
      Res h[256];
      cudaMemcpy(h, d, sizeof(Res) * sms, cudaMemcpyDeviceToHost);
      double mean = 0;
      for (int i = 0; i < sms; ++i) mean += (double)h[i].cycles;
      printf("pattern %d two_acc %d: %.1f cycles/MMA\n", pattern, ce, mean / sms / (iters * 36.0));
    }
  printf("tcgen05.mma M128 x N x K16 (SS, SW128 K-major), %d SMs, %d MMAs per CTA\n", sms, iters * 36);
  printf("%5s %8s %14s %14s %12s\n", "N", "shifted", "cycles/MMA", "math cyc@4096", "TF/s all SMs");
  for (int shifted = 0; shifted < 2; ++shifted)
    for (int N : {16, 32, 48, 64, 80, 96, 128, 192, 256}) {
      for (int rep = 0; rep < 2; ++rep) mma_rate_kernel<<<sms, 128, 64 * 1024>>>(N, shifted, iters, 0, 0, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
      Res h[256];
      cudaMemcpy(h, d, sizeof(Res) * sms, cudaMemcpyDeviceToHost);
      double mean = 0;
      for (int i = 0; i < sms; ++i) mean += (double)h[i].cycles;
      mean /= sms;
      const double per = mean / (iters * 36.0);
      int clk_khz = 0;
      cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
      const double tf = 2.0 * 128 * N * 16 / per * (clk_khz * 1e3) * sms / 1e12;
      printf("%5d %8d %14.1f %14.1f %12.0f\n", N, shifted, per, 128.0 * N * 16 / 4096.0, tf);
    }
  return 0;
}
