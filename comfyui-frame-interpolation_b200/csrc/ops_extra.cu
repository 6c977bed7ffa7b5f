// The two remaining names of the reference's op surface (vfi_models/ops/__init__.py:19-21): FunctionAdaCoF (ST-MFNet) and
// batch_edt (EISAI).  Contiguous NCHW fp32 like the reference's cupy launches; both are small HBM / latency-bound kernels.
//
//   adacof   kernel_AdaCoF_updateOutput, cupy_ops/adacof.py:5-62: out[n,c,i,j] = sum_{k,l} w[n,kF+l,i,j] * bilinear(in[n,c],
//            i + k d + alpha, j + l d + beta) with A = (int) alpha (truncation toward zero, NOT floor), clamped tap
//            indices and the unclamped fractions (alpha - A), as written there.  One thread per output pixel: weight and
//            offsets are read once per tap and reused over the channels (the reference re-reads them per channel).
//   edt_pass kernel_dt, cupy_ops/batch_edt.py:9-41: out[b,i,j] = min(diam2, min_j' data[b,i,j'] + (j - j')^2): one of the two
//            passes of the separable squared distance transform (batch_edt runs it on rows, transposes, runs it again).
#include "ptx.cuh"
#include "vfi_internal.h"

namespace vfi {

namespace {

inline int xgrid(size_t total, int threads) {
  size_t b = (total + threads - 1) / threads;
  return (int)(b < 1 ? 1 : (b > 1048576 ? 1048576 : b));
}

__global__ void adacof_kernel(const float* __restrict__ in, const float* __restrict__ weight, const float* __restrict__ off_i,
                              const float* __restrict__ off_j, float* __restrict__ out, int N, int C, int Hin, int Win, int F,
                              int dil, int Ho, int Wo) {
  const size_t total = (size_t)N * Ho * Wo;
  const size_t plane_o = (size_t)Ho * Wo, plane_i = (size_t)Hin * Win;
  for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(id % Wo);
    const int i = (int)((id / Wo) % Ho);
    const int n = (int)(id / plane_o);
    const size_t px = (size_t)i * Wo + j;
    for (int c0 = 0; c0 < C; c0 += 8) {  // eight channels per sweep over the taps
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const int nc = min(8, C - c0);
      for (int k = 0; k < F; ++k)
        for (int l = 0; l < F; ++l) {
          const size_t t = ((size_t)n * F * F + (size_t)k * F + l) * plane_o + px;
          const float w = weight[t], alpha = off_i[t], beta = off_j[t];
          const int A = (int)alpha, B = (int)beta;
          const int y0 = min(max(i + k * dil + A, 0), Hin - 1), y1 = min(max(i + k * dil + A + 1, 0), Hin - 1);
          const int x0 = min(max(j + l * dil + B, 0), Win - 1), x1 = min(max(j + l * dil + B + 1, 0), Win - 1);
          const float fa = alpha - (float)A, fb = beta - (float)B;
          const float w00 = (1.f - fa) * (1.f - fb), w10 = fa * (1.f - fb), w01 = (1.f - fa) * fb, w11 = fa * fb;
          const float* base = in + ((size_t)n * C + c0) * plane_i;
          for (int c = 0; c < nc; ++c) {
            const float* p = base + (size_t)c * plane_i;
            acc[c] += w * (p[(size_t)y0 * Win + x0] * w00 + p[(size_t)y1 * Win + x0] * w10 + p[(size_t)y0 * Win + x1] * w01 +
                           p[(size_t)y1 * Win + x1] * w11);
          }
        }
      for (int c = 0; c < nc; ++c) out[((size_t)n * C + c0 + c) * plane_o + px] = acc[c];
    }
  }
}

__global__ void edt_pass_kernel(const float* __restrict__ data, float* __restrict__ out, int bs, int h, int w, float diam2) {
  const size_t total = (size_t)bs * h * w;
  for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (size_t)gridDim.x * blockDim.x) {
    const int pj = (int)(id % w);
    const float* row = data + (id - pj);
    float best = diam2;
    for (int j = 0; j < w; ++j) {
      const float cost = row[j] + (float)((pj - j) * (pj - j));
      if (cost < best) best = cost;
    }
    out[id] = best;
  }
}

}  // namespace

cudaError_t launch_adacof(const float* in, const float* weight, const float* off_i, const float* off_j, float* out, int N,
                          int C, int Hin, int Win, int F, int dil, int Ho, int Wo, cudaStream_t st) {
  VFI_LAUNCH(adacof_kernel, xgrid((size_t)N * Ho * Wo, 128), 128, 0, st, in, weight, off_i, off_j, out, N, C, Hin, Win, F, dil,
             Ho, Wo);
  return cudaGetLastError();
}

cudaError_t launch_edt_pass(const float* data, float* out, int bs, int h, int w, float diam2, cudaStream_t st) {
  VFI_LAUNCH(edt_pass_kernel, xgrid((size_t)bs * h * w, 256), 256, 0, st, data, out, bs, h, w, diam2);
  return cudaGetLastError();
}

}  // namespace vfi
