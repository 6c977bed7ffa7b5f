// The reference's custom CUDA ops (vfi_models/ops/cupy_ops, CUDA-C strings JIT-compiled through cupy/NVRTC), rebuilt as
// ahead-of-time sm_100a kernels behind the C ABI.  Same tensor contract as the reference: contiguous NCHW fp32.
//   softsplat_sum  : forward summation splat                      cupy_ops/softsplat.py:140-192
//   costvol_l1     : 9x9 L1 cost volume (81 ch)                   cupy_ops/costvol.py:4-43
//   corr_dot       : 9x9 dot-product correlation (81 ch)          cupy_ops/correlation.py:4-99
//   sepconv        : adaptive separable convolution, Kahan sums   cupy_ops/sepconv.py:86-117
// One thread per pixel computes everything that does not depend on the channel ONCE (the reference launches one
// thread per (pixel, channel) and recomputes flow taps / weights C times) and keeps the 81 volume entries in registers.
#include "vfi_internal.h"

namespace vfi {
namespace {

__global__ void softsplat_sum_kernel(const float* __restrict__ in, const float* __restrict__ flow,
                                     float* __restrict__ out, int N, int C, int H, int W) {
  const size_t hw = (size_t)H * W;
  const size_t total = (size_t)N * hw;
  for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(id % W);
    const int y = (int)((id / W) % H);
    const int n = (int)(id / hw);
    const size_t pix = (size_t)y * W + x;
    const float fx = (float)x + __ldg(flow + ((size_t)n * 2 + 0) * hw + pix);
    const float fy = (float)y + __ldg(flow + ((size_t)n * 2 + 1) * hw + pix);
    if (!isfinite(fx) || !isfinite(fy)) continue;  // softsplat.py:158-159
    const int x0 = (int)floorf(fx), y0 = (int)floorf(fy);
    const int x1 = x0 + 1, y1 = y0 + 1;
    const float wnw = ((float)x1 - fx) * ((float)y1 - fy);
    const float wne = (fx - (float)x0) * ((float)y1 - fy);
    const float wsw = ((float)x1 - fx) * (fy - (float)y0);
    const float wse = (fx - (float)x0) * (fy - (float)y0);
    const bool inx0 = x0 >= 0 && x0 < W, inx1 = x1 >= 0 && x1 < W;
    const bool iny0 = y0 >= 0 && y0 < H, iny1 = y1 >= 0 && y1 < H;
    const float* src = in + (size_t)n * C * hw + pix;
    float* dst = out + (size_t)n * C * hw;
    for (int c = 0; c < C; ++c) {
      const float v = __ldg(src + (size_t)c * hw);
      float* d = dst + (size_t)c * hw;
      if (inx0 && iny0) atomicAdd(d + (size_t)y0 * W + x0, v * wnw);
      if (inx1 && iny0) atomicAdd(d + (size_t)y0 * W + x1, v * wne);
      if (inx0 && iny1) atomicAdd(d + (size_t)y1 * W + x0, v * wsw);
      if (inx1 && iny1) atomicAdd(d + (size_t)y1 * W + x1, v * wse);
    }
  }
}

// kDot = false: mean_c |one - two(shifted)|, outside -> mean_c |one|      (costvol)
// kDot = true : mean_c one * two(shifted), outside -> 0                   (correlation, zero padded)
template <bool kDot>
__global__ void volume81_kernel(const float* __restrict__ one, const float* __restrict__ two, float* __restrict__ out,
                                int N, int C, int H, int W) {
  const size_t hw = (size_t)H * W;
  const size_t total = (size_t)N * hw;
  for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(id % W);
    const int y = (int)((id / W) % H);
    const int n = (int)(id / hw);
    float acc[81];
#pragma unroll
    for (int k = 0; k < 81; ++k) acc[k] = 0.f;
    for (int c = 0; c < C; ++c) {
      const float a = __ldg(one + ((size_t)n * C + c) * hw + (size_t)y * W + x);
      const float* t = two + ((size_t)n * C + c) * hw;
#pragma unroll
      for (int dy = -4; dy <= 4; ++dy) {
        const int yy = y + dy;
        const bool yin = yy >= 0 && yy < H;
#pragma unroll
        for (int dx = -4; dx <= 4; ++dx) {
          const int xx = x + dx;
          const float b = (yin && xx >= 0 && xx < W) ? __ldg(t + (size_t)yy * W + xx) : 0.f;
          const int k = (dy + 4) * 9 + (dx + 4);
          if (kDot)
            acc[k] = fmaf(a, b, acc[k]);
          else
            acc[k] += fabsf(a - b);
        }
      }
    }
    const float inv = 1.f / (float)C;
    float* o = out + (size_t)n * 81 * hw + (size_t)y * W + x;
#pragma unroll
    for (int k = 0; k < 81; ++k) o[(size_t)k * hw] = kDot ? acc[k] / (float)C : acc[k] / (float)C;
    (void)inv;
  }
}

// in [N, C, H+Kv-1, W+Kh-1], ver [N, Kv, H, W], hor [N, Kh, H, W] -> out [N, C, H, W]; C <= 4 per pass
template <int CB>
__global__ void sepconv_kernel(const float* __restrict__ in, const float* __restrict__ ver,
                               const float* __restrict__ hor, float* __restrict__ out, int N, int C, int c0, int H,
                               int W, int Kv, int Kh) {
  const size_t hw = (size_t)H * W;
  const size_t total = (size_t)N * hw;
  const int Hp = H + Kv - 1, Wp = W + Kh - 1;
  const size_t hwp = (size_t)Hp * Wp;
  for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(id % W);
    const int y = (int)((id / W) % H);
    const int n = (int)(id / hw);
    const size_t pix = (size_t)y * W + x;
    float sum[CB], comp[CB];
#pragma unroll
    for (int c = 0; c < CB; ++c) sum[c] = comp[c] = 0.f;
    const float* vp = ver + (size_t)n * Kv * hw + pix;
    const float* hp = hor + (size_t)n * Kh * hw + pix;
    for (int fy = 0; fy < Kv; ++fy) {
      const float v = __ldg(vp + (size_t)fy * hw);
      for (int fx = 0; fx < Kh; ++fx) {
        const float h = __ldg(hp + (size_t)fx * hw);
#pragma unroll
        for (int c = 0; c < CB; ++c) {
          if (c0 + c < C) {
            const float i = __ldg(in + ((size_t)n * C + c0 + c) * hwp + (size_t)(y + fy) * Wp + (x + fx));
            // Kahan summation exactly as sepconv.py:103-110
            float yk = i * v * h;
            yk = yk - comp[c];
            const float t = sum[c] + yk;
            comp[c] = (t - sum[c]) - yk;
            sum[c] = t;
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < CB; ++c)
      if (c0 + c < C) out[((size_t)n * C + c0 + c) * hw + pix] = sum[c];
  }
}

inline int grid_for(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  const size_t cap = 148 * 16;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace

cudaError_t launch_softsplat_sum(const float* in, const float* flow, float* out, int N, int C, int H, int W,
                                 cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(out, 0, (size_t)N * C * H * W * sizeof(float), st);
  if (e != cudaSuccess) return e;
  const size_t total = (size_t)N * H * W;
  softsplat_sum_kernel<<<grid_for(total, 256), 256, 0, st>>>(in, flow, out, N, C, H, W);
  return cudaGetLastError();
}

cudaError_t launch_volume81(bool dot, const float* one, const float* two, float* out, int N, int C, int H, int W,
                            cudaStream_t st) {
  const size_t total = (size_t)N * H * W;
  if (dot)
    volume81_kernel<true><<<grid_for(total, 128), 128, 0, st>>>(one, two, out, N, C, H, W);
  else
    volume81_kernel<false><<<grid_for(total, 128), 128, 0, st>>>(one, two, out, N, C, H, W);
  return cudaGetLastError();
}

cudaError_t launch_sepconv(const float* in, const float* ver, const float* hor, float* out, int N, int C, int H, int W,
                           int Kv, int Kh, cudaStream_t st) {
  const size_t total = (size_t)N * H * W;
  for (int c0 = 0; c0 < C; c0 += 4) {
    sepconv_kernel<4><<<grid_for(total, 128), 128, 0, st>>>(in, ver, hor, out, N, C, c0, H, W, Kv, Kh);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace vfi
