// Full-resolution, HBM-bound kernels of the RIFE-4.6 path (fp32 flow / mask / image arithmetic):
//   prep_frames : clamp(0,1) + zero-pad to x64 + RGB->float4           rife_arch.py:476-485
//   front       : [warp(img0), warp(img1), t, mask, flow] -> bilinear 1/s -> 16-channel block input, written
//                 directly in the space-to-depth form the stride-2 conv0.0 reads
//                                                                       rife_arch.py:31-70, :238-249, :589-596
//   upflow      : bilinear x s of the block output, flow = up*s, flow += / mask +=   rife_arch.py:263-266, :694-696
//   final       : warp both frames with the final flow, sigmoid blend, crop, clamp    rife_arch.py:703-704, :713-717, :732
//   warp        : stand-alone backward bilinear warp (border, align_corners=True), NHWC fp32, any C
#include "ptx.cuh"
#include "vfi_internal.h"

namespace vfi {
namespace {

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

__global__ void prep_frames_kernel(const float* __restrict__ frames, int n, int H, int W, int cstride,
                                   float4* __restrict__ imgs, int Hp, int Wp) {
  const size_t total = (size_t)n * Hp * Wp;
  for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(id % Wp);
    const size_t r = id / Wp;
    const int y = (int)(r % Hp);
    const int f = (int)(r / Hp);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y < H && x < W) {
      const float* s = frames + (((size_t)f * H + y) * W + x) * cstride;
      v.x = clamp01(__ldg(s));
      v.y = clamp01(__ldg(s + 1));
      v.z = clamp01(__ldg(s + 2));
    }
    imgs[id] = v;
  }
}

// grid_sample(bilinear, border, align_corners=True) at (x + fx, y + fy) of a float4 image [Hp][Wp]
__device__ __forceinline__ float4 sample_border(const float4* __restrict__ img, int Hp, int Wp, float sx, float sy) {
  sx = fminf(fmaxf(sx, 0.f), (float)(Wp - 1));
  sy = fminf(fmaxf(sy, 0.f), (float)(Hp - 1));
  const float fx0 = floorf(sx), fy0 = floorf(sy);
  const int x0 = (int)fx0, y0 = (int)fy0;
  const int x1 = min(x0 + 1, Wp - 1), y1 = min(y0 + 1, Hp - 1);
  const float ax = sx - fx0, ay = sy - fy0;
  const float w00 = (1.f - ax) * (1.f - ay), w01 = ax * (1.f - ay), w10 = (1.f - ax) * ay, w11 = ax * ay;
  const float4 a = __ldg(img + (size_t)y0 * Wp + x0);
  const float4 b = __ldg(img + (size_t)y0 * Wp + x1);
  const float4 c = __ldg(img + (size_t)y1 * Wp + x0);
  const float4 d = __ldg(img + (size_t)y1 * Wp + x1);
  float4 o;
  o.x = a.x * w00 + b.x * w01 + c.x * w10 + d.x * w11;
  o.y = a.y * w00 + b.y * w01 + c.y * w10 + d.y * w11;
  o.z = a.z * w00 + b.z * w01 + c.z * w10 + d.z * w11;
  o.w = 0.f;
  return o;
}

// one thread = one cell of the 1/s grid; channels: w0.rgb, w1.rgb, t, mask, flow/s (4) [, 4 zero pad]
template <typename T, bool kFirst>
__global__ void front_kernel(const float4* __restrict__ imgs, const float4* __restrict__ flow,
                             const float* __restrict__ mask, const BatchTasks tasks, int Hp, int Wp, int s,
                             T* __restrict__ x_s2d) {
  const int Hs = Hp / s, Ws = Wp / s;
  const size_t total = (size_t)tasks.n * Hs * Ws;
  const size_t plane = (size_t)Hp * Wp;
  const float inv_s = 1.f / (float)s;
  for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (size_t)gridDim.x * blockDim.x) {
    const int xl = (int)(id % Ws);
    const size_t r = id / Ws;
    const int yl = (int)(r % Hs);
    const int b = (int)(r / Hs);
    const float4* img0 = imgs + (size_t)tasks.f0[b] * plane;
    const float4* img1 = imgs + (size_t)tasks.f1[b] * plane;
    const float t = tasks.t[b];
    // bilinear 1/s, align_corners=False: the two source taps per axis are s*i + s/2 - 1 and s*i + s/2, weight 1/2
    const int ntap = (s == 1) ? 1 : 2;
    const int by = (s == 1) ? yl : s * yl + s / 2 - 1;
    const int bx = (s == 1) ? xl : s * xl + s / 2 - 1;
    float ch[12];
    float rowacc[2][12];
#pragma unroll
    for (int ty = 0; ty < 2; ++ty) {
      float colv[2][12];
#pragma unroll
      for (int tx = 0; tx < 2; ++tx) {
        if (ty < ntap && tx < ntap) {
          const int Y = by + ty, X = bx + tx;
          float* v = colv[tx];
          if (kFirst) {
            const float4 a = __ldg(img0 + (size_t)Y * Wp + X);
            const float4 c = __ldg(img1 + (size_t)Y * Wp + X);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = c.x; v[4] = c.y; v[5] = c.z;
            v[6] = t; v[7] = 0.f; v[8] = 0.f; v[9] = 0.f; v[10] = 0.f; v[11] = 0.f;
          } else {
            const float4 f = __ldg(flow + (size_t)b * plane + (size_t)Y * Wp + X);
            const float m = __ldg(mask + (size_t)b * plane + (size_t)Y * Wp + X);
            const float4 a = sample_border(img0, Hp, Wp, (float)X + f.x, (float)Y + f.y);
            const float4 c = sample_border(img1, Hp, Wp, (float)X + f.z, (float)Y + f.w);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = c.x; v[4] = c.y; v[5] = c.z;
            v[6] = t; v[7] = m; v[8] = f.x; v[9] = f.y; v[10] = f.z; v[11] = f.w;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 12; ++i) rowacc[ty][i] = (ntap == 1) ? colv[0][i] : (colv[0][i] + colv[1][i]);
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) ch[i] = (ntap == 1) ? rowacc[0][i] : 0.25f * (rowacc[0][i] + rowacc[1][i]);
#pragma unroll
    for (int i = 8; i < 12; ++i) ch[i] *= inv_s;  // flow is also divided by the scale (rife_arch.py:242-248)

    // space-to-depth store: cell (yl, xl) -> [b, yl/2, xl/2, ((yl&1)*2 + (xl&1))*16 + c]
    const size_t cell = ((size_t)b * (Hs >> 1) + (yl >> 1)) * (Ws >> 1) + (xl >> 1);
    T* dst = x_s2d + cell * 64 + ((yl & 1) * 2 + (xl & 1)) * 16;
    uint4 lo, hi;
    lo.x = Pack2<T>::pack(ch[0], ch[1]);
    lo.y = Pack2<T>::pack(ch[2], ch[3]);
    lo.z = Pack2<T>::pack(ch[4], ch[5]);
    lo.w = Pack2<T>::pack(ch[6], ch[7]);
    hi.x = Pack2<T>::pack(ch[8], ch[9]);
    hi.y = Pack2<T>::pack(ch[10], ch[11]);
    hi.z = 0u;
    hi.w = 0u;
    reinterpret_cast<uint4*>(dst)[0] = lo;
    reinterpret_cast<uint4*>(dst)[1] = hi;
  }
}

// flow/mask (full resolution, fp32) (+)= bilinear-upsampled block output
__global__ void upflow_kernel(const float4* __restrict__ tflow, const float* __restrict__ tmask,
                              float4* __restrict__ flow, float* __restrict__ mask, int B, int Hp, int Wp, int s,
                              int first) {
  const int Hs = Hp / s, Ws = Wp / s;
  const size_t total = (size_t)B * Hp * Wp;
  const float fs = (float)s, inv_s = 1.f / fs;
  for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(id % Wp);
    const size_t r = id / Wp;
    const int y = (int)(r % Hp);
    const int b = (int)(r / Hp);
    // F.interpolate(scale_factor=s, bilinear, align_corners=False): src = (dst + 0.5)/s - 0.5, clamped at 0
    const float sy = fmaxf(((float)y + 0.5f) * inv_s - 0.5f, 0.f);
    const float sx = fmaxf(((float)x + 0.5f) * inv_s - 0.5f, 0.f);
    const int y0 = min((int)sy, Hs - 1), x0 = min((int)sx, Ws - 1);
    const int y1 = min(y0 + 1, Hs - 1), x1 = min(x0 + 1, Ws - 1);
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const size_t base = (size_t)b * Hs * Ws;
    const float4 a = __ldg(tflow + base + (size_t)y0 * Ws + x0);
    const float4 bq = __ldg(tflow + base + (size_t)y0 * Ws + x1);
    const float4 c = __ldg(tflow + base + (size_t)y1 * Ws + x0);
    const float4 d = __ldg(tflow + base + (size_t)y1 * Ws + x1);
    const float ma = __ldg(tmask + base + (size_t)y0 * Ws + x0);
    const float mb = __ldg(tmask + base + (size_t)y0 * Ws + x1);
    const float mc = __ldg(tmask + base + (size_t)y1 * Ws + x0);
    const float md = __ldg(tmask + base + (size_t)y1 * Ws + x1);
    float4 u;
    u.x = hy * (hx * a.x + lx * bq.x) + ly * (hx * c.x + lx * d.x);
    u.y = hy * (hx * a.y + lx * bq.y) + ly * (hx * c.y + lx * d.y);
    u.z = hy * (hx * a.z + lx * bq.z) + ly * (hx * c.z + lx * d.z);
    u.w = hy * (hx * a.w + lx * bq.w) + ly * (hx * c.w + lx * d.w);
    const float um = hy * (hx * ma + lx * mb) + ly * (hx * mc + lx * md);
    float4 f = make_float4(u.x * fs, u.y * fs, u.z * fs, u.w * fs);
    float m = um;
    if (!first) {
      const float4 f0 = flow[id];
      f.x += f0.x; f.y += f0.y; f.z += f0.z; f.w += f0.w;
      m += mask[id];
    }
    flow[id] = f;
    mask[id] = m;
  }
}

__global__ void final_kernel(const float4* __restrict__ imgs, const float4* __restrict__ flow,
                             const float* __restrict__ mask, const BatchTasks tasks, int Hp, int Wp, int H, int W,
                             float* __restrict__ out) {
  const size_t total = (size_t)tasks.n * H * W;
  const size_t plane = (size_t)Hp * Wp;
  for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(id % W);
    const size_t r = id / W;
    const int y = (int)(r % H);
    const int b = (int)(r / H);
    const size_t pid = (size_t)b * plane + (size_t)y * Wp + x;
    const float4 f = __ldg(flow + pid);
    const float m = __ldg(mask + pid);
    const float4 a = sample_border(imgs + (size_t)tasks.f0[b] * plane, Hp, Wp, (float)x + f.x, (float)y + f.y);
    const float4 c = sample_border(imgs + (size_t)tasks.f1[b] * plane, Hp, Wp, (float)x + f.z, (float)y + f.w);
    const float sg = 1.f / (1.f + expf(-m));
    float* o = out + id * 3;
    o[0] = clamp01(a.x * sg + c.x * (1.f - sg));
    o[1] = clamp01(a.y * sg + c.y * (1.f - sg));
    o[2] = clamp01(a.z * sg + c.z * (1.f - sg));
  }
}

// stand-alone warp, NHWC fp32 with C channels (C % 4 == 0 uses 128-bit loads): img [B,H,W,C], flow [B,H,W,2]
template <int VEC>
__global__ void warp_kernel(const float* __restrict__ img, const float* __restrict__ flow, float* __restrict__ out,
                            int B, int H, int W, int C) {
  const int cv = C / VEC;  // vectors per pixel
  const size_t total = (size_t)B * H * W * cv;
  for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (size_t)gridDim.x * blockDim.x) {
    const int v = (int)(id % cv);
    const size_t pix = id / cv;
    const int x = (int)(pix % W);
    const size_t r = pix / W;
    const int y = (int)(r % H);
    const int b = (int)(r / H);
    const float2 f = __ldg(reinterpret_cast<const float2*>(flow) + pix);
    const float sx = fminf(fmaxf((float)x + f.x, 0.f), (float)(W - 1));
    const float sy = fminf(fmaxf((float)y + f.y, 0.f), (float)(H - 1));
    const float fx0 = floorf(sx), fy0 = floorf(sy);
    const int x0 = (int)fx0, y0 = (int)fy0;
    const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
    const float ax = sx - fx0, ay = sy - fy0;
    const float w00 = (1.f - ax) * (1.f - ay), w01 = ax * (1.f - ay), w10 = (1.f - ax) * ay, w11 = ax * ay;
    const float* base = img + (size_t)b * H * W * C + (size_t)v * VEC;
    const float* p00 = base + ((size_t)y0 * W + x0) * C;
    const float* p01 = base + ((size_t)y0 * W + x1) * C;
    const float* p10 = base + ((size_t)y1 * W + x0) * C;
    const float* p11 = base + ((size_t)y1 * W + x1) * C;
    float* o = out + pix * C + (size_t)v * VEC;
    if (VEC == 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p00));
      const float4 bq = __ldg(reinterpret_cast<const float4*>(p01));
      const float4 c = __ldg(reinterpret_cast<const float4*>(p10));
      const float4 d = __ldg(reinterpret_cast<const float4*>(p11));
      float4 q;
      q.x = a.x * w00 + bq.x * w01 + c.x * w10 + d.x * w11;
      q.y = a.y * w00 + bq.y * w01 + c.y * w10 + d.y * w11;
      q.z = a.z * w00 + bq.z * w01 + c.z * w10 + d.z * w11;
      q.w = a.w * w00 + bq.w * w01 + c.w * w10 + d.w * w11;
      *reinterpret_cast<float4*>(o) = q;
    } else {
      o[0] = __ldg(p00) * w00 + __ldg(p01) * w01 + __ldg(p10) * w10 + __ldg(p11) * w11;
    }
  }
}

inline int grid_for(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  const size_t cap = 148 * 32;  // grid-stride loops: a few waves of the 148 SMs
  return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace

cudaError_t launch_prep_frames(const float* frames, int n, int H, int W, int cstride, float4* imgs, int Hp, int Wp,
                               cudaStream_t st) {
  const size_t total = (size_t)n * Hp * Wp;
  prep_frames_kernel<<<grid_for(total, 256), 256, 0, st>>>(frames, n, H, W, cstride, imgs, Hp, Wp);
  return cudaGetLastError();
}

cudaError_t launch_front(int op_type, const float4* imgs, const float4* flow, const float* mask, BatchTasks tasks,
                         int Hp, int Wp, int s, bool first, void* x_s2d, cudaStream_t st) {
  const size_t total = (size_t)tasks.n * (Hp / s) * (Wp / s);
  const int g = grid_for(total, 128);
  if (op_type == OP_BF16) {
    if (first)
      front_kernel<__nv_bfloat16, true><<<g, 128, 0, st>>>(imgs, flow, mask, tasks, Hp, Wp, s, (__nv_bfloat16*)x_s2d);
    else
      front_kernel<__nv_bfloat16, false><<<g, 128, 0, st>>>(imgs, flow, mask, tasks, Hp, Wp, s, (__nv_bfloat16*)x_s2d);
  } else {
    if (first)
      front_kernel<__half, true><<<g, 128, 0, st>>>(imgs, flow, mask, tasks, Hp, Wp, s, (__half*)x_s2d);
    else
      front_kernel<__half, false><<<g, 128, 0, st>>>(imgs, flow, mask, tasks, Hp, Wp, s, (__half*)x_s2d);
  }
  return cudaGetLastError();
}

cudaError_t launch_upflow(const float4* tmp_flow, const float* tmp_mask, float4* flow, float* mask, int B, int Hp,
                          int Wp, int s, bool first, cudaStream_t st) {
  const size_t total = (size_t)B * Hp * Wp;
  upflow_kernel<<<grid_for(total, 256), 256, 0, st>>>(tmp_flow, tmp_mask, flow, mask, B, Hp, Wp, s, first ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t launch_final(const float4* imgs, const float4* flow, const float* mask, BatchTasks tasks, int Hp, int Wp,
                         int H, int W, float* out, cudaStream_t st) {
  const size_t total = (size_t)tasks.n * H * W;
  final_kernel<<<grid_for(total, 256), 256, 0, st>>>(imgs, flow, mask, tasks, Hp, Wp, H, W, out);
  return cudaGetLastError();
}

cudaError_t launch_warp(const float* img, const float* flow, float* out, int B, int H, int W, int C, cudaStream_t st) {
  if (C % 4 == 0) {
    const size_t total = (size_t)B * H * W * (C / 4);
    warp_kernel<4><<<grid_for(total, 256), 256, 0, st>>>(img, flow, out, B, H, W, C);
  } else {
    const size_t total = (size_t)B * H * W * C;
    warp_kernel<1><<<grid_for(total, 256), 256, 0, st>>>(img, flow, out, B, H, W, C);
  }
  return cudaGetLastError();
}

}  // namespace vfi
