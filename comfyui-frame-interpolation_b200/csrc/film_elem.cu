// HBM-bound kernels of the FILM path (everything of film_arch.Interpolator.forward that is not a >= 64-channel conv).
// Activations are 16-bit NHWC channel slices {pointer, pixel pitch}; images and flows are fp32 NHWC.
//
//   gather_rgb   frames[idx] -> [n,H,W,3]                      preprocess_frames vfi_utils.py:139-140 ([..., :3])
//   pool_rgb     2x2 mean of an fp32 RGB image                 build_image_pyramid film_arch.py:655-674
//   conv_rgb     Conv2d(3,64,3,'same') + LeakyReLU(0.2)        SubTreeExtractor convs[0][0] film_arch.py:94-98
//   pool16       2x2 mean of a 16-bit feature slice            SubTreeExtractor.forward film_arch.py:118-119
//   flow_up      bilinear resize of 2*v to the next level      film_arch.py:598, :611, :751 (align_corners=False)
//   warp16       backward warp of a feature slice by a flow    warp film_arch.py:677-724 (border, pixel offsets)
//   misc64       warped images + scaled flows, one 64-ch block Interpolator.debug_forward film_arch.py:425-447
//   nearest16    nearest-neighbour resize of a feature tensor  Fusion.forward film_arch.py:290
//   flow_head    Conv2d(c,2,1) + residual flow accumulation    FlowEstimator._convs[-1] :527, :602-603, :615-616
//   out_rgb      Conv2d(64,3,1) [+ clamp(0,1)]                 Fusion.output_conv :230, :295; film/__init__.py:38
#include "ptx.cuh"
#include "vfi_internal.h"

namespace vfi {

namespace {

template <typename T>
struct Vec8 {  // eight 16-bit values = one 16-byte access
  uint4 u;
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 p = Pack2<T>::unpack(w[i]);
      f[2 * i] = p.x;
      f[2 * i + 1] = p.y;
    }
  }
  __device__ __forceinline__ void pack(const float (&f)[8]) {
    u.x = Pack2<T>::pack(f[0], f[1]);
    u.y = Pack2<T>::pack(f[2], f[3]);
    u.z = Pack2<T>::pack(f[4], f[5]);
    u.w = Pack2<T>::pack(f[6], f[7]);
  }
};

inline int grid_for(size_t total, int threads) {
  size_t b = (total + threads - 1) / threads;
  return (int)(b < 1 ? 1 : (b > 1048576 ? 1048576 : b));
}

__global__ void gather_rgb_kernel(const float* __restrict__ frames, int cstride, const FilmFrameIdx idx, int n, size_t hw,
                                  float* __restrict__ out) {
  const size_t total = (size_t)n * hw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int img = (int)(i / hw);
    const size_t px = i - (size_t)img * hw;
    const float* s = frames + ((size_t)idx.i[img] * hw + px) * cstride;
    float* d = out + i * 3;
    d[0] = s[0];
    d[1] = s[1];
    d[2] = s[2];
  }
}

// avg_pool2d(2, 2): output floor(H/2) x floor(W/2); (a + b + c + d) * 0.25 in ATen's order of accumulation (row-major window)
__global__ void pool_rgb_kernel(const float* __restrict__ in, float* __restrict__ out, int n, int H, int W) {
  const int Ho = H >> 1, Wo = W >> 1;
  const size_t total = (size_t)n * Ho * Wo;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wo);
    const int y = (int)((i / Wo) % Ho);
    const int b = (int)(i / ((size_t)Wo * Ho));
    const float* p = in + (((size_t)b * H + 2 * y) * W + 2 * x) * 3;
    const float* q = p + (size_t)W * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) out[i * 3 + c] = (p[c] + p[3 + c] + q[c] + q[3 + c]) * 0.25f;
  }
}

// Conv2d(3, 64, 3, padding='same') + LeakyReLU(0.2) on the CUDA cores (K = 27: no tensor-core shape); two threads per
// pixel, 32 output channels each, weights [64][3][3][3] (PyTorch order) transposed in shared memory to [27][64].
template <typename T>
__global__ void conv_rgb_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                                T* __restrict__ out, int out_pitch, int n, int H, int W) {
  __shared__ float ws[27 * 64];
  __shared__ float bs[64];
  for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) {
    const int k = i / 64, o = i - k * 64;  // k = (c, ky, kx) flattened as in the PyTorch weight
    ws[i] = w[o * 27 + k];
  }
  for (int i = threadIdx.x; i < 64; i += blockDim.x) bs[i] = bias[i];
  __syncthreads();
  const size_t total = (size_t)n * H * W * 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int half = (int)(i & 1);
    const size_t px = i >> 1;
    const int x = (int)(px % W);
    const int y = (int)((px / W) % H);
    const int b = (int)(px / ((size_t)W * H));
    float v[27];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int yy = y + ky - 1, xx = x + kx - 1;
        const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
        const float* s = img + (((size_t)b * H + (ok ? yy : 0)) * W + (ok ? xx : 0)) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c * 9 + ky * 3 + kx] = ok ? s[c] : 0.f;
      }
    T* o = out + px * (size_t)out_pitch + half * 32;
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = bs[half * 32 + c8 * 8 + j];
#pragma unroll
      for (int k = 0; k < 27; ++k) {
        const float* wr = ws + k * 64 + half * 32 + c8 * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(v[k], wr[j], acc[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = lrelu02(acc[j]);
      Vec8<T> pk;
      pk.pack(acc);
      *reinterpret_cast<uint4*>(o + c8 * 8) = pk.u;
    }
  }
}

template <typename T>
__global__ void pool16_kernel(const T* __restrict__ in, int in_pitch, T* __restrict__ out, int out_pitch, int C8, int n,
                              int H, int W) {
  const int Ho = H >> 1, Wo = W >> 1;
  const size_t total = (size_t)n * Ho * Wo * C8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    size_t px = i / C8;
    const int x = (int)(px % Wo);
    const int y = (int)((px / Wo) % Ho);
    const int b = (int)(px / ((size_t)Wo * Ho));
    const T* p = in + (((size_t)b * H + 2 * y) * W + 2 * x) * (size_t)in_pitch + c * 8;
    float a[8], s[8];
    Vec8<T> v;
    v.u = *reinterpret_cast<const uint4*>(p);
    v.unpack(s);
    v.u = *reinterpret_cast<const uint4*>(p + in_pitch);
    v.unpack(a);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += a[j];
    v.u = *reinterpret_cast<const uint4*>(p + (size_t)W * in_pitch);
    v.unpack(a);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += a[j];
    v.u = *reinterpret_cast<const uint4*>(p + (size_t)W * in_pitch + in_pitch);
    v.unpack(a);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = (s[j] + a[j]) * 0.25f;
    v.pack(s);
    *reinterpret_cast<uint4*>(out + px * (size_t)out_pitch + c * 8) = v.u;
  }
}

// F.interpolate(2 * v, size=(H, W), mode='bilinear') (align_corners=False): ATen's source index is
// max(scale * (dst + 0.5) - 0.5, 0) with scale = in / out (float), the second tap clamped to the last row / column.
__global__ void flow_up_kernel(const float2* __restrict__ v, int h, int w, float2* __restrict__ out, int B, int H, int W) {
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  const size_t total = (size_t)B * H * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int b = (int)(i / ((size_t)W * H));
    const float fy = fmaxf(sy * ((float)y + 0.5f) - 0.5f, 0.f), fx = fmaxf(sx * ((float)x + 0.5f) - 0.5f, 0.f);
    const int y0 = min((int)fy, h - 1), x0 = min((int)fx, w - 1);
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float2* p = v + (size_t)b * h * w;
    const float2 a = p[(size_t)y0 * w + x0], c = p[(size_t)y0 * w + x1];
    const float2 d = p[(size_t)y1 * w + x0], e = p[(size_t)y1 * w + x1];
    float2 o;
    o.x = (1.f - ly) * ((1.f - lx) * (2.f * a.x) + lx * (2.f * c.x)) + ly * ((1.f - lx) * (2.f * d.x) + lx * (2.f * e.x));
    o.y = (1.f - ly) * ((1.f - lx) * (2.f * a.y) + lx * (2.f * c.y)) + ly * ((1.f - lx) * (2.f * d.y) + lx * (2.f * e.y));
    out[i] = o;
  }
}

// bilinear sample position of film_arch.warp: pixel (x + fx, y + fy), coordinates clamped to the image (grid_sample
// padding_mode='border'); a tap beyond the last row / column has weight 0 there.
struct Taps {
  int x0, x1, y0, y1;
  float wx, wy;
};
__device__ __forceinline__ Taps film_taps(int x, int y, float fx, float fy, int W, int H) {
  Taps t;
  const float sx = fminf(fmaxf((float)x + fx, 0.f), (float)(W - 1));
  const float sy = fminf(fmaxf((float)y + fy, 0.f), (float)(H - 1));
  const float flx = floorf(sx), fly = floorf(sy);
  t.x0 = (int)flx;
  t.y0 = (int)fly;
  t.x1 = min(t.x0 + 1, W - 1);
  t.y1 = min(t.y0 + 1, H - 1);
  t.wx = sx - flx;
  t.wy = sy - fly;
  return t;
}

template <typename T>
__global__ void warp16_kernel(const T* __restrict__ src, int src_pitch, int C8, const float2* __restrict__ flow,
                              float fscale, T* __restrict__ dst, int dst_pitch, int B, int H, int W) {
  const size_t total = (size_t)B * H * W * C8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    const size_t px = i / C8;
    const int x = (int)(px % W);
    const int y = (int)((px / W) % H);
    const int b = (int)(px / ((size_t)W * H));
    const float2 f = flow[px];
    const Taps t = film_taps(x, y, f.x * fscale, f.y * fscale, W, H);
    const T* base = src + (size_t)b * H * W * src_pitch + c * 8;
    float a[8], o[8];
    Vec8<T> v;
    const float w00 = (1.f - t.wx) * (1.f - t.wy), w01 = t.wx * (1.f - t.wy), w10 = (1.f - t.wx) * t.wy, w11 = t.wx * t.wy;
    v.u = *reinterpret_cast<const uint4*>(base + ((size_t)t.y0 * W + t.x0) * src_pitch);
    v.unpack(a);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = a[j] * w00;
    v.u = *reinterpret_cast<const uint4*>(base + ((size_t)t.y0 * W + t.x1) * src_pitch);
    v.unpack(a);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(a[j], w01, o[j]);
    v.u = *reinterpret_cast<const uint4*>(base + ((size_t)t.y1 * W + t.x0) * src_pitch);
    v.unpack(a);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(a[j], w10, o[j]);
    v.u = *reinterpret_cast<const uint4*>(base + ((size_t)t.y1 * W + t.x1) * src_pitch);
    v.unpack(a);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(a[j], w11, o[j]);
    v.pack(o);
    *reinterpret_cast<uint4*>(dst + px * (size_t)dst_pitch + c * 8) = v.u;
  }
}

__device__ __forceinline__ void sample_rgb(const float* img, const Taps& t, int W, float (&o)[3]) {
  const float w00 = (1.f - t.wx) * (1.f - t.wy), w01 = t.wx * (1.f - t.wy), w10 = (1.f - t.wx) * t.wy, w11 = t.wx * t.wy;
  const float* p00 = img + ((size_t)t.y0 * W + t.x0) * 3;
  const float* p01 = img + ((size_t)t.y0 * W + t.x1) * 3;
  const float* p10 = img + ((size_t)t.y1 * W + t.x0) * 3;
  const float* p11 = img + ((size_t)t.y1 * W + t.x1) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c] = p00[c] * w00 + p01[c] * w01 + p10[c] * w10 + p11[c] * w11;
}

// The 64-channel block of an aligned-pyramid level that is not warped features: channels
//   0-2 warp(img0, 0.5 * bwd flow), 3-5 warp(img1, 0.5 * fwd flow), 6-7 0.5 * bwd flow, 8-9 0.5 * fwd flow, 10-63 zero
// (film_arch.py:425-447: backward_flow = bwd * 0.5 reads image 0, forward_flow = fwd * (1 - 0.5) reads image 1).
template <typename T>
__global__ void misc64_kernel(const float* __restrict__ img0, const float* __restrict__ img1,
                              const float2* __restrict__ bflow, const float2* __restrict__ fflow, T* __restrict__ dst,
                              int dst_pitch, int B, int H, int W) {
  const size_t total = (size_t)B * H * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int b = (int)(i / ((size_t)W * H));
    float2 fb = bflow[i], ff = fflow[i];
    fb.x *= 0.5f;
    fb.y *= 0.5f;
    ff.x *= 0.5f;
    ff.y *= 0.5f;
    float a[3], c[3];
    sample_rgb(img0 + (size_t)b * H * W * 3, film_taps(x, y, fb.x, fb.y, W, H), W, a);
    sample_rgb(img1 + (size_t)b * H * W * 3, film_taps(x, y, ff.x, ff.y, W, H), W, c);
    const float v0[8] = {a[0], a[1], a[2], c[0], c[1], c[2], fb.x, fb.y};
    const float v1[8] = {ff.x, ff.y, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    Vec8<T> p0, p1;
    p0.pack(v0);
    p1.pack(v1);
    uint4* d = reinterpret_cast<uint4*>(dst + i * (size_t)dst_pitch);
    d[0] = p0.u;
    d[1] = p1.u;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int k = 2; k < 8; ++k) d[k] = z;
  }
}

// F.interpolate(mode='nearest', size=(H, W)): ATen nearest_idx - identity when sizes match, dst >> 1 for an exact
// doubling, otherwise min(floor(dst * (float)in / out), in - 1).
__device__ __forceinline__ int nearest_src(int dst, int in, int out, float scale) {
  if (out == in) return dst;
  if (out == 2 * in) return dst >> 1;
  return min((int)floorf((float)dst * scale), in - 1);
}
__global__ void nearest16_kernel(const uint4* __restrict__ in, int h, int w, uint4* __restrict__ out, int C8, int B, int H,
                                 int W) {
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  const size_t total = (size_t)B * H * W * C8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    const size_t px = i / C8;
    const int x = (int)(px % W);
    const int y = (int)((px / W) % H);
    const int b = (int)(px / ((size_t)W * H));
    const int ys = nearest_src(y, h, H, sy), xs = nearest_src(x, w, W, sx);
    out[i] = in[(((size_t)b * h + ys) * w + xs) * C8 + c];
  }
}

// Conv2d(C, 2, 1) without activation on the CUDA cores + the flow recurrence v = residual + up(2 * v_coarser)
// (film_arch.py:602-603, :615-616; flow_pyramid_synthesis :745-755 rebuilds exactly these v).
template <typename T>
__global__ void flow_head_kernel(const T* __restrict__ x, int pitch, int C, const float* __restrict__ w,
                                 const float* __restrict__ bias, const float2* __restrict__ v_up, float2* __restrict__ v_out,
                                 size_t npx) {
  __shared__ float ws[2 * 128];
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const float b0 = bias[0], b1 = bias[1];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (size_t)gridDim.x * blockDim.x) {
    const T* p = x + i * (size_t)pitch;
    float a0 = 0.f, a1 = 0.f;
    for (int c = 0; c < C; c += 8) {
      Vec8<T> v;
      v.u = *reinterpret_cast<const uint4*>(p + c);
      float f[8];
      v.unpack(f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a0 = fmaf(f[j], ws[c + j], a0);
        a1 = fmaf(f[j], ws[C + c + j], a1);
      }
    }
    float2 o = make_float2(a0 + b0, a1 + b1);
    if (v_up) {
      const float2 u = v_up[i];
      o.x += u.x;
      o.y += u.y;
    }
    v_out[i] = o;
  }
}

template <typename T>
__global__ void out_rgb_kernel(const T* __restrict__ x, int pitch, const float* __restrict__ w,
                               const float* __restrict__ bias, int clamp01, float* __restrict__ out, size_t npx) {
  __shared__ float ws[3 * 64];
  for (int i = threadIdx.x; i < 3 * 64; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (size_t)gridDim.x * blockDim.x) {
    const T* p = x + i * (size_t)pitch;
    float a[3] = {bias[0], bias[1], bias[2]};
#pragma unroll
    for (int c = 0; c < 64; c += 8) {
      Vec8<T> v;
      v.u = *reinterpret_cast<const uint4*>(p + c);
      float f[8];
      v.unpack(f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a[0] = fmaf(f[j], ws[c + j], a[0]);
        a[1] = fmaf(f[j], ws[64 + c + j], a[1]);
        a[2] = fmaf(f[j], ws[128 + c + j], a[2]);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) out[i * 3 + c] = clamp01 ? fminf(fmaxf(a[c], 0.f), 1.f) : a[c];
  }
}

}  // namespace

#define VFI_T(op, expr_h, expr_b) ((op) == OP_BF16 ? (expr_b) : (expr_h))

cudaError_t launch_film_gather_rgb(const float* frames, int cstride, const FilmFrameIdx& idx, int n, int H, int W,
                                   float* out, cudaStream_t st) {
  const size_t hw = (size_t)H * W;
  VFI_LAUNCH(gather_rgb_kernel, grid_for((size_t)n * hw, 256), 256, 0, st, frames, cstride, idx, n, hw, out);
  return cudaGetLastError();
}

cudaError_t launch_film_pool_rgb(const float* in, float* out, int n, int H, int W, cudaStream_t st) {
  if ((H >> 1) < 1 || (W >> 1) < 1) return cudaErrorInvalidValue;
  VFI_LAUNCH(pool_rgb_kernel, grid_for((size_t)n * (H >> 1) * (W >> 1), 256), 256, 0, st, in, out, n, H, W);
  return cudaGetLastError();
}

cudaError_t launch_film_conv_rgb(int op_type, const float* img, const float* w, const float* bias, void* out,
                                 int out_pitch, int n, int H, int W, cudaStream_t st) {
  const int g = grid_for((size_t)n * H * W * 2, 256);
  if (op_type == OP_BF16)
    VFI_LAUNCH(conv_rgb_kernel<__nv_bfloat16>, g, 256, 0, st, img, w, bias, (__nv_bfloat16*)out, out_pitch, n, H, W);
  else
    VFI_LAUNCH(conv_rgb_kernel<__half>, g, 256, 0, st, img, w, bias, (__half*)out, out_pitch, n, H, W);
  return cudaGetLastError();
}

cudaError_t launch_film_pool16(int op_type, const void* in, int in_pitch, void* out, int out_pitch, int C, int n, int H,
                               int W, cudaStream_t st) {
  if ((H >> 1) < 1 || (W >> 1) < 1 || (C & 7)) return cudaErrorInvalidValue;
  const int g = grid_for((size_t)n * (H >> 1) * (W >> 1) * (C / 8), 256);
  if (op_type == OP_BF16)
    VFI_LAUNCH(pool16_kernel<__nv_bfloat16>, g, 256, 0, st, (const __nv_bfloat16*)in, in_pitch, (__nv_bfloat16*)out, out_pitch,
                                                    C / 8, n, H, W);
  else
    VFI_LAUNCH(pool16_kernel<__half>, g, 256, 0, st, (const __half*)in, in_pitch, (__half*)out, out_pitch, C / 8, n, H, W);
  return cudaGetLastError();
}

cudaError_t launch_film_flow_up(const float* v, int h, int w, float* out, int B, int H, int W, cudaStream_t st) {
  VFI_LAUNCH(flow_up_kernel, grid_for((size_t)B * H * W, 256), 256, 0, st, (const float2*)v, h, w, (float2*)out, B, H, W);
  return cudaGetLastError();
}

cudaError_t launch_film_warp16(int op_type, const void* src, int src_pitch, int C, const float* flow, float fscale,
                               void* dst, int dst_pitch, int B, int H, int W, cudaStream_t st) {
  if (C & 7) return cudaErrorInvalidValue;
  const int g = grid_for((size_t)B * H * W * (C / 8), 256);
  if (op_type == OP_BF16)
    VFI_LAUNCH(warp16_kernel<__nv_bfloat16>, g, 256, 0, st, (const __nv_bfloat16*)src, src_pitch, C / 8, (const float2*)flow,
                                                    fscale, (__nv_bfloat16*)dst, dst_pitch, B, H, W);
  else
    VFI_LAUNCH(warp16_kernel<__half>, g, 256, 0, st, (const __half*)src, src_pitch, C / 8, (const float2*)flow, fscale,
                                             (__half*)dst, dst_pitch, B, H, W);
  return cudaGetLastError();
}

cudaError_t launch_film_misc64(int op_type, const float* img0, const float* img1, const float* bflow,
                               const float* fflow, void* dst, int dst_pitch, int B, int H, int W, cudaStream_t st) {
  const int g = grid_for((size_t)B * H * W, 256);
  if (op_type == OP_BF16)
    VFI_LAUNCH(misc64_kernel<__nv_bfloat16>, g, 256, 0, st, img0, img1, (const float2*)bflow, (const float2*)fflow,
                                                    (__nv_bfloat16*)dst, dst_pitch, B, H, W);
  else
    VFI_LAUNCH(misc64_kernel<__half>, g, 256, 0, st, img0, img1, (const float2*)bflow, (const float2*)fflow, (__half*)dst,
                                             dst_pitch, B, H, W);
  return cudaGetLastError();
}

cudaError_t launch_film_nearest16(const void* in, int h, int w, void* out, int C, int B, int H, int W, cudaStream_t st) {
  if (C & 7) return cudaErrorInvalidValue;
  VFI_LAUNCH(nearest16_kernel, grid_for((size_t)B * H * W * (C / 8), 256), 256, 0, st, (const uint4*)in, h, w, (uint4*)out, C / 8,
                                                                               B, H, W);
  return cudaGetLastError();
}

cudaError_t launch_film_flow_head(int op_type, const void* x, int pitch, int C, const float* w, const float* bias,
                                  const float* v_up, float* v_out, int B, int H, int W, cudaStream_t st) {
  if (C > 128 || (C & 7)) return cudaErrorInvalidValue;
  const size_t npx = (size_t)B * H * W;
  const int g = grid_for(npx, 128);
  if (op_type == OP_BF16)
    VFI_LAUNCH(flow_head_kernel<__nv_bfloat16>, g, 128, 0, st, (const __nv_bfloat16*)x, pitch, C, w, bias, (const float2*)v_up,
                                                       (float2*)v_out, npx);
  else
    VFI_LAUNCH(flow_head_kernel<__half>, g, 128, 0, st, (const __half*)x, pitch, C, w, bias, (const float2*)v_up, (float2*)v_out,
                                                npx);
  return cudaGetLastError();
}

cudaError_t launch_film_out_rgb(int op_type, const void* x, int pitch, const float* w, const float* bias, int clamp01,
                                float* out, int B, int H, int W, cudaStream_t st) {
  const size_t npx = (size_t)B * H * W;
  const int g = grid_for(npx, 128);
  if (op_type == OP_BF16)
    VFI_LAUNCH(out_rgb_kernel<__nv_bfloat16>, g, 128, 0, st, (const __nv_bfloat16*)x, pitch, w, bias, clamp01, out, npx);
  else
    VFI_LAUNCH(out_rgb_kernel<__half>, g, 128, 0, st, (const __half*)x, pitch, w, bias, clamp01, out, npx);
  return cudaGetLastError();
}

}  // namespace vfi
