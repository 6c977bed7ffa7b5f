// Whole-path host emulation of the Sepconv trunk (TEST INFRASTRUCTURE): csrc/sepconv.cu (weight loading, the forward
// schedule, buffer sizing, the C ABI), csrc/sepconv_elem.cu and csrc/streamconv.cu's packer + CUDA-core checker kernel,
// compiled for the host through cuda_shim.h.  What is NOT emulated: the tcgen05 kernel (the checker stands in, exactly as
// `vfi_sepconv_debug_set_ref(ctx, 1)` selects on a GPU) and ops.cu's tiled separable-convolution kernel (restated below).
#include "cuda_shim.h"

#include <string>

#include "../../comfyui-frame-interpolation_b200/csrc/vfi_internal.h"

struct vfi_ctx {
  vfi::SepState* sep = nullptr;
  int launches = 0;
};
namespace vfi {
static std::string g_err;
void set_error(const std::string& s) { g_err = s; }
CtxInfo ctx_info(::vfi_ctx*) { return CtxInfo{0, 148, nullptr, nullptr, nullptr}; }
SepState*& ctx_sep(::vfi_ctx* c) { return c->sep; }
void ctx_add_launches(::vfi_ctx* c, int n) { c->launches += n; }
// sepconv_out, cupy_ops/sepconv.py:86-117 (the role of ops.cu's kernel in this emulation)
cudaError_t launch_sepconv(const float* in, const float* ver, const float* hor, float* out, int N, int C, int H, int W, int Kv,
                           int Kh, cudaStream_t) {
  const int Hp = H + Kv - 1, Wp = W + Kh - 1;
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c)
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
          double acc = 0.0;
          for (int fy = 0; fy < Kv; ++fy) {
            const float v = ver[(((size_t)n * Kv + fy) * H + y) * W + x];
            const float* row = in + (((size_t)n * C + c) * Hp + y + fy) * Wp + x;
            double r = 0.0;
            for (int fx = 0; fx < Kh; ++fx) r += (double)row[fx] * hor[(((size_t)n * Kh + fx) * H + y) * W + x];
            acc += r * v;
          }
          out[(((size_t)n * C + c) * H + y) * W + x] = (float)acc;
        }
  return cudaSuccess;
}
}  // namespace vfi
extern "C" const char* vfi_last_error(void) { return vfi::g_err.c_str(); }

#include "../../comfyui-frame-interpolation_b200/csrc/sepconv_elem.cu"
#include "../../comfyui-frame-interpolation_b200/csrc/streamconv.cu"
#include "../../comfyui-frame-interpolation_b200/csrc/sepconv.cu"

extern "C" {
// weights in, one pair in, one frame out - everything through the product's own vfi_sepconv_load / vfi_sepconv_forward
// coef (optional): [4][51][He][We] fp32 = the coefficient planes of the heads Verone, Vertwo, Horone, Hortwo
int emu_sepconv(const float* const* tensors, const int64_t* numel, int n_tensors, const float* frames, int H, int W, int C,
                float* out, float* coef) {
  vfi_ctx ctx;
  int rc = vfi_sepconv_load(&ctx, tensors, numel, n_tensors, VFI_OPERAND_F16);
  if (rc) return rc;
  rc = vfi_sepconv_debug_set_ref(&ctx, 1);
  if (rc) return rc;
  const int32_t f0[1] = {0}, f1[1] = {1};
  rc = vfi_sepconv_forward(&ctx, frames, 2, H, W, C, f0, f1, 1, out, nullptr);
  if (!rc && coef) {
    const size_t n = (size_t)51 * (H + (H & 1)) * (W + (W & 1));
    for (int k = 0; k < 4; ++k) std::memcpy(coef + k * n, ctx.sep->coef[k].p, n * sizeof(float));
  }
  vfi::sepconv_destroy(ctx.sep);
  return rc ? rc : 1000 + ctx.launches;   // 1000 + number of launches on success
}
}
