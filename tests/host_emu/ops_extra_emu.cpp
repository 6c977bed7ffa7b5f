// Host build of csrc/ops_extra.cu (AdaCoF, distance-transform pass) through cuda_shim.h, C entry points for ctypes.
#include "cuda_shim.h"

#include <string>

#include "../../comfyui-frame-interpolation_b200/csrc/vfi_internal.h"
namespace vfi {
void set_error(const std::string&) {}
}
#include "../../comfyui-frame-interpolation_b200/csrc/ops_extra.cu"

extern "C" {
int emu_adacof(const float* in, const float* w, const float* oi, const float* oj, float* out, int N, int C, int Hin, int Win,
               int F, int dil, int Ho, int Wo) {
  return (int)vfi::launch_adacof(in, w, oi, oj, out, N, C, Hin, Win, F, dil, Ho, Wo, nullptr);
}
int emu_edt_pass(const float* data, float* out, int bs, int h, int w, float diam2) {
  return (int)vfi::launch_edt_pass(data, out, bs, h, w, diam2, nullptr);
}
}
