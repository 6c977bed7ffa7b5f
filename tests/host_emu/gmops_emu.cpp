// Host build of csrc/gmops.cu (the fp32 building blocks of the GMFSS path) through cuda_shim_block.h: every block on a host
// thread, its threads as fibers; exports the same vfi_gm_* C ABI, so comfyui-frame-interpolation_b200/gmfss.py runs its whole
// schedule on CPU tensors against this library (tests/test_gmfss_host.py).  TEST INFRASTRUCTURE.
#include "cuda_shim_block.h"

#include <string>

#include "../../include/vfi_b200.h"
#include "../../comfyui-frame-interpolation_b200/csrc/vfi_internal.h"
namespace vfi {
static thread_local std::string g_gm_err;
void set_error(const std::string& s) { g_gm_err = s; }
}
extern "C" const char* vfi_last_error(void) { return vfi::g_gm_err.c_str(); }
#include "../../comfyui-frame-interpolation_b200/csrc/gmops.cu"
