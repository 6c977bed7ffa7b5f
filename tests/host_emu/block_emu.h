// Cooperative emulation of ONE thread block whose threads synchronise with __syncthreads() (TEST INFRASTRUCTURE; used by
// oracle/ref_ops.py for the reference's correlation kernel).  Each emulated thread is a ucontext fiber; __syncthreads()
// yields to a round-robin scheduler, so the threads of a block advance barrier phase by barrier phase in thread order -
// the lock-step a 32-thread block (one warp) has on the GPU.  `__shared__` variables are plain statics here (one block runs
// at a time), dynamic shared memory is the buffer `emu_dyn_smem`.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <vector>

struct EmuDim3 {
  unsigned x = 1, y = 1, z = 1;
};
// thread_local: the full-geometry launcher of cuda_shim_block.h runs different blocks on different host threads
static thread_local EmuDim3 blockIdx, threadIdx, blockDim, gridDim;
static thread_local char* emu_dyn_smem = nullptr;
#define __global__
#define __shared__ static thread_local
#define __restrict__

namespace block_emu {
struct Fiber {
  ucontext_t ctx;
  std::vector<char> stack;
  bool done = false;
};
static thread_local ucontext_t sched_ctx;
static thread_local std::vector<Fiber>* fibers = nullptr;
static thread_local int current = -1;
static thread_local std::function<void()>* body = nullptr;
static void trampoline() {
  (*body)();
  (*fibers)[current].done = true;
  swapcontext(&(*fibers)[current].ctx, &sched_ctx);
}
// runs `kernel()` for every thread of one block (threadIdx.x = 0..nthreads-1), phase by phase
static void run_block(int nthreads, std::function<void()> kernel) {
  static thread_local std::vector<Fiber> fs;   // fiber stacks: allocated once per host thread, reused by every block
  if ((int)fs.size() < nthreads) fs.resize(nthreads);
  fibers = &fs;
  body = &kernel;
  for (int t = 0; t < nthreads; ++t) {
    if (fs[t].stack.empty()) fs[t].stack.resize(128 * 1024);
    fs[t].done = false;
    getcontext(&fs[t].ctx);
    fs[t].ctx.uc_stack.ss_sp = fs[t].stack.data();
    fs[t].ctx.uc_stack.ss_size = fs[t].stack.size();
    fs[t].ctx.uc_link = &sched_ctx;
    makecontext(&fs[t].ctx, trampoline, 0);
  }
  bool any = true;
  while (any) {
    any = false;
    for (int t = 0; t < nthreads; ++t) {
      if (fs[t].done) continue;
      any = true;
      current = t;
      threadIdx.x = (unsigned)t;
      swapcontext(&sched_ctx, &fs[t].ctx);
    }
  }
  fibers = nullptr;
}
}  // namespace block_emu
static inline void __syncthreads() {
  if (block_emu::fibers) {
    const int me = block_emu::current;
    swapcontext(&(*block_emu::fibers)[me].ctx, &block_emu::sched_ctx);
    threadIdx.x = (unsigned)me;
  }
}
