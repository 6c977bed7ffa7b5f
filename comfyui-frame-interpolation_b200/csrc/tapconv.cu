// tapconv: tap-list implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 / TMEM), sm_100a only.
//
// Replaces, for the RIFE IFBlock (reference rife_arch.py:237-276), the cuDNN/ATen kernels behind
//   conv0.0 / conv0.1  Conv2d(3x3, stride 2)+LeakyReLU        rife_arch.py:96-107, :181-184
//   convblock          8 x ResConv  lrelu(conv3x3(x)*beta + x) rife_arch.py:20-28, :201-210
//   lastconv           ConvTranspose2d(c,24,4,2,1)+PixelShuffle rife_arch.py:215-218
//
// One CTA tile = 16 x 8 grid cells = the M=128 rows of one tcgen05.mma.  The input window of the tile
// (tile + halo, all input channels) is staged ONCE in shared memory as 8-channel planes
//      A[chunk][halo_pixel][8 ch]          (16 bytes per pixel per plane)
// which is exactly the K-major / no-swizzle UMMA operand layout (core matrix = 8 pixels x 16 B = 128 contiguous
// bytes).  Every filter tap is then just a different descriptor START ADDRESS into the same staged window
// (shift by (dy*halo_w + dx) pixels = 16-byte units), so the 3x3 window is read from L2 once (x1.4 halo), not 9x.
// Weights of the CTA's output-channel slice stay resident in shared memory for the whole launch
// (one cp.async.bulk per CTA), accumulators live in TMEM (double buffered), the epilogue reads the residual
// from the staged window, not from global memory.
//
// Warp roles (384 threads, 1 CTA/SM, persistent over tiles):
//   warp 0      : TMEM alloc/dealloc; lane 0 issues all tcgen05.mma + tcgen05.commit
//   warps 1..3  : producers - cp.async (LDGSTS.128, zero-fill outside the image) of the input window
//   warps 4..11 : epilogue  - two sets of four (one per accumulator buffer): tcgen05.ld -> +shift (+residual)
//                 -> LeakyReLU -> 16-bit -> global stores (or the fp32 4x4 flow/mask patch for lastconv)
#include "ptx.cuh"
#include "vfi_internal.h"

namespace vfi {

namespace {

// control block layout at the start of dynamic shared memory
constexpr int kMaxStages = 4;
struct Ctrl {
  uint64_t w_full;
  uint64_t a_full[kMaxStages];
  uint64_t a_empty[kMaxStages];
  uint64_t t_full[2];
  uint64_t t_empty[2];
  uint32_t tmem_base;
};
constexpr uint32_t kCtrlBytes = 256;
constexpr uint32_t kAtabBytes = 8192;  // descriptor table: stages x K=16 steps x {a_lo, a_hi, b_lo, b_hi} (<= 512 entries)
constexpr uint32_t kRowoffBytes = 512; // 128 rows
static_assert(sizeof(Ctrl) <= kCtrlBytes, "control block");

__device__ __forceinline__ size_t out_pixel_offset(const TapConvParams& p, int b, int gy, int gx) {
  // element offset of channel 0 of grid cell (b, gy, gx) in the output tensor
  if (p.out_s2d) {
    // space-to-depth store: the consumer is a stride-2 conv that reads [B, H/2, W/2, 4*n_total]
    size_t cell = ((size_t)b * (p.H >> 1) + (gy >> 1)) * (p.W >> 1) + (gx >> 1);
    return cell * (size_t)(4 * p.n_total) + (size_t)(((gy & 1) * 2 + (gx & 1)) * p.n_total);
  }
  return (((size_t)b * p.H + gy) * p.W + gx) * (size_t)p.n_total;
}

template <typename T>
__global__ void __launch_bounds__(384, 1) tapconv_kernel(const __grid_constant__ TapConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int split = blockIdx.x % p.nsplit;
  const int first = blockIdx.x / p.nsplit;
  if (first >= p.ctas_per_split) return;  // whole CTA leaves together
  const int S = p.stages;
  const bool residual = (p.epi_mode == EPI_RESCONV);

  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_w = smem_base + offsetof(Ctrl, w_full);
  const uint32_t bar_afull = smem_base + offsetof(Ctrl, a_full);
  const uint32_t bar_aempty = smem_base + offsetof(Ctrl, a_empty);
  const uint32_t bar_tfull = smem_base + offsetof(Ctrl, t_full);
  const uint32_t bar_tempty = smem_base + offsetof(Ctrl, t_empty);

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_afull + 8 * s, p.layout == LAYOUT_SWZ ? 1 : kProducerThreads);
      mbar_init(bar_aempty + 8 * s, residual ? 4 : 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 4);
    }
    mbar_fence_init();
  }
  // Descriptor table, built once: for every (stage, K=16 step) the A and B shared-memory descriptors.  The MMA
  // issue loop is then "load 16 bytes, issue" (ncu r01: the issuing warp was the bottleneck when it computed
  // descriptors on the fly - 1240 instructions per tile).
  {
    uint4* tab = reinterpret_cast<uint4*>(smem + kCtrlBytes + kRowoffBytes);
    const int K16 = p.ktotal16;
    for (int id = threadIdx.x; id < S * K16; id += blockDim.x) {
      const int st = id / K16;
      int j = id - st * K16, e = 0;
      const int jj = j;
      while (j >= p.taps[e].nk16) {
        j -= p.taps[e].nk16;
        ++e;
      }
      const TapEntry te = p.taps[e];
      const uint32_t stage_addr = smem_base + p.off_a + (uint32_t)st * p.stage_bytes;
      uint4 d;
      if (p.layout == LAYOUT_SWZ) {
        // rows = window pixels, 128 B (64 ch, SWIZZLE_128B) or 64 B (32-channel tail, SWIZZLE_64B) each; a tap is a
        // start-address shift of whole rows, a K=16 step a 32-byte shift inside the row (the hardware applies the
        // XOR swizzle on the absolute shared-memory address, exactly as TMA wrote it).
        const int ch0 = te.chunk0 * 8 + 16 * j;               // first input channel of this step
        const int kb = ch0 >> 6;
        const bool tail = (kb == p.nkb - 1) && (p.cin & 63);  // 32-channel tail block
        const uint32_t rowb = tail ? 64u : 128u;
        const uint32_t a_addr = stage_addr + p.kb_off[kb] +
                                (uint32_t)((te.dy - p.halo_y0) * p.halo_pitch + (te.dx - p.halo_x0)) * rowb +
                                (uint32_t)(ch0 - kb * 64) * 2u;
        const uint32_t a_sbo = (uint32_t)p.halo_pitch * rowb;
        d.x = ((a_addr >> 4) & 0x3FFFu) | (1u << 16);
        d.y = (a_sbo >> 4) | (1u << 14) | ((tail ? 4u : 2u) << 29);
        if (p.desc_mode & 1) d.y |= ((a_addr >> 7) & 7u) << 17;  // base_offset field, bits [49,52)
        const uint32_t b_addr = smem_base + p.off_w + (uint32_t)(jj >> 2) * ((uint32_t)p.n_cta * 128u) + (uint32_t)(jj & 3) * 32u;
        d.z = ((b_addr >> 4) & 0x3FFFu) | (1u << 16);
        d.w = (1024u >> 4) | (1u << 14) | (2u << 29);
      } else {
        const uint32_t plane16 = (uint32_t)p.plane_bytes >> 4;
        const uint32_t a_off = (uint32_t)(te.chunk0 + 2 * j) * plane16 +
                               (uint32_t)((te.dy - p.halo_y0) * p.halo_w + (te.dx - p.halo_x0));
        d.x = (plane16 << 16) | ((stage_addr >> 4) + a_off);
        d.y = (((uint32_t)p.halo_w * 16u) >> 4) | (1u << 14);   // SBO | version
        d.z = ((((uint32_t)p.n_cta * 16u) >> 4) << 16) | (((smem_base + p.off_w) >> 4) + (uint32_t)jj * (((uint32_t)p.n_cta * 32u) >> 4));
        d.w = (128u >> 4) | (1u << 14);
      }
      tab[id] = d;
    }
  }
  if (warp == 0) tmem_alloc(smem_base + offsetof(Ctrl, tmem_base), p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;

  const uint32_t a_smem = smem_base + p.off_a;
  const uint32_t w_smem = smem_base + p.off_w;
  const int tiles_per_img = p.tiles_y * p.tiles_x;

  if (warp == 0) {
    // ======================================================= MMA issuer
    // The whole warp runs the (warp-uniform) control flow; one elected lane issues tcgen05.mma / tcgen05.commit.
    const uint32_t leader = elect_one_sync();
    mbar_wait(bar_w, 0, 1);
    const uint4* tab0 = reinterpret_cast<const uint4*>(smem + kCtrlBytes + kRowoffBytes);
    const int K16 = p.ktotal16;
    const uint32_t idesc = p.idesc;
    uint32_t k = 0;
    for (int t = first; t < p.ntiles; t += p.ctas_per_split, ++k) {
      const uint32_t stage = k % S, use = k / S, acc = k & 1, vuse = k >> 1;
      mbar_wait(bar_tempty + 8 * acc, (vuse & 1) ^ 1, 2);
      mbar_wait(bar_afull + 8 * stage, use & 1, 3);
      fence_proxy_async();  // cp.async (generic proxy) writes -> tensor-core (async proxy) reads
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * p.acc_stride;
      const uint4* tab = tab0 + stage * K16;
      // K16 is a multiple of 9 for every layer (9 taps x c/16 or 9 x c/32): issue in groups of nine, the next
      // group's descriptors are fetched into registers before the current group is issued, so the shared-memory
      // latency never sits between two MMAs (ncu r01 v2: the issuing warp spent ~140 cycles per MMA).
      uint4 cur[9], nxt[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) cur[i] = tab[i];
      const int ngroups = K16 / 9;
      for (int g = 0; g < ngroups; ++g) {
        if (g + 1 < ngroups) {
#pragma unroll
          for (int i = 0; i < 9; ++i) nxt[i] = tab[(g + 1) * 9 + i];
        }
        if (leader) {
          umma_f16_split(d_tmem, cur[0].x, cur[0].y, cur[0].z, cur[0].w, idesc, g > 0 ? 1u : 0u);
#pragma unroll
          for (int i = 1; i < 9; ++i) umma_f16_split(d_tmem, cur[i].x, cur[i].y, cur[i].z, cur[i].w, idesc, 1u);
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) cur[i] = nxt[i];
      }
      if (leader) {
        if (!residual) umma_commit(bar_aempty + 8 * stage);  // window free once the MMAs have read it
        umma_commit(bar_tfull + 8 * acc);                    // accumulator ready for the epilogue
      }
      __syncwarp();
    }
  } else if (warp < 4) {
    // ======================================================= producers
    const int ptid = threadIdx.x - 32;
    if (ptid == 0) {
      mbar_arrive_expect_tx(bar_w, p.w_bytes);
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w) + (size_t)split * p.w_bytes;
      for (uint32_t off = 0; off < p.w_bytes; off += 16384u) {
        const uint32_t n = min(16384u, p.w_bytes - off);
        bulk_g2s(w_smem + off, wsrc + off, n, bar_w);
      }
    }
    if (p.layout == LAYOUT_SWZ) {
      // one thread feeds the whole pipeline: a 4-D tensor copy per k-block drops the window (zero-filled outside
      // the image = conv padding) into shared memory already in the swizzled K-major operand layout
      if (ptid == 0) {
        uint32_t k = 0;
        for (int t = first; t < p.ntiles; t += p.ctas_per_split, ++k) {
          const uint32_t stage = k % S, use = k / S;
          const int b = t / tiles_per_img;
          const int rem = t - b * tiles_per_img;
          const int tyi = rem / p.tiles_x, txi = rem - tyi * p.tiles_x;
          const int gy0 = tyi * kTileH + p.halo_y0, gx0 = txi * kTileW + p.halo_x0;
          mbar_wait(bar_aempty + 8 * stage, (use & 1) ^ 1, 4);
          mbar_arrive_expect_tx(bar_afull + 8 * stage, p.tx_bytes);
          const uint32_t dst = a_smem + stage * p.stage_bytes;
          for (int kb = 0; kb < p.nkb; ++kb) {
            const bool tail = (kb == p.nkb - 1) && (p.cin & 63);
            tma_load_4d(dst + p.kb_off[kb], tail ? &p.tm32 : &p.tm64, bar_afull + 8 * stage, kb * 64, gx0, gy0, b);
          }
        }
      }
    } else {
    // thread -> fixed 8-channel chunk `ch`, pixels px0, px0+ppi, ... of the window (96 % cpp == 0 for every layer)
    const uint32_t cpp = p.cpp, ppi = kProducerThreads / cpp;
    const uint32_t ch = (uint32_t)ptid % cpp, px0 = (uint32_t)ptid / cpp;
    const uint32_t halo_w = (uint32_t)p.halo_w;
    const uint32_t hy0 = px0 / halo_w, hx0 = px0 - hy0 * halo_w;
    const uint32_t dst0 = ch * (uint32_t)p.plane_bytes + px0 * 16u;
    const uint32_t cin2 = (uint32_t)p.cin * 2u;
    const uint32_t row_bytes = (uint32_t)p.W * cin2;
    const uint32_t step_src = ppi * cin2, step_dst = ppi * 16u;
    const uint32_t wrap_adj = row_bytes - halo_w * cin2;  // next window row, back to its first column
    const int n_it = ((int)p.halo_px - (int)px0 + (int)ppi - 1) / (int)ppi;
    const uint8_t* in = reinterpret_cast<const uint8_t*>(p.in);
    const size_t img_bytes = (size_t)p.H * row_bytes;
    uint32_t k = 0;
    for (int t = first; t < p.ntiles; t += p.ctas_per_split, ++k) {
      const uint32_t stage = k % S, use = k / S;
      const int b = t / tiles_per_img;
      const int rem = t - b * tiles_per_img;
      const int tyi = rem / p.tiles_x, txi = rem - tyi * p.tiles_x;
      const int gy0 = tyi * kTileH + p.halo_y0, gx0 = txi * kTileW + p.halo_x0;
      const uint8_t* img = in + (size_t)b * img_bytes + ch * 16u;
      uint32_t dst = a_smem + stage * p.stage_bytes + dst0;
      uint32_t hx = hx0;
      const bool interior = gy0 >= 0 && gx0 >= 0 && gy0 + p.halo_h <= p.H && gx0 + p.halo_w <= p.W;
      mbar_wait(bar_aempty + 8 * stage, (use & 1) ^ 1, 4);
      if (interior) {
        // whole window inside the image: no per-pixel bounds logic, pointer walks the window row by row
        const uint8_t* src = img + (size_t)(gy0 + (int)hy0) * row_bytes + (size_t)(gx0 + (int)hx0) * cin2;
        for (int it = 0; it < n_it; ++it) {
          cp_async16(dst, src, 16u);
          dst += step_dst;
          src += step_src;
          hx += ppi;
          if (hx >= halo_w) {
            hx -= halo_w;
            src += wrap_adj;
            if (hx >= halo_w) {
              hx -= halo_w;
              src += wrap_adj;
            }
          }
        }
      } else {
        uint32_t hy = hy0;
        for (int it = 0; it < n_it; ++it) {
          const int gy = gy0 + (int)hy, gx = gx0 + (int)hx;
          const bool ok = ((unsigned)gy < (unsigned)p.H) && ((unsigned)gx < (unsigned)p.W);
          const uint32_t off = ok ? (uint32_t)gy * row_bytes + (uint32_t)gx * cin2 : 0u;
          cp_async16(dst, img + off, ok ? 16u : 0u);
          dst += step_dst;
          hx += ppi;
          while (hx >= halo_w) {
            hx -= halo_w;
            ++hy;
          }
        }
      }
      cp_async_arrive_noinc(bar_afull + 8 * stage);
    }
    }
  } else {
    // ======================================================= epilogue (warps 4..7 <-> TMEM lane quarters 0..3)
    // two epilogue warp sets (warps 4..7 and 8..11): set s drains accumulator buffer s, i.e. every other tile, so
    // one tile's epilogue overlaps the next tile's (ncu r01 v5: a single set was busy 100 % of the time)
    const int q = (warp - 4) & 3;
    const uint32_t eset = (uint32_t)(warp - 4) >> 2;
    const int etid = threadIdx.x - 128;
    float* ss = reinterpret_cast<float*>(smem + p.off_ss);  // per-channel shift of this CTA's output slice
    for (int i = etid; i < p.n_cta; i += 256) ss[i] = p.shift[split * p.n_cta + i];
    asm volatile("bar.sync 1, 256;" ::: "memory");

    const int r = q * 32 + lane;           // accumulator row == TMEM lane == tile cell
    const int py = r >> 3, px = r & 7;
    const int n0 = split * p.n_cta;
    const uint32_t center = (uint32_t)((py - p.halo_y0) * p.halo_w + (px - p.halo_x0)) * 16u;
    const uint32_t res_off = (uint32_t)(n0 >> 3) * (uint32_t)p.plane_bytes + center;
    const uint32_t center_px = (uint32_t)((py - p.halo_y0) * p.halo_pitch + (px - p.halo_x0));  // LAYOUT_SWZ
    const int nchunks = p.n_cta >> 4;

    uint32_t k = 0;
    for (int t = first; t < p.ntiles; t += p.ctas_per_split, ++k) {
      const uint32_t stage = k % S, use = k / S, acc = k & 1, vuse = k >> 1;
      if (acc != eset) continue;
      const int b = t / tiles_per_img;
      const int rem = t - b * tiles_per_img;
      const int tyi = rem / p.tiles_x, txi = rem - tyi * p.tiles_x;
      const int gy = tyi * kTileH + py, gx = txi * kTileW + px;
      const bool valid = (gy < p.H) && (gx < p.W);

      mbar_wait(bar_tfull + 8 * acc, vuse & 1, 5);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * p.acc_stride + ((uint32_t)(q * 32) << 16);

      if (p.epi_mode == EPI_LASTCONV) {
        // n = c5*16 + (y4*4 + x4): 16-column chunk c5 = one component (4 flow + mask) of the 4x4 sub-pixel patch
        const int Hs = p.H * 4, Ws = p.W * 4;
        if (p.n_cta == 80) {
          uint32_t v[5][16];
#pragma unroll
          for (int c = 0; c < 5; ++c) tmem_ld16(taddr + c * 16, v[c]);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
          if (valid) {
#pragma unroll
            for (int y4 = 0; y4 < 4; ++y4) {
              const size_t o = ((size_t)b * Hs + (gy * 4 + y4)) * Ws + gx * 4;
#pragma unroll
              for (int x4 = 0; x4 < 4; ++x4) {
                const int pos = y4 * 4 + x4;
                float4 f;
                f.x = __uint_as_float(v[0][pos]) + ss[0 * 16 + pos];
                f.y = __uint_as_float(v[1][pos]) + ss[1 * 16 + pos];
                f.z = __uint_as_float(v[2][pos]) + ss[2 * 16 + pos];
                f.w = __uint_as_float(v[3][pos]) + ss[3 * 16 + pos];
                p.out_flow[o + x4] = f;
                p.out_mask[o + x4] = __uint_as_float(v[4][pos]) + ss[4 * 16 + pos];
              }
            }
          }
        } else {
          // output channels split across CTAs (large c): this CTA owns components [split*n_cta/16, ...)
          const int ncomp = p.n_cta >> 4;
          for (int cc = 0; cc < ncomp; ++cc) {
            uint32_t v[16];
            tmem_ld16(taddr + cc * 16, v);
            tmem_ld_wait();
            const int c5 = split * ncomp + cc;
            if (valid) {
#pragma unroll
              for (int pos = 0; pos < 16; ++pos) {
                const size_t o = ((size_t)b * Hs + (gy * 4 + (pos >> 2))) * Ws + gx * 4 + (pos & 3);
                const float val = __uint_as_float(v[pos]) + ss[cc * 16 + pos];
                if (c5 < 4)
                  reinterpret_cast<float*>(p.out_flow)[o * 4 + c5] = val;
                else
                  p.out_mask[o] = val;
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
        }
        continue;
      }

      if (residual) mbar_wait(bar_afull + 8 * stage, use & 1, 6);  // acquire the staged window for generic reads
      T* orow = reinterpret_cast<T*>(p.out) + (valid ? out_pixel_offset(p, b, gy, gx) + (size_t)n0 : 0);
      const uint32_t ra0 = a_smem + stage * p.stage_bytes + res_off;
      for (int cc = 0; cc < nchunks; cc += 2) {
        const bool two = (cc + 1 < nchunks);
        uint32_t v[2][16];
        tmem_ld16(taddr + cc * 16, v[0]);
        if (two) tmem_ld16(taddr + cc * 16 + 16, v[1]);
        // shared-memory operands of this step are fetched while the TMEM load is in flight
        float4 sh[2][4];
        uint4 rr[2][2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h == 0 || two) {
            const float4* sp = reinterpret_cast<const float4*>(ss + (cc + h) * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i) sh[h][i] = sp[i];
            if (residual) {
              uint32_t ra, rb;  // shared-memory addresses of the two 16-byte chunks (16 residual channels)
              if (p.layout == LAYOUT_SWZ) {
                const int ch0 = n0 + (cc + h) * 16;
                const int kb = ch0 >> 6;
                const bool tail = (kb == p.nkb - 1) && (p.cin & 63);
                const uint32_t rowb = tail ? 64u : 128u, msk = tail ? 3u : 7u;
                const uint32_t row = a_smem + stage * p.stage_bytes + p.kb_off[kb] + center_px * rowb;
                const uint32_t c0 = (uint32_t)(ch0 - kb * 64) >> 3, sw = (row >> 7) & msk;
                ra = row + ((c0 ^ sw) << 4);
                rb = row + (((c0 + 1u) ^ sw) << 4);
              } else {
                ra = ra0 + (uint32_t)(2 * (cc + h)) * (uint32_t)p.plane_bytes;
                rb = ra + (uint32_t)p.plane_bytes;
              }
              asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                           : "=r"(rr[h][0].x), "=r"(rr[h][0].y), "=r"(rr[h][0].z), "=r"(rr[h][0].w)
                           : "r"(ra));
              asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                           : "=r"(rr[h][1].x), "=r"(rr[h][1].y), "=r"(rr[h][1].z), "=r"(rr[h][1].w)
                           : "r"(rb));
            }
          }
        }
        tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h == 0 || two) {
            const float shf[16] = {sh[h][0].x, sh[h][0].y, sh[h][0].z, sh[h][0].w, sh[h][1].x, sh[h][1].y,
                                   sh[h][1].z, sh[h][1].w, sh[h][2].x, sh[h][2].y, sh[h][2].z, sh[h][2].w,
                                   sh[h][3].x, sh[h][3].y, sh[h][3].z, sh[h][3].w};
            const uint32_t rw[8] = {rr[h][0].x, rr[h][0].y, rr[h][0].z, rr[h][0].w,
                                    rr[h][1].x, rr[h][1].y, rr[h][1].z, rr[h][1].w};
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float a0 = __uint_as_float(v[h][2 * i]) + shf[2 * i];
              float a1 = __uint_as_float(v[h][2 * i + 1]) + shf[2 * i + 1];
              if (residual) {
                const float2 rf = Pack2<T>::unpack(rw[i]);
                a0 += rf.x;
                a1 += rf.y;
              }
              o[i] = Pack2<T>::pack(fmaxf(a0, 0.2f * a0), fmaxf(a1, 0.2f * a1));  // LeakyReLU(0.2)
            }
            // each thread owns one grid cell: 32 contiguous bytes (16 channels) of its channel vector per step.
            // (ncu r01: a shared-memory transpose for 512-byte-contiguous stores cost more instructions and
            // shared-memory wavefronts than it saved; L2 merges the 16-byte halves of a sector.)
            if (valid) {
              uint4* dst = reinterpret_cast<uint4*>(orow + (cc + h) * 16);
              dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
              dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
            }
          }
        }
      }
      tc_fence_before();
      if (residual && p.layout == LAYOUT_SWZ) fence_proxy_async();  // generic reads before the next TMA write
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_tempty + 8 * acc);
        if (residual) mbar_arrive(bar_aempty + 8 * stage);
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------
// CUDA-core checker with the SAME parameters, packed weights and epilogue: one thread per (cell, n).
// Test infrastructure for the tensor-core kernel (debug entry point only; never on the product path).
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float ld16bit(const T* p);
template <>
__device__ __forceinline__ float ld16bit<__half>(const __half* p) { return __half2float(*p); }
template <>
__device__ __forceinline__ float ld16bit<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T>
__device__ __forceinline__ T cvt16bit(float v);
template <>
__device__ __forceinline__ __half cvt16bit<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 cvt16bit<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T>
__global__ void tapconv_ref_kernel(const __grid_constant__ TapConvParams p) {
  const size_t total = (size_t)p.B * p.H * p.W * p.n_total;
  for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(id % p.n_total);
    size_t cell = id / p.n_total;
    const int gx = (int)(cell % p.W);
    cell /= p.W;
    const int gy = (int)(cell % p.H);
    const int b = (int)(cell / p.H);
    const int split = n / p.n_cta, nl = n - split * p.n_cta;
    const T* w = reinterpret_cast<const T*>(p.w) + (size_t)split * (p.w_bytes / 2);
    const T* in = reinterpret_cast<const T*>(p.in);
    float acc = 0.f;
    int j = 0;  // running K=16 step
    for (int e = 0; e < p.ntaps; ++e) {
      const TapEntry te = p.taps[e];
      const int y = gy + te.dy, x = gx + te.dx;
      const bool ok = y >= 0 && y < p.H && x >= 0 && x < p.W;
      for (int i = 0; i < te.nk16; ++i, ++j) {
        if (!ok) continue;
        for (int c = 0; c < 16; ++c) {
          const int cin_idx = te.chunk0 * 8 + i * 16 + c;
          const float a = ld16bit<T>(in + (((size_t)b * p.H + y) * p.W + x) * p.cin + cin_idx);
          size_t widx;
          if (p.layout == LAYOUT_SWZ) {  // [j/4][n][128 B row, 16-byte chunks XOR (n & 7)]
            const int chunk = (j & 3) * 2 + (c >> 3);
            widx = ((size_t)(j >> 2) * p.n_cta + nl) * 64 + (size_t)((chunk ^ (nl & 7)) * 8 + (c & 7));
          } else {
            const int kk = j * 2 + (c >> 3);  // 8-channel K chunk
            widx = ((size_t)kk * p.n_cta + nl) * 8 + (c & 7);
          }
          const float wv = ld16bit<T>(w + widx);
          acc = fmaf(a, wv, acc);
        }
      }
    }
    if (p.epi_mode == EPI_LASTCONV) {
      const int c5 = n >> 4, pos = n & 15;
      const int Hs = p.H * 4, Ws = p.W * 4;
      const size_t o = ((size_t)b * Hs + gy * 4 + (pos >> 2)) * Ws + gx * 4 + (pos & 3);
      const float v = acc + p.shift[n];
      if (c5 < 4)
        reinterpret_cast<float*>(p.out_flow)[o * 4 + c5] = v;
      else
        p.out_mask[o] = v;
    } else {
      float v = acc + p.shift[n];
      if (p.epi_mode == EPI_RESCONV) v += ld16bit<T>(in + (((size_t)b * p.H + gy) * p.W + gx) * p.cin + n);
      reinterpret_cast<T*>(p.out)[out_pixel_offset(p, b, gy, gx) + n] = cvt16bit<T>(lrelu02(v));
    }
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
bool make_tmap(CUtensorMap* tm, CUtensorMapDataType dt, const void* base, int C, int W, int H, int B, int box_c,
               int box_w, int box_h, CUtensorMapSwizzle swz) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) {
      set_error("cuTensorMapEncodeTiled is not available from this driver");
      return false;
    }
    fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  const CUresult r = fn(tm, dt, 4, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return false;
  }
  return true;
}

uint32_t ceil_magic(uint32_t d) { return (uint32_t)((0x100000000ull + d - 1) / d); }
uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

}  // namespace

int tapconv_plan(const TapConvLayer& L, TapConvParams* pp) {
  TapConvParams& p = *pp;
  p.cin = L.cin;
  p.n_total = L.n_total;
  p.n_cta = L.n_cta;
  p.nsplit = L.nsplit;
  p.ntaps = L.ntaps;
  p.ktotal16 = L.ktotal16;
  p.halo_y0 = L.halo_y0;
  p.halo_x0 = L.halo_x0;
  p.halo_h = L.halo_h;
  p.halo_w = L.halo_w;
  p.halo_px = L.halo_h * L.halo_w;
  p.epi_mode = L.epi_mode;
  p.out_s2d = L.out_s2d;
  for (int e = 0; e < L.ntaps; ++e) p.taps[e] = L.taps[e];
  int plane_px = p.halo_px | 1;  // odd number of 16-byte slots: the 8 planes one pixel is scattered to hit 8 banks
  p.plane_bytes = plane_px * 16;
  p.cpp = (uint32_t)L.cin / 8;
  p.cpp_magic = ceil_magic(p.cpp);
  p.halow_magic = ceil_magic((uint32_t)L.halo_w);
  p.layout = L.layout;
  p.desc_mode = L.desc_mode;
  p.halo_pitch = L.halo_pitch > 0 ? L.halo_pitch : L.halo_w;
  uint32_t walign = 128;
  if (L.layout == LAYOUT_SWZ) {
    // weights: [ceil(K16/4)][n_cta] rows of 128 B (four K=16 steps), 128B-swizzled; window: one region per k-block
    p.w_bytes = (uint32_t)((L.ktotal16 + 3) / 4) * (uint32_t)L.n_cta * 128u;
    p.nkb = (L.cin + 63) / 64;
    uint32_t off = 0;
    p.tx_bytes = 0;
    for (int kb = 0; kb < p.nkb; ++kb) {
      const bool tail = (kb == p.nkb - 1) && (L.cin & 63);
      const uint32_t bytes = (uint32_t)p.halo_pitch * (uint32_t)L.halo_h * (tail ? 64u : 128u);
      p.kb_off[kb] = off;
      off += align_up(bytes, 1024);
      p.tx_bytes += bytes;
    }
    p.stage_bytes = off;
    walign = 1024;
  } else {
    p.w_bytes = (uint32_t)L.ktotal16 * 2u * (uint32_t)L.n_cta * 16u;
    p.stage_bytes = align_up(p.cpp * (uint32_t)p.plane_bytes, 128);
  }
  p.epi_pitch = (uint32_t)L.n_cta * 2u + 16u;
  const uint32_t epi_bytes = 0u;  // the epilogue stores straight from registers
  p.off_ss = kCtrlBytes + kRowoffBytes + kAtabBytes;
  p.cpo_magic = ceil_magic((uint32_t)L.n_cta / 8);
  p.off_w = align_up(p.off_ss + (uint32_t)L.n_cta * 4u, walign);
  p.off_a = align_up(p.off_w + p.w_bytes, walign);
  int stages = 0;
  for (int s = kMaxStages; s >= 1; --s) {
    if (p.off_a + (uint32_t)s * p.stage_bytes + epi_bytes <= (uint32_t)kSmemLimit) {
      stages = s;
      break;
    }
  }
  while (stages > 1 && stages * L.ktotal16 > (int)(kAtabBytes / 16)) --stages;  // descriptor table capacity
  if (stages * L.ktotal16 > (int)(kAtabBytes / 16)) stages = 0;
  p.stages = stages;
  p.off_epi = p.off_a + (uint32_t)stages * p.stage_bytes;
  // accumulators: two buffers of n_cta fp32 columns, allocation is a power of two >= 32
  uint32_t stride = 16;
  while (stride < (uint32_t)L.n_cta) stride <<= 1;
  p.acc_stride = stride;
  p.tmem_cols = stride * 2 < 32 ? 32 : stride * 2;
  return stages;
}

cudaError_t launch_tapconv(const TapConvLayer& L, int op_type, const void* in, void* out, float4* out_flow,
                           float* out_mask, int B, int H, int W, int num_sms, bool use_ref, cudaStream_t st) {
  TapConvParams p{};
  if (tapconv_plan(L, &p) < 1) {
    set_error("tapconv: layer does not fit in shared memory");
    return cudaErrorInvalidConfiguration;
  }
  if (L.out_s2d && ((H | W) & 1)) {
    set_error("tapconv: space-to-depth output needs even H and W");
    return cudaErrorInvalidValue;
  }
  p.in = in;
  p.out = out;
  p.out_flow = out_flow;
  p.out_mask = out_mask;
  p.w = L.w;
  p.shift = L.shift;
  p.B = B;
  p.H = H;
  p.W = W;
  p.tiles_y = (H + kTileH - 1) / kTileH;
  p.tiles_x = (W + kTileW - 1) / kTileW;
  p.ntiles = B * p.tiles_y * p.tiles_x;
  // instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): D=f32, A/B = f16|bf16, both K-major,
  // N>>3 at [17,23), M>>4 at [24,29)
  const uint32_t fmt = (op_type == OP_BF16) ? 1u : 0u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(L.n_cta >> 3) << 17) | ((128u >> 4) << 24);

  if (use_ref) {
    const size_t total = (size_t)B * H * W * L.n_total;
    const int blocks = (int)((total + 255) / 256 < 65535 ? (total + 255) / 256 : 65535);
    if (op_type == OP_BF16)
      tapconv_ref_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(p);
    else
      tapconv_ref_kernel<__half><<<blocks, 256, 0, st>>>(p);
    return cudaGetLastError();
  }

  if (L.layout == LAYOUT_SWZ) {
    const CUtensorMapDataType dt = (op_type == OP_BF16) ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    if (!make_tmap(&p.tm64, dt, in, L.cin, W, H, B, 64, p.halo_pitch, L.halo_h, CU_TENSOR_MAP_SWIZZLE_128B))
      return cudaErrorInvalidValue;
    if (L.cin & 63) {
      if (!make_tmap(&p.tm32, dt, in, L.cin, W, H, B, 32, p.halo_pitch, L.halo_h, CU_TENSOR_MAP_SWIZZLE_64B))
        return cudaErrorInvalidValue;
    } else {
      p.tm32 = p.tm64;
    }
  }
  int cps = num_sms / L.nsplit;
  if (cps < 1) cps = 1;
  if (cps > p.ntiles) cps = p.ntiles;
  p.ctas_per_split = cps;
  const int grid = cps * L.nsplit;
  const size_t smem = p.off_epi;
  cudaError_t err;
  if (op_type == OP_BF16) {
    err = cudaFuncSetAttribute(tapconv_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (err != cudaSuccess) return err;
    tapconv_kernel<__nv_bfloat16><<<grid, 384, smem, st>>>(p);
  } else {
    err = cudaFuncSetAttribute(tapconv_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (err != cudaSuccess) return err;
    tapconv_kernel<__half><<<grid, 384, smem, st>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace vfi
