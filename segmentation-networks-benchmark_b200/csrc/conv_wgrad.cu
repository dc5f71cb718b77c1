// Weight gradients of the tap-list convolutions on the sm_100a tensor cores (tcgen05 + TMEM), fed by TMA.
//
// Backward half of the LinkNet34 training step (BASELINE configs[1]; the reference reaches it through autograd + cuDNN,
// torch_train.py:186-189).  For every forward kind of conv_tcgen05.cu (conv3x3, conv1x1, conv2x2, ConvTranspose phases)
//     dW[phase][tap][co][ci] = sum_{n,y,x} dY_phase[n][y][x][co] * X[n][y + dy(tap) + off][x + dx(tap) + off][ci]
// which is a GEMM whose K dimension is the PIXEL index: both operands are read straight from the NHWC bf16 slabs the
// forward pass left behind, as MN-major UMMA operands (the channel index is contiguous in memory):
//     A (M side) = X    : one (8+2) x (16+2) halo box per 64-channel chunk; a tap is a different start row of the shared-
//                         memory descriptor, exactly as in the forward halo kernel (K groups of 8 pixels = one patch row,
//                         SBO = one halo row = 10 pixels)
//     B (N side) = dY   : one 8 x 16 patch box per 64-channel chunk (TMA out-of-bounds zero fill gives the conv padding on
//                         X and the ragged patch edges on dY)
//     D[ci][co] per tap : fp32 in TMEM, one accumulator of 64 columns per tap (up to 8 taps per CTA)
// A CTA owns (phase, 64/128 input channels, 64 output channels, a group of taps, a range of pixel patches), runs a TMA ->
// MMA ring over its patches, then its four epilogue warps add the accumulators into the packed fp32 gradient
// [phase * taps + tap][co][ci] with red.global.add.f32: TMEM lanes are consecutive ci, so every warp-wide add is one
// coalesced 128-byte line.  Split-K over patch ranges fills the 148 SMs when the weight tile count is small.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstring>
#include <new>

#include "sm100_ptx.cuh"
#include "snb_internal.h"

namespace snb {

constexpr int kWgMaxPhases = 4;
constexpr int kWgMaxTaps = 9;
constexpr int kWgHaloW = 10, kWgHaloH = 18;
constexpr int kWgChunk = 64;                                   // channels per operand chunk = one 128-byte swizzle span
constexpr int kWgHBytes = ((kWgHaloW * kWgHaloH * 128) + 1023) / 1024 * 1024;   // 23552: one halo chunk
constexpr int kWgPBytes = 128 * 128;                           // one patch chunk (128 pixels x 128 bytes)
constexpr int kWgN = 64;                                       // output channels per CTA
constexpr int kWgMaxStages = 6;
constexpr int kWgSmemBudget = 227 * 1024;

struct alignas(64) WgradParams {
  CUtensorMap map_h;                     // X   (C, W, H, N), box (64, 10, 18, 1)
  CUtensorMap map_p[kWgMaxPhases];       // dY per phase (C, W, H, N), box (64, 8, 16, 1)
  int32_t n_phases, taps;                // taps per phase
  int32_t m_blocks, n_blocks, mb;        // M blocks (mb chunks of 64 X channels each), N blocks (64 dY channels)
  int32_t tap_groups, tg;                // tap groups per phase, taps per group
  int32_t tiles_x, tiles_y, n_img, patches;
  int32_t k_splits, stages, load_off;
  int32_t h_c, p_c;                      // real channel counts of X and dY (stores are clipped to them)
  int64_t out_tap_stride, out_row_stride;
  float* out;
  int8_t tap_dy[kWgMaxPhases][kWgMaxTaps];
  int8_t tap_dx[kWgMaxPhases][kWgMaxTaps];
};

// instruction descriptor: bf16 x bf16 -> fp32, both operands MN-major (bits 15 / 16)
__host__ __device__ constexpr uint32_t make_idesc_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// MN-major SW128 descriptor halves: lo = start >> 4 | LBO >> 4 << 16 (distance between 64-element MN blocks),
// hi = SBO >> 4 (distance between groups of 8 K rows) | version 1 | layout SW128
__device__ __forceinline__ uint64_t mn_desc(uint32_t addr16, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  const uint32_t lo = (addr16 & 0x3FFFu) | ((lbo_bytes >> 4) << 16);
  const uint32_t hi = (sbo_bytes >> 4) | (1u << 14) | (2u << 29);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

// M = 64 * MB (MB = 1: 64 x N accumulators live in lanes 0-15 of every 32-lane quarter)
template <int MB>
__global__ void __launch_bounds__(256, 1) conv_wgrad_kernel(const __grid_constant__ WgradParams p) {
  constexpr int M = 64 * MB;
  constexpr uint32_t IDESC = make_idesc_mn(M, kWgN);
  constexpr int STAGE_BYTES = MB * kWgHBytes + kWgPBytes;
  constexpr int TCOLS = 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * STAGE_BYTES);
  uint64_t* full = bars;                         // [kWgMaxStages]
  uint64_t* empty = full + kWgMaxStages;         // [kWgMaxStages]
  uint64_t* done = empty + kWgMaxStages;         // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- which unit is this CTA
  int u = blockIdx.x;
  const int split = u % p.k_splits; u /= p.k_splits;
  const int grp = u % p.tap_groups; u /= p.tap_groups;
  const int nb = u % p.n_blocks; u /= p.n_blocks;
  const int mblk = u % p.m_blocks; u /= p.m_blocks;
  const int ph = u;
  const int tap0 = grp * p.tg;
  const int n_taps = min(p.tg, p.taps - tap0);
  const int pb = static_cast<int>(static_cast<int64_t>(p.patches) * split / p.k_splits);
  const int pe = static_cast<int>(static_cast<int64_t>(p.patches) * (split + 1) / p.k_splits);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_h);
    tma_prefetch_desc(&p.map_p[ph]);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kWgMaxStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, TCOLS);
    tmem_relinquish();
  }
  tc05_fence_before();
  __syncthreads();
  tc05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t s = 0, par = 1;
      for (int t = pb; t < pe; ++t) {
        int q = t;
        const int x0 = (q % p.tiles_x) * 8; q /= p.tiles_x;
        const int y0 = (q % p.tiles_y) * 16; q /= p.tiles_y;
        const int img = q;
        mbar_wait(&empty[s], par);
        uint8_t* st = smem + s * STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], MB * kWgHaloW * kWgHaloH * 128 + kWgPBytes);
#pragma unroll
        for (int c = 0; c < MB; ++c)
          tma_load_4d(&p.map_h, &full[s], st + c * kWgHBytes, (mblk * MB + c) * kWgChunk, x0 - 1 + p.load_off,
                      y0 - 1 + p.load_off, img);
        tma_load_4d(&p.map_p[ph], &full[s], st + MB * kWgHBytes, nb * kWgChunk, x0, y0, img);
        if (++s == static_cast<uint32_t>(p.stages)) { s = 0; par ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (single thread)
    if (elect_one()) {
      uint32_t aoff[kWgMaxTaps];
#pragma unroll
      for (int i = 0; i < kWgMaxTaps; ++i) {
        const int t = min(tap0 + i, p.taps - 1);
        aoff[i] = static_cast<uint32_t>(((p.tap_dy[ph][t] + 1) * kWgHaloW + (p.tap_dx[ph][t] + 1)) * 128) >> 4;
      }
      const uint32_t base16 = (smem_u32(smem) & 0x3FFFFu) >> 4;
      uint32_t s = 0, par = 0;
      for (int t = pb; t < pe; ++t) {
        mbar_wait(&full[s], par);
        tc05_fence_after();
        const uint32_t a16 = base16 + s * (STAGE_BYTES >> 4);
        const uint32_t b16 = a16 + MB * (kWgHBytes >> 4);
#pragma unroll 1
        for (int i = 0; i < n_taps; ++i) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            // K = 16 pixels = patch rows 2j, 2j+1: two halo rows down (2 * 10 * 128 B) on the X side, 16 rows on the dY side
            const uint64_t adesc = mn_desc(a16 + aoff[i] + j * (2 * kWgHaloW * 128 >> 4), kWgHBytes, kWgHaloW * 128);
            const uint64_t bdesc = mn_desc(b16 + j * (16 * 128 >> 4), kWgPBytes, 8 * 128);
            umma_bf16_ss(adesc, bdesc, tmem_base + i * kWgN, IDESC, (t > pb || j > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty[s]);
        if (++s == static_cast<uint32_t>(p.stages)) { s = 0; par ^= 1; }
      }
      umma_commit(done);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: TMEM -> red.global.add
    if (pe > pb) {
      mbar_wait(done, 0);
      tc05_fence_after();
      const int q = warp & 3;
      const bool row_ok = MB == 2 || lane < 16;
      const int m_ch = mblk * M + (MB == 2 ? q * 32 + lane : q * 16 + lane);
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      for (int i = 0; i < n_taps; ++i) {
        float* dst = p.out + static_cast<int64_t>(ph * p.taps + tap0 + i) * p.out_tap_stride + m_ch;
#pragma unroll 1
        for (int c0 = 0; c0 < kWgN; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(t_addr + i * kWgN + c0, v);
          tmem_ld_wait();
          if (row_ok && m_ch < p.h_c) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n_ch = nb * kWgN + c0 + j;
              if (n_ch < p.p_c) atomicAdd(dst + static_cast<int64_t>(n_ch) * p.out_row_stride, __uint_as_float(v[j]));
            }
          }
        }
      }
    }
  }

  tc05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc05_fence_after();
    tmem_dealloc(tmem_base, TCOLS);
  }
}

}  // namespace snb

struct snb_wgrad {
  snb::WgradParams params;
  const void* fn;
  int smem, grid;
  double flops;
};

using namespace snb;

extern "C" int snb_wgrad_create(const snb_wgrad_desc* d, snb_wgrad** out) {
  if (!d || !out) return fail(SNB_E_INVALID, "snb_wgrad_create: null argument");
  *out = nullptr;
  if (d->kind < SNB_CONV_3X3 || d->kind > SNB_CONV_2X2_ADJ) return fail(SNB_E_INVALID, "unknown conv kind %d", d->kind);
  if (d->n <= 0 || d->h <= 0 || d->w <= 0 || d->cin <= 0 || d->cout <= 0) return fail(SNB_E_INVALID, "bad shape");
  if (!d->d_in || !d->d_dout || !d->d_dweight) return fail(SNB_E_INVALID, "null tensor pointer");
  if (d->in_cstride < d->cin || d->in_cstride % 8 || d->dout_cstride < d->cout || d->dout_cstride % 8)
    return fail(SNB_E_INVALID, "slab strides must be multiples of 8 channels covering the channels");
  if ((reinterpret_cast<uintptr_t>(d->d_in) & 15) || (reinterpret_cast<uintptr_t>(d->d_dout) & 15) ||
      (reinterpret_cast<uintptr_t>(d->d_dweight) & 3))
    return fail(SNB_E_INVALID, "tensor pointers must be 16-byte aligned");
  if (d->dw_cout < d->cout || d->dw_cin < d->cin) return fail(SNB_E_INVALID, "packed gradient smaller than cout x cin");

  snb_wgrad* c = new (std::nothrow) snb_wgrad();
  if (!c) return fail(SNB_E_INVALID, "out of host memory");
  WgradParams& p = c->params;
  std::memset(&p, 0, sizeof(p));
  ConvTapGeom tg;
  if (int rc = conv_tap_geometry(d->kind, d->valid, d->h, d->w, &tg)) { delete c; return rc; }
  p.n_phases = tg.n_phases;
  p.taps = tg.taps;
  std::memcpy(p.tap_dy, tg.tap_dy, sizeof(p.tap_dy));
  std::memcpy(p.tap_dx, tg.tap_dx, sizeof(p.tap_dx));
  p.load_off = tg.load_off;
  p.h_c = static_cast<int32_t>(d->cin);
  p.p_c = static_cast<int32_t>(d->cout);
  p.mb = d->cin > 64 ? 2 : 1;
  const int M = 64 * p.mb;
  p.m_blocks = static_cast<int32_t>((d->cin + M - 1) / M);
  p.n_blocks = static_cast<int32_t>((d->cout + kWgN - 1) / kWgN);
  p.tg = std::min<int32_t>(p.taps, 512 / kWgN);                    // accumulators of 64 columns, 512 TMEM columns
  p.tap_groups = (p.taps + p.tg - 1) / p.tg;
  p.tg = (p.taps + p.tap_groups - 1) / p.tap_groups;              // balance the groups (9 taps -> 5 + 4)
  p.tiles_x = static_cast<int32_t>((tg.grid_w + 7) / 8);
  p.tiles_y = static_cast<int32_t>((tg.grid_h + 15) / 16);
  p.n_img = static_cast<int32_t>(d->n);
  const int64_t patches = (int64_t)p.tiles_x * p.tiles_y * p.n_img;
  if (patches > INT32_MAX) { delete c; return fail(SNB_E_UNSUPPORTED, "too many patches"); }
  p.patches = static_cast<int32_t>(patches);
  const int sms = sm_count();
  if (sms <= 0) { delete c; return fail(SNB_E_CUDA, "no CUDA device"); }
  const int64_t base_units = (int64_t)p.n_phases * p.m_blocks * p.n_blocks * p.tap_groups;
  // split-K: fill the SMs, but keep >= 2 patches per CTA (prologue + an epilogue of M x 64 x taps atomics per CTA)
  int64_t splits = std::max<int64_t>(1, (2 * sms + base_units - 1) / base_units);
  splits = std::min<int64_t>(splits, std::max<int64_t>(1, patches / 2));
  p.k_splits = static_cast<int32_t>(splits);
  const int stage_bytes = p.mb * kWgHBytes + kWgPBytes;
  p.stages = std::min<int>(kWgMaxStages, (kWgSmemBudget - 2048) / stage_bytes);
  p.stages = static_cast<int32_t>(std::max<int64_t>(2, std::min<int64_t>(p.stages, (patches + splits - 1) / splits + 1)));
  c->smem = 1024 + p.stages * stage_bytes + 1024;
  p.out = d->d_dweight;
  p.out_row_stride = d->dw_cin;
  p.out_tap_stride = d->dw_cout * d->dw_cin;

  int rc;
  {
    // channel extents are rounded up to whole 16-byte vectors inside the slab: the extra channels only feed accumulator
    // rows / columns that are never stored
    uint64_t dims[4] = {(uint64_t)std::min<int64_t>(d->in_cstride, (d->cin + 7) / 8 * 8), (uint64_t)d->w, (uint64_t)d->h, (uint64_t)d->n};
    uint64_t str[3] = {(uint64_t)d->in_cstride * 2, (uint64_t)d->w * d->in_cstride * 2, (uint64_t)d->h * d->w * d->in_cstride * 2};
    uint32_t box[4] = {kWgChunk, kWgHaloW, kWgHaloH, 1};
    rc = encode_map(&p.map_h, const_cast<void*>(d->d_in), 4, dims, str, box, 128, 2);
    if (rc) { delete c; return rc; }
  }
  // dY: the forward's output tensor; a ConvTranspose phase is a stride-2 view of it
  const int s = tg.n_phases == 4 ? 2 : 1;
  for (int ph = 0; ph < tg.n_phases; ++ph) {
    const int py = ph / 2, px = ph % 2;
    const char* base = static_cast<const char*>(d->d_dout) + ((int64_t)py * tg.out_w + px) * d->dout_cstride * 2;
    const int64_t pw = s == 1 ? tg.out_w : (tg.out_w - px + 1) / 2, phh = s == 1 ? tg.out_h : (tg.out_h - py + 1) / 2;
    uint64_t dims[4] = {(uint64_t)std::min<int64_t>(d->dout_cstride, (d->cout + 7) / 8 * 8), (uint64_t)pw, (uint64_t)phh, (uint64_t)d->n};
    uint64_t str[3] = {(uint64_t)s * d->dout_cstride * 2, (uint64_t)s * tg.out_w * d->dout_cstride * 2,
                       (uint64_t)tg.out_h * tg.out_w * d->dout_cstride * 2};
    uint32_t box[4] = {kWgChunk, 8, 16, 1};
    rc = encode_map(&p.map_p[ph], const_cast<char*>(base), 4, dims, str, box, 128, 2);
    if (rc) { delete c; return rc; }
  }
  c->fn = p.mb == 2 ? reinterpret_cast<const void*>(&conv_wgrad_kernel<2>) : reinterpret_cast<const void*>(&conv_wgrad_kernel<1>);
  cudaError_t e = cudaFuncSetAttribute(c->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBudget);
  if (e != cudaSuccess) { delete c; return fail(SNB_E_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); }
  c->grid = static_cast<int>(base_units * splits);
  c->flops = 2.0 * (double)d->n * tg.grid_h * tg.grid_w * (double)d->cin * d->cout * p.taps * p.n_phases;
  *out = c;
  return SNB_OK;
}

extern "C" int snb_wgrad_launch(const snb_wgrad* c, void* stream) {
  if (!c) return fail(SNB_E_INVALID, "snb_wgrad_launch: null handle");
  void* args[1] = {const_cast<WgradParams*>(&c->params)};
  cudaError_t e = cudaLaunchKernel(c->fn, dim3(c->grid), dim3(256), args, c->smem, as_stream(stream));
  if (e != cudaSuccess) return fail(SNB_E_CUDA, "wgrad launch failed: %s", cudaGetErrorString(e));
  return SNB_OK;
}

extern "C" void snb_wgrad_destroy(snb_wgrad* c) { delete c; }

extern "C" double snb_wgrad_flops(const snb_wgrad* c) { return c ? c->flops : 0.0; }
