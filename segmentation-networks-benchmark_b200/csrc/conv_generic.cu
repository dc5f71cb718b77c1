// Generic convolution gradients on NHWC bf16 slabs (the backward half of LinkNet34's training step, BASELINE configs[1]).
//
// ONE geometry covers every layer of lib/models/linknet.py (torchvision resnet34 blocks :39-48, DecoderBlockLinkNet :16-31,
// the head :57-62): a convolution  small[n, oy, ox, co] = sum_{ky,kx,ci} big[n, oy*s + ky - p, ox*s + kx - p, ci] * W[co][ci][ky][kx]
// with kernel kh x kw, stride s, padding p, mapping the "big" tensor (stride side) to the "small" one.  nn.Conv2d is that
// map with big = input; nn.ConvTranspose2d is its adjoint with big = output and weight W[co = Cin_t][ci = Cout_t].
// Three implicit GEMMs C[M][N] = sum_K A * B share one 64 x 64 x 16 register-tiled CUDA-core kernel (fp32 accumulate):
//   FWD   (ConvTranspose dgrad):  M = small pixels, N = co, K = (tap, ci):  small = conv(big, W)
//   DGRAD (Conv2d dgrad, ConvTranspose forward):  M = big pixels, N = ci, K = (tap, co):  big = conv^T(small, W)
//   WGRAD: per tap  M = co, N = ci, K = small pixels:  dW[co][ci][ky][kx] = sum_p small[p][co] * big[p*s + k - p][ci]
//          written with fp32 atomics (split-K over pixel ranges) straight into the PyTorch-layout gradient tensor.
// Weights are read in the PyTorch layout (float [co][ci][kh][kw]), so no packing step sits between the optimiser and the
// backward pass.  This is the round-1 correctness path: CUDA-core FMAs at a fraction of the tensor peak; the tcgen05
// dgrad (forward kernels on transformed weights) and wgrad (K = pixels, MN-major operands) replace it next round
// (DESIGN.md section 7).
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "snb_internal.h"

namespace snb {

struct ConvGeom {
  int n, bh, bw, bc;      // big tensor: images, height, width, channels (ci)
  int sh, sw, sc;         // small tensor: height, width, channels (co)
  int kh, kw, stride, pad;
  int64_t big_cs, small_cs;   // channel strides (elements per pixel) of the two slabs
};

constexpr int kGT = 64;    // tile edge
constexpr int kGK = 16;    // K chunk

__device__ __forceinline__ float ldbf(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// MODE 0 = FWD, 1 = DGRAD, 2 = WGRAD
template <int MODE>
__global__ void __launch_bounds__(256) conv_generic_kernel(ConvGeom g, const __nv_bfloat16* __restrict__ big,
                                                           const __nv_bfloat16* __restrict__ small_,
                                                           const float* __restrict__ w, void* __restrict__ out,
                                                           int64_t out_cs, const float* __restrict__ bias, int k_split) {
  __shared__ __align__(16) float As[kGK][kGT + 4];
  __shared__ __align__(16) float Bs[kGK][kGT + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int taps = g.kh * g.kw;
  const int64_t small_px = (int64_t)g.n * g.sh * g.sw, big_px = (int64_t)g.n * g.bh * g.bw;
  // tile origin
  int64_t m0;
  int n0, tap_w = 0;
  int64_t k_begin = 0, k_end;
  if (MODE == 2) {
    // grid.x = co tiles, grid.y = ci tiles, grid.z = taps * k_split
    m0 = (int64_t)blockIdx.x * kGT;
    n0 = blockIdx.y * kGT;
    tap_w = blockIdx.z / k_split;
    const int ks = blockIdx.z % k_split;
    const int64_t per = (small_px + k_split - 1) / k_split;
    k_begin = ks * per;
    k_end = min(small_px, k_begin + per);
  } else {
    m0 = (int64_t)blockIdx.x * kGT;
    n0 = blockIdx.y * kGT;
    k_end = (int64_t)taps * (MODE == 0 ? g.bc : g.sc);
  }
  const int M_ch = MODE == 2 ? g.sc : 0;                 // WGRAD: M = co
  const int N_ch = MODE == 0 ? g.sc : g.bc;               // FWD: N = co; DGRAD / WGRAD: N = ci
  const int64_t M_total = MODE == 0 ? small_px : (MODE == 1 ? big_px : g.sc);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader roles: thread loads 4 elements of the A tile and 4 of the B tile per K chunk
  const int lrow = threadIdx.x / 4;          // 0..63: M (or N) index inside the tile for pixel-major tiles
  const int lk4 = (threadIdx.x % 4) * 4;     // 0,4,8,12: K offset
  const int wk = threadIdx.x / 16;           // WGRAD: K (pixel) index 0..15
  const int wc4 = (threadIdx.x % 16) * 4;    // WGRAD: channel offset 0..60

  // FWD / DGRAD: decode this thread's M pixel once
  int pn = 0, py = 0, px = 0;
  bool m_ok = false;
  if (MODE != 2) {
    const int64_t m = m0 + lrow;
    m_ok = m < M_total;
    if (m_ok) {
      const int hh = MODE == 0 ? g.sh : g.bh, ww = MODE == 0 ? g.sw : g.bw;
      px = (int)(m % ww);
      py = (int)((m / ww) % hh);
      pn = (int)(m / ((int64_t)ww * hh));
    }
  }
  const int kc_ch = MODE == 0 ? g.bc : g.sc;   // channels per tap along K (FWD: ci, DGRAD: co)
  // 8-byte loads of 4 consecutive channels: pixel rows of both slabs start 8-byte aligned
  const bool vec4 = MODE == 2 && g.big_cs % 4 == 0 && g.small_cs % 4 == 0 &&
                    (reinterpret_cast<uintptr_t>(big) & 7) == 0 && (reinterpret_cast<uintptr_t>(small_) & 7) == 0;

  for (int64_t k0 = k_begin; k0 < k_end; k0 += kGK) {
    if (MODE == 2) {
      // A[k = pixel][m = co] = small[p][co];  B[k = pixel][n = ci] = big[p*s + tap - pad][ci]
      const int64_t p = k0 + wk;
      float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
      if (p < k_end) {
        const int ox = (int)(p % g.sw), oy = (int)((p / g.sw) % g.sh), nn = (int)(p / ((int64_t)g.sw * g.sh));
        const __nv_bfloat16* sp = small_ + p * g.small_cs;
        if (vec4 && m0 + wc4 + 3 < g.sc) {
          const uint2 u = __ldg(reinterpret_cast<const uint2*>(sp + m0 + wc4));
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
          const float2 lo = __bfloat1622float2(h[0]), hi = __bfloat1622float2(h[1]);
          a[0] = lo.x; a[1] = lo.y; a[2] = hi.x; a[3] = hi.y;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (m0 + wc4 + e < g.sc) a[e] = ldbf(sp + m0 + wc4 + e);
        }
        const int iy = oy * g.stride + tap_w / g.kw - g.pad, ix = ox * g.stride + tap_w % g.kw - g.pad;
        if (iy >= 0 && iy < g.bh && ix >= 0 && ix < g.bw) {
          const __nv_bfloat16* bp = big + (((int64_t)nn * g.bh + iy) * g.bw + ix) * g.big_cs;
          if (vec4 && n0 + wc4 + 3 < g.bc) {
            const uint2 u = __ldg(reinterpret_cast<const uint2*>(bp + n0 + wc4));
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
            const float2 lo = __bfloat1622float2(h[0]), hi = __bfloat1622float2(h[1]);
            b[0] = lo.x; b[1] = lo.y; b[2] = hi.x; b[3] = hi.y;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (n0 + wc4 + e < g.bc) b[e] = ldbf(bp + n0 + wc4 + e);
          }
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        As[wk][wc4 + e] = a[e];
        Bs[wk][wc4 + e] = b[e];
      }
    } else {
      // A[m = pixel][k = (tap, c)]: 4 consecutive channels of one tap (kc_ch % 4 is not required: per-element guards)
      float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int64_t k = k0 + lk4 + e;
        if (m_ok && k < k_end) {
          const int tap = (int)(k / kc_ch), c = (int)(k % kc_ch);
          const int ky = tap / g.kw, kx = tap % g.kw;
          if (MODE == 0) {
            const int iy = py * g.stride + ky - g.pad, ix = px * g.stride + kx - g.pad;
            if (iy >= 0 && iy < g.bh && ix >= 0 && ix < g.bw)
              a[e] = ldbf(big + (((int64_t)pn * g.bh + iy) * g.bw + ix) * g.big_cs + c);
          } else {
            const int ty2 = py + g.pad - ky, tx2 = px + g.pad - kx;
            if (ty2 >= 0 && tx2 >= 0 && ty2 % g.stride == 0 && tx2 % g.stride == 0) {
              const int oy = ty2 / g.stride, ox = tx2 / g.stride;
              if (oy < g.sh && ox < g.sw) a[e] = ldbf(small_ + (((int64_t)pn * g.sh + oy) * g.sw + ox) * g.small_cs + c);
            }
          }
        }
        As[lk4 + e][lrow] = a[e];
      }
      // B[k = (tap, c)][n]: the weight W[co][ci][ky][kx]; FWD: n = co, c = ci; DGRAD: n = ci, c = co
      float b[4] = {0.f, 0.f, 0.f, 0.f};
      const int nn = n0 + lrow;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int64_t k = k0 + lk4 + e;
        if (nn < N_ch && k < k_end) {
          const int tap = (int)(k / kc_ch), c = (int)(k % kc_ch);
          const int co = MODE == 0 ? nn : c, ci = MODE == 0 ? c : nn;
          b[e] = __ldg(w + ((int64_t)co * g.bc + ci) * taps + tap);
        }
        Bs[lk4 + e][lrow] = b[e];
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kGK; ++kk) {
      // rows are 68 floats = 17 x 16 bytes apart, so the 4-float groups are 16-byte aligned
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  if (MODE == 2) {
    float* dw = static_cast<float*>(out);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int co = (int)m0 + ty * 4 + i;
      if (co >= M_ch) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ci = n0 + tx * 4 + j;
        if (ci < g.bc) atomicAdd(dw + ((int64_t)co * g.bc + ci) * taps + tap_w, acc[i][j]);
      }
    }
  } else {
    __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t m = m0 + ty * 4 + i;
      if (m >= M_total) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = n0 + tx * 4 + j;
        if (c < N_ch) o[m * out_cs + c] = __float2bfloat16(acc[i][j] + (bias ? __ldg(bias + c) : 0.f));
      }
    }
  }
}

static int geom_check(const snb_conv_geom* d, ConvGeom* g) {
  if (!d) return fail(SNB_E_INVALID, "null geometry");
  if (d->n <= 0 || d->big_h <= 0 || d->big_w <= 0 || d->big_c <= 0 || d->small_c <= 0 || d->kh <= 0 || d->kw <= 0 ||
      d->stride <= 0 || d->pad < 0)
    return fail(SNB_E_INVALID, "bad convolution geometry");
  const int64_t sh = (d->big_h + 2 * d->pad - d->kh) / d->stride + 1, sw = (d->big_w + 2 * d->pad - d->kw) / d->stride + 1;
  if (sh != d->small_h || sw != d->small_w || sh <= 0 || sw <= 0)
    return fail(SNB_E_SHAPE, "small extent %lld x %lld does not match conv(big %lld x %lld, k %lld x %lld, s %lld, p %lld) = %lld x %lld",
                (long long)d->small_h, (long long)d->small_w, (long long)d->big_h, (long long)d->big_w, (long long)d->kh,
                (long long)d->kw, (long long)d->stride, (long long)d->pad, (long long)sh, (long long)sw);
  if (d->big_cstride < d->big_c || d->small_cstride < d->small_c) return fail(SNB_E_INVALID, "channel stride smaller than the channel count");
  if (d->n * d->big_h * d->big_w > INT32_MAX || d->big_c * d->kh * d->kw > INT32_MAX) return fail(SNB_E_UNSUPPORTED, "tensor too large");
  g->n = (int)d->n; g->bh = (int)d->big_h; g->bw = (int)d->big_w; g->bc = (int)d->big_c;
  g->sh = (int)d->small_h; g->sw = (int)d->small_w; g->sc = (int)d->small_c;
  g->kh = (int)d->kh; g->kw = (int)d->kw; g->stride = (int)d->stride; g->pad = (int)d->pad;
  g->big_cs = d->big_cstride; g->small_cs = d->small_cstride;
  return SNB_OK;
}

}  // namespace snb

using namespace snb;

extern "C" int snb_conv_generic_fwd(const snb_conv_geom* d, const void* d_big, const float* d_weight, const float* d_bias,
                                    void* d_small, void* stream) {
  ConvGeom g;
  if (int rc = geom_check(d, &g)) return rc;
  if (!d_big || !d_weight || !d_small) return fail(SNB_E_INVALID, "snb_conv_generic_fwd: null argument");
  const int64_t m = (int64_t)g.n * g.sh * g.sw;
  dim3 grid((unsigned)((m + kGT - 1) / kGT), (unsigned)((g.sc + kGT - 1) / kGT));
  conv_generic_kernel<0><<<grid, 256, 0, as_stream(stream)>>>(g, static_cast<const __nv_bfloat16*>(d_big), nullptr, d_weight,
                                                             d_small, g.small_cs, d_bias, 1);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_conv_generic_dgrad(const snb_conv_geom* d, const void* d_small, const float* d_weight, const float* d_bias,
                                      void* d_big, void* stream) {
  ConvGeom g;
  if (int rc = geom_check(d, &g)) return rc;
  if (!d_big || !d_weight || !d_small) return fail(SNB_E_INVALID, "snb_conv_generic_dgrad: null argument");
  const int64_t m = (int64_t)g.n * g.bh * g.bw;
  dim3 grid((unsigned)((m + kGT - 1) / kGT), (unsigned)((g.bc + kGT - 1) / kGT));
  conv_generic_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(g, nullptr, static_cast<const __nv_bfloat16*>(d_small), d_weight,
                                                             d_big, g.big_cs, d_bias, 1);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_conv_generic_wgrad(const snb_conv_geom* d, const void* d_big, const void* d_small, float* d_dweight,
                                      void* stream) {
  ConvGeom g;
  if (int rc = geom_check(d, &g)) return rc;
  if (!d_big || !d_small || !d_dweight) return fail(SNB_E_INVALID, "snb_conv_generic_wgrad: null argument");
  cudaStream_t st = as_stream(stream);
  const int taps = g.kh * g.kw;
  SNB_CUDA_CHECK(cudaMemsetAsync(d_dweight, 0, sizeof(float) * (size_t)g.sc * g.bc * taps, st));
  const int64_t px = (int64_t)g.n * g.sh * g.sw;
  const int tiles = ((g.sc + kGT - 1) / kGT) * ((g.bc + kGT - 1) / kGT) * taps;
  // split K (pixels): measured on B200 (tools/train_step_bench.py), more and shorter splits win up to ~8 CTAs per SM: the
  // kernel is latency-bound per CTA (no double buffering), the 64 x 64 fp32 atomics per split are not the limit
  static const int64_t min_px = [] { const char* e = std::getenv("SNB_WGRAD_MIN_PX"); return e ? (int64_t)std::atoi(e) : 256; }();
  static const int64_t ctas_per_sm = [] { const char* e = std::getenv("SNB_WGRAD_CTAS"); return e ? (int64_t)std::atoi(e) : 4; }();
  int k_split = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)sm_count() * ctas_per_sm / std::max(1, tiles), px / min_px));
  k_split = std::min(k_split, 65535 / std::max(1, taps));
  dim3 grid((unsigned)((g.sc + kGT - 1) / kGT), (unsigned)((g.bc + kGT - 1) / kGT), (unsigned)(taps * k_split));
  conv_generic_kernel<2><<<grid, 256, 0, st>>>(g, static_cast<const __nv_bfloat16*>(d_big),
                                               static_cast<const __nv_bfloat16*>(d_small), nullptr, d_dweight, 0, nullptr,
                                               k_split);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
