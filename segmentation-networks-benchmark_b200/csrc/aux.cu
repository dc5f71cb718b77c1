// Layout / pooling helpers around the convolution kernels: nn.MaxPool2d(2,2) (lib/models/unet16.py:64) on NHWC
// bf16 channel slabs and the NHWC-bf16 -> NCHW-fp32 exit conversion of the nn.Module interface.
// Pure HBM streaming: 16-byte vectors of 8 channels, one output vector per thread, x/channel fastest.
#include <cuda_bf16.h>

#include "snb_internal.h"

namespace snb {

__device__ __forceinline__ uint4 bf16x8_max(uint4 a, uint4 b) {
  uint4 r;
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}

__device__ __forceinline__ uint4 f32x4_max(uint4 a, uint4 b) {
  return make_uint4(__float_as_uint(fmaxf(__uint_as_float(a.x), __uint_as_float(b.x))),
                    __float_as_uint(fmaxf(__uint_as_float(a.y), __uint_as_float(b.y))),
                    __float_as_uint(fmaxf(__uint_as_float(a.z), __uint_as_float(b.z))),
                    __float_as_uint(fmaxf(__uint_as_float(a.w), __uint_as_float(b.w))));
}

// CV / in_sv / out_sv count 16-byte vectors (8 bf16 or 4 floats)
template <bool F32>
__global__ void maxpool2x2_kernel(const uint4* __restrict__ in, int H, int W, int CV, int in_sv,
                                  uint4* __restrict__ out, int out_sv, int64_t total) {
  const int OH = H / 2, OW = W / 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int cv = (int)(r % CV); r /= CV;
    const int ox = (int)(r % OW); r /= OW;
    const int oy = (int)(r % OH);
    const int64_t n = r / OH;
    const int64_t base = ((n * H + 2 * oy) * W + 2 * ox) * in_sv + cv;
    const uint4 a = __ldg(in + base), b = __ldg(in + base + in_sv);
    const uint4 c = __ldg(in + base + (int64_t)W * in_sv), d = __ldg(in + base + (int64_t)W * in_sv + in_sv);
    out[((n * OH + oy) * OW + ox) * out_sv + cv] =
        F32 ? f32x4_max(f32x4_max(a, b), f32x4_max(c, d)) : bf16x8_max(bf16x8_max(a, b), bf16x8_max(c, d));
  }
}

__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ in, int H, int W, int C, int cstride,
                                    float* __restrict__ out, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int x = (int)(r % W); r /= W;
    const int y = (int)(r % H); r /= H;
    const int c = (int)(r % C);
    const int64_t n = r / C;
    out[i] = __bfloat162float(in[((n * H + y) * W + x) * cstride + c]);
  }
}

// one thread = 8 channels of one pixel (16-byte vector in, 16-byte vector out)
__global__ void bn_relu_kernel(const uint4* __restrict__ in, int CV, int CPV, int in_sv, const float* __restrict__ scale,
                               const float* __restrict__ shift, uint4* __restrict__ out, int out_sv, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CPV);
    const int64_t pix = i / CPV;
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    if (cv < CV) {
      const uint4 v = __ldg(in + pix * in_sv + cv);
      const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale) + 2 * cv), s1 = __ldg(reinterpret_cast<const float4*>(scale) + 2 * cv + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(shift) + 2 * cv), b1 = __ldg(reinterpret_cast<const float4*>(shift) + 2 * cv + 1);
      const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&v);
      const float2 x0 = __bfloat1622float2(pv[0]), x1 = __bfloat1622float2(pv[1]);
      const float2 x2 = __bfloat1622float2(pv[2]), x3 = __bfloat1622float2(pv[3]);
      __nv_bfloat162 o[4];
      o[0] = __floats2bfloat162_rn(fmaxf(fmaf(x0.x, s0.x, b0.x), 0.f), fmaxf(fmaf(x0.y, s0.y, b0.y), 0.f));
      o[1] = __floats2bfloat162_rn(fmaxf(fmaf(x1.x, s0.z, b0.z), 0.f), fmaxf(fmaf(x1.y, s0.w, b0.w), 0.f));
      o[2] = __floats2bfloat162_rn(fmaxf(fmaf(x2.x, s1.x, b1.x), 0.f), fmaxf(fmaf(x2.y, s1.y, b1.y), 0.f));
      o[3] = __floats2bfloat162_rn(fmaxf(fmaf(x3.x, s1.z, b1.z), 0.f), fmaxf(fmaf(x3.y, s1.w, b1.w), 0.f));
      r = *reinterpret_cast<uint4*>(o);
    }
    out[pix * out_sv + cv] = r;
  }
}

// space-to-depth: out[n][y][x][(py*2+px)*C + c] = in[n][2y+py][2x+px][c]; a stride-2 conv3x3 over `in` becomes a
// 4-tap stride-1 conv over `out` (SNB_CONV_2X2 taps), a stride-2 conv1x1 a plain conv1x1 on channels [0, C)
__global__ void space_to_depth_kernel(const uint4* __restrict__ in, int H, int W, int CV, int in_sv,
                                      uint4* __restrict__ out, int out_sv, int64_t total) {
  const int OH = (H + 1) / 2, OW = (W + 1) / 2;     // odd sizes: the missing last row / column reads as zero
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int cv = (int)(r % CV); r /= CV;
    const int q = (int)(r % 4); r /= 4;
    const int ox = (int)(r % OW); r /= OW;
    const int oy = (int)(r % OH);
    const int64_t n = r / OH;
    const int y = 2 * oy + (q >> 1), x = 2 * ox + (q & 1);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (y < H && x < W) v = __ldg(in + ((n * H + y) * W + x) * in_sv + cv);
    out[((n * OH + oy) * OW + ox) * out_sv + q * CV + cv] = v;
  }
}

__device__ __forceinline__ uint32_t add_bf16x2(uint32_t a, uint32_t b) {
  const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&a));
  const float2 fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&b));
  const __nv_bfloat162 r = __floats2bfloat162_rn(fa.x + fb.x, fa.y + fb.y);
  return *reinterpret_cast<const uint32_t*>(&r);
}

// inverse of space_to_depth_kernel: out[n][2y+py][2x+px][c] (+)= in[n][y][x][(py*2+px)*C + c]; H, W = the OUTPUT size
// (rows / columns of the last block beyond an odd size are dropped)
__global__ void depth_to_space_kernel(const uint4* __restrict__ in, int H, int W, int CV, int in_sv, uint4* __restrict__ out,
                                      int out_sv, int accumulate, int64_t total) {
  const int IH = (H + 1) / 2, IW = (W + 1) / 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int cv = (int)(r % CV); r /= CV;
    const int x = (int)(r % W); r /= W;
    const int y = (int)(r % H);
    const int64_t n = r / H;
    const int q = (y & 1) * 2 + (x & 1);
    uint4 v = __ldg(in + ((n * IH + (y >> 1)) * IW + (x >> 1)) * in_sv + q * CV + cv);
    uint4* dst = out + ((n * H + y) * W + x) * out_sv + cv;
    if (accumulate) {
      const uint4 o = *dst;
      v = make_uint4(add_bf16x2(v.x, o.x), add_bf16x2(v.y, o.y), add_bf16x2(v.z, o.z), add_bf16x2(v.w, o.w));
    }
    *dst = v;
  }
}

// nn.MaxPool2d(kernel_size=3, stride=2, padding=1) on NHWC bf16 (torchvision resnet stem)
__global__ void maxpool3x3s2_kernel(const uint4* __restrict__ in, int H, int W, int CV, int in_sv,
                                    uint4* __restrict__ out, int out_sv, int64_t total) {
  const int OH = (H + 1) / 2, OW = (W + 1) / 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int cv = (int)(r % CV); r /= CV;
    const int ox = (int)(r % OW); r /= OW;
    const int oy = (int)(r % OH);
    const int64_t n = r / OH;
    uint4 m = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);   // bf16 -inf pairs
    for (int dy = -1; dy <= 1; ++dy) {
      const int y = 2 * oy + dy;
      if (y < 0 || y >= H) continue;
      for (int dx = -1; dx <= 1; ++dx) {
        const int x = 2 * ox + dx;
        if (x < 0 || x >= W) continue;
        m = bf16x8_max(m, __ldg(in + ((n * H + y) * W + x) * in_sv + cv));
      }
    }
    out[((n * OH + oy) * OW + ox) * out_sv + cv] = m;
  }
}

// stride-2 7x7 p3 stem as a GEMM operand: rows[n][y][x][k], k = (ky*7 + kx)*C + c for k < 49*C, zero up to KP
__global__ void stem7x7_rows_kernel(const float* __restrict__ src, int C, int H, int W, int KP, uint4* __restrict__ dst,
                                    int64_t total) {
  const int OH = (H + 1) / 2, OW = (W + 1) / 2, VPP = KP / 8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % VPP);
    int64_t r = i / VPP;
    const int ox = (int)(r % OW); r /= OW;
    const int oy = (int)(r % OH);
    const int64_t n = r / OH;
    const float* img = src + n * C * (int64_t)H * W;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = v * 8 + e;
      const int tap = k / C, c = k - tap * C;
      const int y = 2 * oy + tap / 7 - 3, x = 2 * ox + tap % 7 - 3;
      f[e] = (tap < 49 && y >= 0 && y < H && x >= 0 && x < W) ? __ldg(img + ((int64_t)c * H + y) * W + x) : 0.f;
    }
    __nv_bfloat162 p0 = __floats2bfloat162_rn(f[0], f[1]), p1 = __floats2bfloat162_rn(f[2], f[3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(f[4], f[5]), p3 = __floats2bfloat162_rn(f[6], f[7]);
    dst[i] = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                        *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
  }
}

// backward of nn.MaxPool2d(3, 2, 1) on NHWC bf16: gather formulation.  An input pixel receives the gradient of every
// window whose arg-max it is; the arg-max is recomputed with the forward's scan order (rows, then columns, strict >), so
// ties (frequent after a ReLU) go to the first maximum, as in PyTorch.
__global__ void __launch_bounds__(256) maxpool3x3s2_bwd_kernel(const uint4* __restrict__ in, int H, int W, int CV, int in_sv,
                                                               const uint4* __restrict__ dout, int dout_sv,
                                                               uint4* __restrict__ din, int din_sv, int64_t total) {
  // one thread = 8 channels of one input pixel
  const int OH = (H + 1) / 2, OW = (W + 1) / 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int cv = (int)(r % CV); r /= CV;
    const int x = (int)(r % W); r /= W;
    const int y = (int)(r % H);
    const int64_t n = r / H;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    // windows containing (y, x): rows 2oy-1 .. 2oy+1, i.e. oy = y / 2 and, for odd y, (y + 1) / 2
    for (int oy = y / 2; oy <= min(OH - 1, (y + 1) / 2); ++oy) {
      for (int ox = x / 2; ox <= min(OW - 1, (x + 1) / 2); ++ox) {
        float best[8];
        bool mine[8];
        bool first = true;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
          const int yy = 2 * oy + dy;
          if (yy < 0 || yy >= H) continue;
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            const int xx = 2 * ox + dx;
            if (xx < 0 || xx >= W) continue;
            const uint4 u = __ldg(in + ((n * H + yy) * W + xx) * in_sv + cv);
            const __nv_bfloat16* ve = reinterpret_cast<const __nv_bfloat16*>(&u);
            const bool here = yy == y && xx == x;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float v = __bfloat162float(ve[e]);
              if (first || v > best[e]) { best[e] = v; mine[e] = here; }     // strict >: the first maximum wins (PyTorch)
            }
            first = false;
          }
        }
        const uint4 g = __ldg(dout + ((n * OH + oy) * OW + ox) * dout_sv + cv);
        const __nv_bfloat16* ge = reinterpret_cast<const __nv_bfloat16*>(&g);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += mine[e] ? __bfloat162float(ge[e]) : 0.f;
      }
    }
    __nv_bfloat162 o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = __floats2bfloat162_rn(acc[2 * e], acc[2 * e + 1]);
    din[((n * H + y) * W + x) * din_sv + cv] = *reinterpret_cast<const uint4*>(o);
  }
}

// elementwise helpers of the backward pass on bf16 slabs: mode 0: out = a + b;  mode 1: out = a * (b > 0 ? 1 : slope)
// (the gradient through a leaky-ReLU whose OUTPUT is b)
__global__ void __launch_bounds__(256) ew_nhwc_kernel(const uint4* __restrict__ a, int a_sv, const uint4* __restrict__ b, int b_sv,
                                                      uint4* __restrict__ out, int out_sv, int CV, int mode, float slope,
                                                      int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    const int64_t pix = i / CV;
    const uint4 ua = __ldg(a + pix * a_sv + cv), ub = __ldg(b + pix * b_sv + cv);
    const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&ua);
    const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&ub);
    __nv_bfloat162 o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 av = __bfloat1622float2(pa[e]), bv = __bfloat1622float2(pb[e]);
      o[e] = mode == 0 ? __floats2bfloat162_rn(av.x + bv.x, av.y + bv.y)
                       : __floats2bfloat162_rn(bv.x > 0.f ? av.x : av.x * slope, bv.y > 0.f ? av.y : av.y * slope);
    }
    out[pix * out_sv + cv] = *reinterpret_cast<const uint4*>(o);
  }
}

static int aux_grid(int64_t total) {
  const int64_t need = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

}  // namespace snb

using namespace snb;

extern "C" int snb_maxpool2x2(const void* d_in, int64_t n, int64_t h, int64_t w, int64_t channels, int64_t in_cstride,
                              void* d_out, int64_t out_cstride, int elem_bytes, void* stream) {
  if (elem_bytes != 2 && elem_bytes != 4) return fail(SNB_E_INVALID, "elem_bytes must be 2 (bf16) or 4 (float)");
  const int64_t V = 16 / elem_bytes;   // elements per 16-byte vector
  if (!d_in || !d_out) return fail(SNB_E_INVALID, "snb_maxpool2x2: null argument");
  if (n <= 0 || h <= 0 || w <= 0 || (h & 1) || (w & 1)) return fail(SNB_E_INVALID, "maxpool needs even positive h, w");
  if (channels <= 0 || channels % V || in_cstride % V || out_cstride % V || in_cstride < channels || out_cstride < channels)
    return fail(SNB_E_INVALID, "channel counts and strides must be multiples of 16 bytes");
  if ((reinterpret_cast<uintptr_t>(d_in) & 15) || (reinterpret_cast<uintptr_t>(d_out) & 15))
    return fail(SNB_E_INVALID, "pointers must be 16-byte aligned");
  const int64_t total = n * (h / 2) * (w / 2) * (channels / V);
  if (elem_bytes == 4)
    maxpool2x2_kernel<true><<<aux_grid(total), 256, 0, as_stream(stream)>>>(
        static_cast<const uint4*>(d_in), (int)h, (int)w, (int)(channels / V), (int)(in_cstride / V),
        static_cast<uint4*>(d_out), (int)(out_cstride / V), total);
  else
    maxpool2x2_kernel<false><<<aux_grid(total), 256, 0, as_stream(stream)>>>(
        static_cast<const uint4*>(d_in), (int)h, (int)w, (int)(channels / V), (int)(in_cstride / V),
        static_cast<uint4*>(d_out), (int)(out_cstride / V), total);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

// Multi-segment index gather: dst[i] = idx[i] >= 0 ? src[idx[i]] : 0 for every segment of a device table, ONE launch for all
// layers of a plan.  Used to (re)pack fp32 parameters into the bf16 [phase][tap][Cout][Cin] operands of the forward and
// input-gradient convolutions after an optimiser step, and to scatter the packed fp32 weight gradients back into the
// parameters' own layouts.  A block handles 1024 consecutive elements of one segment.
namespace snb {
template <bool DST_BF16>
__global__ void __launch_bounds__(256) gather_segments_kernel(const snb_gather_seg* __restrict__ segs, int n_segs) {
  int lo = 0, hi = n_segs - 1;
  while (lo < hi) {            // last segment whose first_block <= blockIdx.x
    const int mid = (lo + hi + 1) >> 1;
    if (segs[mid].first_block <= (int64_t)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const snb_gather_seg sg = segs[lo];
  const int64_t base = ((int64_t)blockIdx.x - sg.first_block) * 1024;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t i = base + k * 256 + threadIdx.x;
    if (i < sg.count) {
      const int32_t j = __ldg(sg.idx + i);
      const float v = j >= 0 ? __ldg(sg.src + j) : 0.f;
      if (DST_BF16) static_cast<__nv_bfloat16*>(sg.dst)[i] = __float2bfloat16(v);
      else static_cast<float*>(sg.dst)[i] = v;
    }
  }
}
}  // namespace snb

extern "C" int snb_gather_segments(const snb_gather_seg* d_segs, int64_t n_segs, int64_t total_blocks, int dst_bf16,
                                   void* stream) {
  if (!d_segs || n_segs <= 0 || total_blocks <= 0 || total_blocks > INT32_MAX || n_segs > INT32_MAX)
    return snb::fail(SNB_E_INVALID, "snb_gather_segments: bad arguments");
  if (dst_bf16) snb::gather_segments_kernel<true><<<(unsigned)total_blocks, 256, 0, snb::as_stream(stream)>>>(d_segs, (int)n_segs);
  else snb::gather_segments_kernel<false><<<(unsigned)total_blocks, 256, 0, snb::as_stream(stream)>>>(d_segs, (int)n_segs);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

// out[n][y][x][c] *= scale[n][c]  (nn.Dropout2d with a given keep mask / (1 - p), lib/models/linknet.py:57,83; the same
// kernel is its backward)
namespace snb {
__global__ void __launch_bounds__(256) scale_nc_kernel(const uint4* __restrict__ in, int in_sv, const float* __restrict__ scale,
                                                       int64_t hw, int CV, uint4* __restrict__ out, int out_sv, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    const int64_t pix = i / CV;
    const int64_t n = pix / hw;
    const uint4 u = __ldg(in + pix * in_sv + cv);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + n * CV * 8) + 2 * cv);
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(scale + n * CV * 8) + 2 * cv + 1);
    const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&u);
    const float2 a = __bfloat1622float2(pv[0]), b = __bfloat1622float2(pv[1]), c = __bfloat1622float2(pv[2]), d = __bfloat1622float2(pv[3]);
    __nv_bfloat162 o[4] = {__floats2bfloat162_rn(a.x * s0.x, a.y * s0.y), __floats2bfloat162_rn(b.x * s0.z, b.y * s0.w),
                           __floats2bfloat162_rn(c.x * s1.x, c.y * s1.y), __floats2bfloat162_rn(d.x * s1.z, d.y * s1.w)};
    out[pix * out_sv + cv] = *reinterpret_cast<const uint4*>(o);
  }
}
}  // namespace snb

static int check_bf16_slabs(const void* d_in, const void* d_out, int64_t n, int64_t h, int64_t w, int64_t channels,
                            int64_t in_cstride, int64_t out_cstride, int64_t out_channels) {
  if (!d_in || !d_out) return fail(SNB_E_INVALID, "null argument");
  if (n <= 0 || h <= 0 || w <= 0) return fail(SNB_E_INVALID, "bad shape");
  if (channels <= 0 || channels % 8 || in_cstride % 8 || out_cstride % 8 || in_cstride < channels || out_cstride < out_channels)
    return fail(SNB_E_INVALID, "channel counts and strides must be multiples of 8");
  if ((reinterpret_cast<uintptr_t>(d_in) & 15) || (reinterpret_cast<uintptr_t>(d_out) & 15))
    return fail(SNB_E_INVALID, "pointers must be 16-byte aligned");
  return SNB_OK;
}

extern "C" int snb_space_to_depth2(const void* d_in, int64_t n, int64_t h, int64_t w, int64_t channels, int64_t in_cstride,
                                   void* d_out, int64_t out_cstride, void* stream) {
  if (int rc = check_bf16_slabs(d_in, d_out, n, h, w, channels, in_cstride, out_cstride, 4 * channels)) return rc;
  const int64_t total = n * ((h + 1) / 2) * ((w + 1) / 2) * 4 * (channels / 8);
  space_to_depth_kernel<<<aux_grid(total), 256, 0, as_stream(stream)>>>(static_cast<const uint4*>(d_in), (int)h, (int)w,
                                                                         (int)(channels / 8), (int)(in_cstride / 8),
                                                                         static_cast<uint4*>(d_out), (int)(out_cstride / 8), total);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_depth_to_space2(const void* d_in, int64_t n, int64_t h, int64_t w, int64_t channels, int64_t in_cstride,
                                   void* d_out, int64_t out_cstride, int accumulate, void* stream) {
  // (h, w, channels) describe the OUTPUT; the input holds ceil(h/2) x ceil(w/2) blocks of 4 * channels
  if (int rc = check_bf16_slabs(d_out, d_in, n, h, w, channels, out_cstride, in_cstride, 4 * channels)) return rc;
  const int64_t total = n * h * w * (channels / 8);
  depth_to_space_kernel<<<aux_grid(total), 256, 0, as_stream(stream)>>>(static_cast<const uint4*>(d_in), (int)h, (int)w,
                                                                         (int)(channels / 8), (int)(in_cstride / 8),
                                                                         static_cast<uint4*>(d_out), (int)(out_cstride / 8),
                                                                         accumulate, total);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_scale_nc_nhwc(const void* d_in, int64_t n, int64_t hw, int64_t channels, int64_t in_cstride,
                                 const float* d_scale, void* d_out, int64_t out_cstride, void* stream) {
  if (int rc = check_bf16_slabs(d_in, d_out, n, hw, 1, channels, in_cstride, out_cstride, channels)) return rc;
  if (!d_scale || (reinterpret_cast<uintptr_t>(d_scale) & 15)) return fail(SNB_E_INVALID, "scale must be a 16-byte aligned float[n][channels]");
  const int64_t total = n * hw * (channels / 8);
  scale_nc_kernel<<<aux_grid(total), 256, 0, as_stream(stream)>>>(static_cast<const uint4*>(d_in), (int)(in_cstride / 8), d_scale, hw,
                                                                   (int)(channels / 8), static_cast<uint4*>(d_out),
                                                                   (int)(out_cstride / 8), total);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_maxpool3x3s2(const void* d_in, int64_t n, int64_t h, int64_t w, int64_t channels, int64_t in_cstride,
                                void* d_out, int64_t out_cstride, void* stream) {
  if (int rc = check_bf16_slabs(d_in, d_out, n, h, w, channels, in_cstride, out_cstride, channels)) return rc;
  const int64_t total = n * ((h + 1) / 2) * ((w + 1) / 2) * (channels / 8);
  maxpool3x3s2_kernel<<<aux_grid(total), 256, 0, as_stream(stream)>>>(static_cast<const uint4*>(d_in), (int)h, (int)w,
                                                                       (int)(channels / 8), (int)(in_cstride / 8),
                                                                       static_cast<uint4*>(d_out), (int)(out_cstride / 8), total);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_stem7x7_rows(const float* d_src, int64_t n, int64_t channels, int64_t h, int64_t w, void* d_dst,
                                int64_t k_pad, void* stream) {
  if (!d_src || !d_dst) return fail(SNB_E_INVALID, "snb_stem7x7_rows: null argument");
  if (n <= 0 || h <= 0 || w <= 0 || channels <= 0) return fail(SNB_E_INVALID, "bad shape");
  if (k_pad < 49 * channels || k_pad % 32) return fail(SNB_E_INVALID, "k_pad must be a multiple of 32 covering 49 * channels");
  if (reinterpret_cast<uintptr_t>(d_dst) & 15) return fail(SNB_E_INVALID, "destination must be 16-byte aligned");
  const int64_t total = n * ((h + 1) / 2) * ((w + 1) / 2) * (k_pad / 8);
  stem7x7_rows_kernel<<<aux_grid(total), 256, 0, as_stream(stream)>>>(d_src, (int)channels, (int)h, (int)w, (int)k_pad,
                                                                       static_cast<uint4*>(d_dst), total);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_bn_relu_nhwc(const void* d_in, int64_t n, int64_t h, int64_t w, int64_t channels, int64_t in_cstride,
                                const float* d_scale, const float* d_shift, void* d_out, int64_t channels_pad,
                                int64_t out_cstride, void* stream) {
  if (!d_in || !d_scale || !d_shift || !d_out) return fail(SNB_E_INVALID, "snb_bn_relu_nhwc: null argument");
  if (n <= 0 || h <= 0 || w <= 0) return fail(SNB_E_INVALID, "bad shape");
  if (channels <= 0 || channels % 8 || channels_pad < channels || channels_pad % 8 || in_cstride % 8 || out_cstride % 8 ||
      in_cstride < channels || out_cstride < channels_pad)
    return fail(SNB_E_INVALID, "channel counts and strides must be multiples of 8");
  if ((reinterpret_cast<uintptr_t>(d_in) & 15) || (reinterpret_cast<uintptr_t>(d_out) & 15) ||
      (reinterpret_cast<uintptr_t>(d_scale) & 15) || (reinterpret_cast<uintptr_t>(d_shift) & 15))
    return fail(SNB_E_INVALID, "pointers must be 16-byte aligned");
  const int64_t total = n * h * w * (channels_pad / 8);
  bn_relu_kernel<<<aux_grid(total), 256, 0, as_stream(stream)>>>(static_cast<const uint4*>(d_in), (int)(channels / 8),
                                                                  (int)(channels_pad / 8), (int)(in_cstride / 8), d_scale,
                                                                  d_shift, static_cast<uint4*>(d_out), (int)(out_cstride / 8), total);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_nhwc_bf16_to_nchw_f32(const void* d_in, int64_t n, int64_t h, int64_t w, int64_t channels,
                                         int64_t in_cstride, float* d_out, void* stream) {
  if (!d_in || !d_out) return fail(SNB_E_INVALID, "snb_nhwc_bf16_to_nchw_f32: null argument");
  if (n <= 0 || h <= 0 || w <= 0 || channels <= 0 || in_cstride < channels) return fail(SNB_E_INVALID, "bad shape");
  const int64_t total = n * channels * h * w;
  nhwc_to_nchw_kernel<<<aux_grid(total), 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(d_in), (int)h, (int)w,
                                                                       (int)channels, (int)in_cstride, d_out, total);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_maxpool3x3s2_backward(const void* d_in, int64_t n, int64_t h, int64_t w, int64_t channels, int64_t in_cstride,
                                         const void* d_dout, int64_t dout_cstride, void* d_din, int64_t din_cstride,
                                         void* stream) {
  if (!d_in || !d_dout || !d_din) return fail(SNB_E_INVALID, "snb_maxpool3x3s2_backward: null argument");
  if (n <= 0 || h <= 0 || w <= 0 || channels <= 0 || in_cstride < channels || dout_cstride < channels || din_cstride < channels)
    return fail(SNB_E_INVALID, "bad shape");
  if (channels % 8 || in_cstride % 8 || dout_cstride % 8 || din_cstride % 8 || (reinterpret_cast<uintptr_t>(d_in) & 15) ||
      (reinterpret_cast<uintptr_t>(d_dout) & 15) || (reinterpret_cast<uintptr_t>(d_din) & 15))
    return fail(SNB_E_INVALID, "channel counts / strides must be multiples of 8 and pointers 16-byte aligned");
  const int64_t total = n * h * w * (channels / 8);
  maxpool3x3s2_bwd_kernel<<<aux_grid(total), 256, 0, as_stream(stream)>>>(
      static_cast<const uint4*>(d_in), (int)h, (int)w, (int)(channels / 8), (int)(in_cstride / 8),
      static_cast<const uint4*>(d_dout), (int)(dout_cstride / 8), static_cast<uint4*>(d_din), (int)(din_cstride / 8), total);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_ew_nhwc(const void* d_a, int64_t a_cstride, const void* d_b, int64_t b_cstride, void* d_out,
                           int64_t out_cstride, int64_t pixels, int64_t channels, int mode, float slope, void* stream) {
  if (!d_a || !d_b || !d_out) return fail(SNB_E_INVALID, "snb_ew_nhwc: null argument");
  if (pixels <= 0 || channels <= 0 || a_cstride < channels || b_cstride < channels || out_cstride < channels || mode < 0 || mode > 1)
    return fail(SNB_E_INVALID, "bad shape or mode");
  if (channels % 8 || a_cstride % 8 || b_cstride % 8 || out_cstride % 8 || (reinterpret_cast<uintptr_t>(d_a) & 15) ||
      (reinterpret_cast<uintptr_t>(d_b) & 15) || (reinterpret_cast<uintptr_t>(d_out) & 15))
    return fail(SNB_E_INVALID, "channel counts / strides must be multiples of 8 and pointers 16-byte aligned");
  const int64_t total = pixels * (channels / 8);
  ew_nhwc_kernel<<<aux_grid(total), 256, 0, as_stream(stream)>>>(static_cast<const uint4*>(d_a), (int)(a_cstride / 8),
                                                                 static_cast<const uint4*>(d_b), (int)(b_cstride / 8),
                                                                 static_cast<uint4*>(d_out), (int)(out_cstride / 8),
                                                                 (int)(channels / 8), mode, slope, total);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
