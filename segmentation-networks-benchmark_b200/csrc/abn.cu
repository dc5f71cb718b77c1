// InPlaceABN (in-place activated batch norm) on NCHW float32 tensors: the operator lib/modules/abn/functions.py:62-122
// drives through the external `inplace_abn` extension (mapillary/inplace_abn; not vendored in the reference, no pinned
// version -> parity unpinned at that boundary).  The arithmetic follows that library's published kernels:
//   mean_var   : per-channel mean and BIASED variance over N x H x W
//   forward    : y = (x - mean) / sqrt(var + eps) * (|weight| + eps) + bias, then the activation, in place
//   edz_eydz   : after undoing the activation (z -> y, dz -> dy): edz = sum(dy), eydz = sum(xhat * dy),
//                xhat = (y - bias) / (|weight| + eps)
//   backward   : dx = (dy - edz / count - xhat * eydz / count) * (|weight| + eps) / sqrt(var + eps),
//                dweight = eydz * sign(weight), dbias = edz     (eval mode: edz = eydz = 0 in dx, functions.py:110-112)
// and the reference's own glue (functions.py:77-87): running_mean/var momentum update with the unbiased variance.
// Everything is HBM-bound: per element 4 B read + 4 B written (forward), 8 B read + 4 B written (backward), plus one
// 4 B (8 B) reduction pass in training mode.  One CTA per (channel, image-slab) slice, float4 loads when H*W % 4 == 0,
// warp-shuffle + shared-memory block reduction, float64 atomics for the cross-CTA sums.
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdint>

#include "snb_internal.h"

namespace snb {

constexpr int kAbnThreads = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of two doubles; result valid in thread 0
__device__ __forceinline__ void block_sum2(double& a, double& b) {
  __shared__ double sa[kAbnThreads / 32], sb[kAbnThreads / 32];
  a = warp_sum(a);
  b = warp_sum(b);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sa[w] = a; sb[w] = b; }
  __syncthreads();
  if (w == 0) {
    a = l < kAbnThreads / 32 ? sa[l] : 0.0;
    b = l < kAbnThreads / 32 ? sb[l] : 0.0;
    a = warp_sum(a);
    b = warp_sum(b);
  }
}

// activation codes: 0 none, 1 leaky_relu(slope), 2 elu
__device__ __forceinline__ float act_fwd(float y, int act, float slope) {
  if (act == 1) return y >= 0.f ? y : y * slope;
  if (act == 2) return y >= 0.f ? y : expm1f(y);
  return y;
}
// undo the activation: z -> y (pre-activation value), dz -> dy (leaky_relu_backward / elu_backward of the backend)
__device__ __forceinline__ void act_bwd(float& z, float& dz, int act, float slope) {
  if (act == 1) {
    if (z < 0.f) { dz *= slope; z /= slope; }
  } else if (act == 2) {
    if (z < 0.f) { dz *= z + 1.f; z = log1pf(z); }
  }
}

// grid = (C, slabs): CTA (c, s) reduces images n = s, s + slabs, ... of channel c.
// sums[2c], sums[2c+1] += (sum (x - p), sum (x - p)^2) with the pivot p = the channel's first value: the variance
// E[(x-p)^2] - E[x-p]^2 then keeps its bits when |mean| >> std (E[x^2] - E[x]^2 cancels catastrophically there).
__global__ void __launch_bounds__(kAbnThreads) abn_stats_kernel(const float* __restrict__ x, int N, int C, int64_t HW,
                                                                 double* __restrict__ sums) {
  const int c = blockIdx.x;
  const float pv = __ldg(x + (int64_t)c * HW);
  double s = 0.0, q = 0.0;
  for (int n = blockIdx.y; n < N; n += gridDim.y) {
    const float* p = x + ((int64_t)n * C + c) * HW;
    if ((HW & 3) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      const float4* p4 = reinterpret_cast<const float4*>(p);
      float fs = 0.f, fq = 0.f;   // float partials per thread over <= HW/1024 elements, folded into double below
      int cnt = 0;
      for (int64_t i = threadIdx.x; i < HW / 4; i += kAbnThreads) {
        float4 v = __ldg(p4 + i);
        v.x -= pv; v.y -= pv; v.z -= pv; v.w -= pv;
        fs += (v.x + v.y) + (v.z + v.w);
        fq += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        if (++cnt == 16) { s += fs; q += fq; fs = fq = 0.f; cnt = 0; }
      }
      s += fs;
      q += fq;
    } else {
      for (int64_t i = threadIdx.x; i < HW; i += kAbnThreads) {
        const double v = (double)__ldg(p + i) - (double)pv;
        s += v;
        q += v * v;
      }
    }
  }
  block_sum2(s, q);
  if (threadIdx.x == 0) {
    atomicAdd(sums + 2 * c, s);
    atomicAdd(sums + 2 * c + 1, q);
  }
}

// mean / biased var from the sums; running stats update of functions.py:84-85 (unbiased variance, momentum)
__global__ void abn_finalize_stats_kernel(const double* __restrict__ sums, const float* __restrict__ x, int64_t HW, int C,
                                          double count, float momentum, float* __restrict__ mean, float* __restrict__ var,
                                          float* __restrict__ running_mean, float* __restrict__ running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double ms = sums[2 * c] / count;                       // mean of (x - pivot)
  double v = sums[2 * c + 1] / count - ms * ms;
  v = v < 0.0 ? 0.0 : v;
  const double m = ms + (double)__ldg(x + (int64_t)c * HW);
  mean[c] = (float)m;
  var[c] = (float)v;
  if (running_mean) running_mean[c] = running_mean[c] * (1.f - momentum) + momentum * (float)m;
  if (running_var) running_var[c] = running_var[c] * (1.f - momentum) + (float)(momentum * v * count / (count - 1.0));
}

// grid = (N * C planes, chunks of the plane)
__global__ void __launch_bounds__(kAbnThreads) abn_forward_kernel(float* __restrict__ x, int C, int64_t HW,
                                                                   const float* __restrict__ mean,
                                                                   const float* __restrict__ var,
                                                                   const float* __restrict__ weight,
                                                                   const float* __restrict__ bias, float eps, int act,
                                                                   float slope) {
  const int64_t plane = blockIdx.x;
  const int c = (int)(plane % C);
  const float g = weight ? fabsf(__ldg(weight + c)) + eps : 1.f;
  const float b = bias ? __ldg(bias + c) : 0.f;
  const float m = __ldg(mean + c);
  const float mul = g / sqrtf(__ldg(var + c) + eps);   // (x - mean) * invstd * gamma + beta, one rounding of the scale
  float* p = x + plane * HW;
  if ((HW & 3) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    float4* p4 = reinterpret_cast<float4*>(p);
    for (int64_t i = blockIdx.y * (int64_t)kAbnThreads + threadIdx.x; i < HW / 4; i += (int64_t)gridDim.y * kAbnThreads) {
      float4 v = p4[i];
      v.x = act_fwd((v.x - m) * mul + b, act, slope);
      v.y = act_fwd((v.y - m) * mul + b, act, slope);
      v.z = act_fwd((v.z - m) * mul + b, act, slope);
      v.w = act_fwd((v.w - m) * mul + b, act, slope);
      p4[i] = v;
    }
  } else {
    for (int64_t i = blockIdx.y * (int64_t)kAbnThreads + threadIdx.x; i < HW; i += (int64_t)gridDim.y * kAbnThreads)
      p[i] = act_fwd((p[i] - m) * mul + b, act, slope);
  }
}

// sums[2c] += sum(dy), sums[2c+1] += sum(xhat * dy) over the slabs of channel c
__global__ void __launch_bounds__(kAbnThreads) abn_bwd_reduce_kernel(const float* __restrict__ z,
                                                                      const float* __restrict__ dz, int N, int C,
                                                                      int64_t HW, const float* __restrict__ weight,
                                                                      const float* __restrict__ bias, float eps, int act,
                                                                      float slope, double* __restrict__ sums) {
  const int c = blockIdx.x;
  const float g = weight ? fabsf(__ldg(weight + c)) + eps : 1.f;
  const float b = bias ? __ldg(bias + c) : 0.f;
  double s = 0.0, q = 0.0;
  for (int n = blockIdx.y; n < N; n += gridDim.y) {
    const int64_t base = ((int64_t)n * C + c) * HW;
    float fs = 0.f, fq = 0.f;
    int cnt = 0;
    for (int64_t i = threadIdx.x; i < HW; i += kAbnThreads) {
      float zz = __ldg(z + base + i), dd = __ldg(dz + base + i);
      act_bwd(zz, dd, act, slope);
      fs += dd;
      fq += (zz - b) / g * dd;
      if (++cnt == 32) { s += fs; q += fq; fs = fq = 0.f; cnt = 0; }
    }
    s += fs;
    q += fq;
  }
  block_sum2(s, q);
  if (threadIdx.x == 0) {
    atomicAdd(sums + 2 * c, s);
    atomicAdd(sums + 2 * c + 1, q);
  }
}

// dx (may alias dz); z is restored to the pre-activation value only in registers
__global__ void __launch_bounds__(kAbnThreads) abn_backward_kernel(const float* __restrict__ z, const float* dz, int C,
                                                                    int64_t HW, const float* __restrict__ var,
                                                                    const float* __restrict__ weight,
                                                                    const float* __restrict__ bias,
                                                                    const double* __restrict__ sums, double count,
                                                                    int training, float eps, int act, float slope,
                                                                    float* dx) {
  const int64_t plane = blockIdx.x;
  const int c = (int)(plane % C);
  const float g = weight ? fabsf(__ldg(weight + c)) + eps : 1.f;
  const float b = bias ? __ldg(bias + c) : 0.f;
  const float mul = g / sqrtf(__ldg(var + c) + eps);
  const float edz = training ? (float)(sums[2 * c] / count) : 0.f;
  const float eydz = training ? (float)(sums[2 * c + 1] / count) : 0.f;
  const int64_t base = plane * HW;
  for (int64_t i = blockIdx.y * (int64_t)kAbnThreads + threadIdx.x; i < HW; i += (int64_t)gridDim.y * kAbnThreads) {
    float zz = __ldg(z + base + i), dd = dz[base + i];
    act_bwd(zz, dd, act, slope);
    const float xhat = (zz - b) / g;
    dx[base + i] = (dd - edz - xhat * eydz) * mul;
  }
}

__global__ void abn_param_grads_kernel(const double* __restrict__ sums, const float* __restrict__ weight, int C,
                                       float* __restrict__ dweight, float* __restrict__ dbias) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (dweight) dweight[c] = (float)(__ldg(weight + c) > 0.f ? sums[2 * c + 1] : -sums[2 * c + 1]);
  if (dbias) dbias[c] = (float)sums[2 * c];
}

// ---- training-mode BatchNorm2d / InPlaceABN on the engine's NHWC bf16 slabs (LinkNet34 train-mode forward) ----------
// One thread = 8 channels (one 16-byte vector) of a strided set of pixels; per-thread float partials, shared-memory
// float atomics across the pixel lanes of a CTA, then one float64 atomic per channel and CTA.
constexpr int kBnMaxCV = 256;   // <= 2048 channels

// What the LAST block of a statistics kernel does with the finished per-channel sums (one launch instead of memset +
// reduce + finalize): the block that draws the last ticket reads the sums, writes the derived per-channel values, and puts
// the workspace back to its "zero at rest" state.  Workspace = double[2 C] sums, then a 64-bit ticket word, then
// float[2 C] reduced values for the backward apply kernel: double[3 C + 2] in total, zeroed ONCE by the caller.
struct BnFinalize {
  int mode;                  // 1: BatchNorm forward (scale / shift / statistics), 2: channel sums, 3: backward (means + dgamma / dbeta)
  int C, abn;
  double count;
  float eps, momentum;
  const __nv_bfloat16* pivot;
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  float* scale;
  float* shift;
  float* mean_out;
  float* var_out;
  float* out0;               // mode 2: sums; mode 3: dgamma
  float* out1;               // mode 3: dbeta
};

__device__ __forceinline__ void bn_last_block_finalize(double* __restrict__ sums, const BnFinalize& f) {
  __shared__ unsigned int s_last;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(sums + 2 * f.C);
  float* red = reinterpret_cast<float*>(sums + 2 * f.C + 1);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int c = threadIdx.x; c < f.C; c += blockDim.x) {
    const double s0 = __ldcg(sums + 2 * c), s1 = __ldcg(sums + 2 * c + 1);
    sums[2 * c] = 0.0;
    sums[2 * c + 1] = 0.0;
    if (f.mode == 1) {
      // mean / biased variance -> fused scale and shift of the normalisation; running statistics as nn.BatchNorm2d and
      // functions.py:84-85 update them (momentum, unbiased variance).  abn != 0: gamma = |weight| + eps (InPlaceABN backend)
      const double ms = s0 / f.count;                              // mean of (x - pivot), pivot = pixel 0
      double v = s1 / f.count - ms * ms;
      v = v < 0.0 ? 0.0 : v;
      const double m = ms + (double)__bfloat162float(f.pivot[c]);
      const float g = f.gamma ? (f.abn ? fabsf(f.gamma[c]) + f.eps : f.gamma[c]) : 1.f;
      const float sc = g * rsqrtf((float)v + f.eps);
      f.scale[c] = sc;
      f.shift[c] = (f.beta ? f.beta[c] : 0.f) - (float)m * sc;
      if (f.mean_out) f.mean_out[c] = (float)m;
      if (f.var_out) f.var_out[c] = (float)v;
      if (f.running_mean) f.running_mean[c] = f.running_mean[c] * (1.f - f.momentum) + f.momentum * (float)m;
      if (f.running_var && f.count > 1.0)
        f.running_var[c] = f.running_var[c] * (1.f - f.momentum) + (float)(f.momentum * v * f.count / (f.count - 1.0));
    } else if (f.mode == 2) {
      f.out0[c] = (float)s0;
    } else {
      red[2 * c] = (float)(s0 / f.count);                          // mean dz, mean dz * xhat: read by the apply kernel
      red[2 * c + 1] = (float)(s1 / f.count);
      if (f.out0) f.out0[c] = (float)((f.abn && f.gamma && f.gamma[c] <= 0.f) ? -s1 : s1);
      if (f.out1) f.out1[c] = (float)s0;
    }
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

// PIVOT: accumulate (x - p), (x - p)^2 with p = the channel's value at pixel 0 (see abn_stats_kernel)
template <bool PIVOT>
__global__ void __launch_bounds__(256) bn_stats_nhwc_kernel(const uint4* __restrict__ in, int CV, int in_sv, int64_t pixels,
                                                            double* __restrict__ sums, const BnFinalize fin) {
  __shared__ float sh[kBnMaxCV * 16];          // [pixel lane][channel][sum, sum of squares]: blockDim.x * 16 floats
  const int v = threadIdx.x % CV, pl = threadIdx.x / CV, ppb = blockDim.x / CV;
  float s[8], q[8], pvt[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = q[e] = pvt[e] = 0.f;
  if (PIVOT && pl < ppb) {
    const uint4 u = __ldg(in + v);
    const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 x = __bfloat1622float2(pv[e]);
      pvt[2 * e] = x.x; pvt[2 * e + 1] = x.y;
    }
  }
  if (pl < ppb) {
    for (int64_t pix = blockIdx.x * (int64_t)ppb + pl; pix < pixels; pix += (int64_t)gridDim.x * ppb) {
      const uint4 u = __ldg(in + pix * in_sv + v);
      const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 x = __bfloat1622float2(pv[e]);
        x.x -= pvt[2 * e]; x.y -= pvt[2 * e + 1];
        s[2 * e] += x.x; s[2 * e + 1] += x.y;
        q[2 * e] += x.x * x.x; q[2 * e + 1] += x.y * x.y;
      }
    }
  }
  // fixed-order fold over the pixel lanes of the CTA (no shared-memory atomics: the same input gives the same statistics,
  // so a train-mode forward is reproducible run to run), then one float64 atomic per channel and CTA
  if (pl < ppb) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sh[pl * CV * 16 + (v * 8 + e) * 2] = s[e];
      sh[pl * CV * 16 + (v * 8 + e) * 2 + 1] = q[e];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < CV * 16; i += blockDim.x) {
    float t = 0.f;
    for (int k = 0; k < ppb; ++k) t += sh[k * CV * 16 + i];
    atomicAdd(sums + i, (double)t);
  }
  bn_last_block_finalize(sums, fin);
}

// out = act(x * scale + shift (+ residual before the activation)) (+ residual after it); act: slope >= 0 -> leaky-ReLU
// with that slope (0 = ReLU), slope < 0 -> identity
__global__ void __launch_bounds__(256) bn_apply_nhwc_kernel(const uint4* __restrict__ in, int CV, int in_sv,
                                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                                            float slope, const uint4* __restrict__ res, int res_sv,
                                                            int res_after_act, uint4* __restrict__ out, int out_sv,
                                                            int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    const int64_t pix = i / CV;
    const uint4 u = __ldg(in + pix * in_sv + cv);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale) + 2 * cv), s1 = __ldg(reinterpret_cast<const float4*>(scale) + 2 * cv + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(shift) + 2 * cv), b1 = __ldg(reinterpret_cast<const float4*>(shift) + 2 * cv + 1);
    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const float sh[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&u);
    float f[8], r[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 x = __bfloat1622float2(pv[e]);
      f[2 * e] = fmaf(x.x, sc[2 * e], sh[2 * e]);
      f[2 * e + 1] = fmaf(x.y, sc[2 * e + 1], sh[2 * e + 1]);
      r[2 * e] = r[2 * e + 1] = 0.f;
    }
    if (res) {
      const uint4 ru = __ldg(res + pix * res_sv + cv);
      const __nv_bfloat162* pr = reinterpret_cast<const __nv_bfloat162*>(&ru);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 x = __bfloat1622float2(pr[e]);
        r[2 * e] = x.x; r[2 * e + 1] = x.y;
      }
    }
    __nv_bfloat162 o[4];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float y = f[e] + (res_after_act ? 0.f : r[e]);
      if (slope >= 0.f) y = y > 0.f ? y : y * slope;
      f[e] = y + (res_after_act ? r[e] : 0.f);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
    out[pix * out_sv + cv] = *reinterpret_cast<uint4*>(o);
  }
}

// ---- backward of snb_bn_train_nhwc: out = act(x * scale + shift + r_before) + r_after ---------------------------------
// dz = dOut * act'(u), u recomputed exactly as the forward computed it (same fma on the same bf16 inputs, so the mask
// is identical); sums[2c] += sum dz, sums[2c+1] += sum dz * xhat, xhat = (x - mean) * rsqrt(var + eps)
__global__ void __launch_bounds__(256) bn_bwd_reduce_nhwc_kernel(const uint4* __restrict__ x, int CV, int x_sv,
                                                                 const uint4* __restrict__ g, int g_sv,
                                                                 const uint4* __restrict__ rb, int rb_sv,
                                                                 const float* __restrict__ scale, const float* __restrict__ shift,
                                                                 const float* __restrict__ mean, const float* __restrict__ var,
                                                                 float eps, float slope, int64_t pixels,
                                                                 double* __restrict__ sums, const BnFinalize fin) {
  __shared__ float sh[kBnMaxCV * 16];          // [pixel lane][channel][sum dz, sum dz * xhat]: blockDim.x * 16 floats
  const int v = threadIdx.x % CV, pl = threadIdx.x / CV, ppb = blockDim.x / CV;
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
  if (pl < ppb) {
    float sc[8], sf[8], mu[8], is[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sc[e] = __ldg(scale + v * 8 + e); sf[e] = __ldg(shift + v * 8 + e);
      mu[e] = __ldg(mean + v * 8 + e); is[e] = rsqrtf(__ldg(var + v * 8 + e) + eps);
    }
    for (int64_t pix = blockIdx.x * (int64_t)ppb + pl; pix < pixels; pix += (int64_t)gridDim.x * ppb) {
      const uint4 xu = __ldg(x + pix * x_sv + v), gu = __ldg(g + pix * g_sv + v);
      uint4 ru = make_uint4(0u, 0u, 0u, 0u);
      if (rb) ru = __ldg(rb + pix * rb_sv + v);
      const __nv_bfloat16* xe = reinterpret_cast<const __nv_bfloat16*>(&xu);
      const __nv_bfloat16* ge = reinterpret_cast<const __nv_bfloat16*>(&gu);
      const __nv_bfloat16* re = reinterpret_cast<const __nv_bfloat16*>(&ru);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xv = __bfloat162float(xe[e]);
        const float u = fmaf(xv, sc[e], sf[e]) + __bfloat162float(re[e]);
        float dz = __bfloat162float(ge[e]);
        if (slope >= 0.f && !(u > 0.f)) dz *= slope;
        s[e] += dz;
        q[e] += dz * (xv - mu[e]) * is[e];
      }
    }
  }
  // fixed-order fold over the pixel lanes of the CTA (no shared-memory atomics: the same input gives the same statistics,
  // so a train-mode forward is reproducible run to run), then one float64 atomic per channel and CTA
  if (pl < ppb) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sh[pl * CV * 16 + (v * 8 + e) * 2] = s[e];
      sh[pl * CV * 16 + (v * 8 + e) * 2 + 1] = q[e];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < CV * 16; i += blockDim.x) {
    float t = 0.f;
    for (int k = 0; k < ppb; ++k) t += sh[k * CV * 16 + i];
    atomicAdd(sums + i, (double)t);
  }
  bn_last_block_finalize(sums, fin);
}

// dx = gamma' * invstd * (dz - mean(dz) - xhat * mean(dz * xhat)); optionally dz itself (the gradient of r_before)
__global__ void __launch_bounds__(256) bn_bwd_apply_nhwc_kernel(const uint4* __restrict__ x, int CV, int x_sv,
                                                                const uint4* __restrict__ g, int g_sv,
                                                                const uint4* __restrict__ rb, int rb_sv,
                                                                const float* __restrict__ scale, const float* __restrict__ shift,
                                                                const float* __restrict__ mean, const float* __restrict__ var,
                                                                float eps, float slope, const float* __restrict__ red,
                                                                uint4* __restrict__ dx, int dx_sv,
                                                                uint4* __restrict__ dres, int dres_sv, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    const int64_t pix = i / CV;
    const uint4 xu = __ldg(x + pix * x_sv + cv), gu = __ldg(g + pix * g_sv + cv);
    uint4 ru = make_uint4(0u, 0u, 0u, 0u);
    if (rb) ru = __ldg(rb + pix * rb_sv + cv);
    const __nv_bfloat16* xe = reinterpret_cast<const __nv_bfloat16*>(&xu);
    const __nv_bfloat16* ge = reinterpret_cast<const __nv_bfloat16*>(&gu);
    const __nv_bfloat16* re = reinterpret_cast<const __nv_bfloat16*>(&ru);
    __nv_bfloat16 od[8], oz[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = cv * 8 + e;
      const float sc = __ldg(scale + c), is = rsqrtf(__ldg(var + c) + eps);
      const float xv = __bfloat162float(xe[e]);
      const float u = fmaf(xv, sc, __ldg(shift + c)) + __bfloat162float(re[e]);
      float dz = __bfloat162float(ge[e]);
      if (slope >= 0.f && !(u > 0.f)) dz *= slope;
      const float xhat = (xv - __ldg(mean + c)) * is;
      const float edz = __ldg(red + 2 * c), eydz = __ldg(red + 2 * c + 1);
      od[e] = __float2bfloat16(sc * (dz - edz - xhat * eydz));      // scale = gamma' * invstd
      oz[e] = __float2bfloat16(dz);
    }
    dx[pix * dx_sv + cv] = *reinterpret_cast<uint4*>(od);
    if (dres) dres[pix * dres_sv + cv] = *reinterpret_cast<uint4*>(oz);
  }
}

static int abn_check(const void* x, int64_t n, int64_t c, int64_t hw, int act) {
  if (!x) return fail(SNB_E_INVALID, "null tensor");
  if (n <= 0 || c <= 0 || hw <= 0 || c > INT32_MAX || n > INT32_MAX || n * c > INT32_MAX)
    return fail(SNB_E_INVALID, "bad shape n=%lld c=%lld hw=%lld", (long long)n, (long long)c, (long long)hw);
  if (act < 0 || act > 2) return fail(SNB_E_INVALID, "activation must be 0 (none), 1 (leaky_relu) or 2 (elu)");
  return SNB_OK;
}

static dim3 plane_grid(int64_t planes, int64_t hw) {
  // enough CTAs for a few waves; a plane is split only when there are few planes
  int64_t per_plane = (hw / 4 + kAbnThreads * 8 - 1) / (kAbnThreads * 8);
  const int64_t cap = std::max<int64_t>(1, (int64_t)sm_count() * 32 / planes);
  per_plane = std::max<int64_t>(1, std::min(per_plane, cap));
  return dim3((unsigned)planes, (unsigned)per_plane);
}

static dim3 stats_grid(int64_t n, int64_t c) {
  const int64_t slabs = std::max<int64_t>(1, std::min<int64_t>(n, (int64_t)sm_count() * 8 / c));
  return dim3((unsigned)c, (unsigned)slabs);
}

}  // namespace snb

using namespace snb;

extern "C" int snb_abn_forward(float* d_x, int64_t n, int64_t c, int64_t hw, const float* d_weight, const float* d_bias,
                               float* d_running_mean, float* d_running_var, int training, float momentum, float eps,
                               int activation, float slope, float* d_mean, float* d_var, double* d_workspace,
                               void* stream) {
  if (int rc = abn_check(d_x, n, c, hw, activation)) return rc;
  if ((d_weight == nullptr) != (d_bias == nullptr)) return fail(SNB_E_INVALID, "weight and bias come together (affine)");
  if (!d_running_mean || !d_running_var) return fail(SNB_E_INVALID, "running statistics are required");
  cudaStream_t st = as_stream(stream);
  const float* mean = d_running_mean;
  const float* var = d_running_var;
  if (training) {
    if (!d_mean || !d_var || !d_workspace) return fail(SNB_E_INVALID, "training mode needs mean / var outputs and a workspace");
    if (n * hw < 2) return fail(SNB_E_INVALID, "training mode needs more than one value per channel");
    SNB_CUDA_CHECK(cudaMemsetAsync(d_workspace, 0, sizeof(double) * 2 * c, st));
    abn_stats_kernel<<<stats_grid(n, c), kAbnThreads, 0, st>>>(d_x, (int)n, (int)c, hw, d_workspace);
    abn_finalize_stats_kernel<<<(unsigned)((c + 127) / 128), 128, 0, st>>>(d_workspace, d_x, hw, (int)c, (double)(n * hw),
                                                                         momentum, d_mean, d_var, d_running_mean,
                                                                         d_running_var);
    mean = d_mean;
    var = d_var;
  }
  abn_forward_kernel<<<plane_grid(n * c, hw), kAbnThreads, 0, st>>>(d_x, (int)c, hw, mean, var, d_weight, d_bias, eps,
                                                                   activation, slope);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_abn_backward(const float* d_z, const float* d_dz, int64_t n, int64_t c, int64_t hw, const float* d_var,
                                const float* d_weight, const float* d_bias, int training, float eps, int activation,
                                float slope, float* d_dx, float* d_dweight, float* d_dbias, double* d_workspace,
                                void* stream) {
  if (int rc = abn_check(d_z, n, c, hw, activation)) return rc;
  if (!d_dz || !d_dx || !d_var || !d_workspace) return fail(SNB_E_INVALID, "snb_abn_backward: null argument");
  if ((d_weight == nullptr) != (d_bias == nullptr)) return fail(SNB_E_INVALID, "weight and bias come together (affine)");
  if ((d_dweight || d_dbias) && !d_weight) return fail(SNB_E_INVALID, "parameter gradients need the parameters");
  cudaStream_t st = as_stream(stream);
  SNB_CUDA_CHECK(cudaMemsetAsync(d_workspace, 0, sizeof(double) * 2 * c, st));
  // eval mode: the reference passes edz = eydz = 0 to the backend (functions.py:110-112, its "TODO"), so dx has no
  // mean terms and dweight = dbias = 0; reproduced as is
  if (training)
    abn_bwd_reduce_kernel<<<stats_grid(n, c), kAbnThreads, 0, st>>>(d_z, d_dz, (int)n, (int)c, hw, d_weight, d_bias, eps,
                                                                   activation, slope, d_workspace);
  abn_backward_kernel<<<plane_grid(n * c, hw), kAbnThreads, 0, st>>>(d_z, d_dz, (int)c, hw, d_var, d_weight, d_bias,
                                                                    d_workspace, (double)(n * hw), training, eps,
                                                                    activation, slope, d_dx);
  if (d_dweight || d_dbias)
    abn_param_grads_kernel<<<(unsigned)((c + 127) / 128), 128, 0, st>>>(d_workspace, d_weight, (int)c, d_dweight, d_dbias);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_bn_train_nhwc(const void* d_in, int64_t pixels, int64_t channels, int64_t in_cstride, const float* d_gamma,
                                 const float* d_beta, int abn, float eps, float momentum, float* d_running_mean,
                                 float* d_running_var, float act_slope, const void* d_residual, int64_t res_cstride,
                                 int res_after_act, void* d_out, int64_t out_cstride, float* d_scale, float* d_shift,
                                 float* d_mean, float* d_var, double* d_workspace, void* stream) {
  if (!d_in || !d_out || !d_scale || !d_shift || !d_workspace) return fail(SNB_E_INVALID, "snb_bn_train_nhwc: null argument");
  if (pixels < 2 || channels <= 0 || channels % 8 || channels / 8 > kBnMaxCV)
    return fail(SNB_E_INVALID, "bad shape: pixels=%lld channels=%lld", (long long)pixels, (long long)channels);
  if (in_cstride % 8 || out_cstride % 8 || in_cstride < channels || out_cstride < channels ||
      (d_residual && (res_cstride % 8 || res_cstride < channels)))
    return fail(SNB_E_INVALID, "channel strides must be multiples of 8 covering the channels");
  if ((reinterpret_cast<uintptr_t>(d_in) & 15) || (reinterpret_cast<uintptr_t>(d_out) & 15) ||
      (reinterpret_cast<uintptr_t>(d_residual) & 15) || (reinterpret_cast<uintptr_t>(d_scale) & 15) ||
      (reinterpret_cast<uintptr_t>(d_shift) & 15))
    return fail(SNB_E_INVALID, "pointers must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  const int cv = (int)(channels / 8);
  // threads per CTA: a multiple of the vectors per pixel
  const int threads = cv >= 256 ? cv : (256 / cv) * cv;
  const int ppb = threads / cv;
  const int64_t want = (pixels + ppb * 4 - 1) / (ppb * 4);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sm_count() * 8));
  // statistics + (last block) scale / shift / running statistics: one launch, the workspace returns to zero
  BnFinalize fin{};
  fin.mode = 1; fin.C = (int)channels; fin.abn = abn; fin.count = (double)pixels; fin.eps = eps; fin.momentum = momentum;
  fin.pivot = static_cast<const __nv_bfloat16*>(d_in); fin.gamma = d_gamma; fin.beta = d_beta;
  fin.running_mean = d_running_mean; fin.running_var = d_running_var; fin.scale = d_scale; fin.shift = d_shift;
  fin.mean_out = d_mean; fin.var_out = d_var;
  bn_stats_nhwc_kernel<true><<<grid, threads, 0, st>>>(static_cast<const uint4*>(d_in), cv, (int)(in_cstride / 8), pixels, d_workspace, fin);
  const int64_t total = pixels * cv;
  const int agrid = (int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16));
  bn_apply_nhwc_kernel<<<agrid, 256, 0, st>>>(static_cast<const uint4*>(d_in), cv, (int)(in_cstride / 8), d_scale, d_shift,
                                              act_slope, static_cast<const uint4*>(d_residual), (int)(res_cstride / 8),
                                              res_after_act, static_cast<uint4*>(d_out), (int)(out_cstride / 8), total);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

static int bn_launch_shape(int64_t pixels, int64_t channels, int* threads, int* grid) {
  const int cv = (int)(channels / 8);
  *threads = cv >= 256 ? cv : (256 / cv) * cv;
  const int ppb = *threads / cv;
  const int64_t want = (pixels + ppb * 4 - 1) / (ppb * 4);
  *grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sm_count() * 8));
  return cv;
}

extern "C" int snb_bn_backward_nhwc(const void* d_x, int64_t x_cstride, const void* d_dout, int64_t dout_cstride, int64_t pixels,
                                    int64_t channels, const float* d_scale, const float* d_shift, const float* d_mean,
                                    const float* d_var, const float* d_gamma, int abn, float eps, float act_slope,
                                    const void* d_res_before, int64_t res_cstride, void* d_dx, int64_t dx_cstride,
                                    void* d_dres, int64_t dres_cstride, float* d_dgamma, float* d_dbeta, double* d_workspace,
                                    void* stream) {
  if (!d_x || !d_dout || !d_scale || !d_shift || !d_mean || !d_var || !d_dx || !d_workspace)
    return fail(SNB_E_INVALID, "snb_bn_backward_nhwc: null argument");
  if (pixels < 2 || channels <= 0 || channels % 8 || channels / 8 > kBnMaxCV) return fail(SNB_E_INVALID, "bad shape");
  if (x_cstride % 8 || dout_cstride % 8 || dx_cstride % 8 || x_cstride < channels || dout_cstride < channels ||
      dx_cstride < channels || (d_res_before && (res_cstride % 8 || res_cstride < channels)) ||
      (d_dres && (dres_cstride % 8 || dres_cstride < channels)))
    return fail(SNB_E_INVALID, "channel strides must be multiples of 8 covering the channels");
  if ((reinterpret_cast<uintptr_t>(d_x) & 15) || (reinterpret_cast<uintptr_t>(d_dout) & 15) ||
      (reinterpret_cast<uintptr_t>(d_dx) & 15) || (reinterpret_cast<uintptr_t>(d_res_before) & 15) ||
      (reinterpret_cast<uintptr_t>(d_dres) & 15))
    return fail(SNB_E_INVALID, "pointers must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  int threads, grid;
  const int cv = bn_launch_shape(pixels, channels, &threads, &grid);
  // reduction + (last block) means for the apply kernel, dgamma / dbeta: one launch, the workspace sums return to zero
  BnFinalize fin{};
  fin.mode = 3; fin.C = (int)channels; fin.abn = abn; fin.count = (double)pixels; fin.gamma = d_gamma;
  fin.out0 = d_dgamma; fin.out1 = d_dbeta;
  bn_bwd_reduce_nhwc_kernel<<<grid, threads, 0, st>>>(static_cast<const uint4*>(d_x), cv, (int)(x_cstride / 8),
                                                     static_cast<const uint4*>(d_dout), (int)(dout_cstride / 8),
                                                     static_cast<const uint4*>(d_res_before), (int)(res_cstride / 8), d_scale,
                                                     d_shift, d_mean, d_var, eps, act_slope, pixels, d_workspace, fin);
  const int64_t total = pixels * cv;
  const int agrid = (int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16));
  bn_bwd_apply_nhwc_kernel<<<agrid, 256, 0, st>>>(static_cast<const uint4*>(d_x), cv, (int)(x_cstride / 8),
                                                 static_cast<const uint4*>(d_dout), (int)(dout_cstride / 8),
                                                 static_cast<const uint4*>(d_res_before), (int)(res_cstride / 8), d_scale, d_shift,
                                                 d_mean, d_var, eps, act_slope,
                                                 reinterpret_cast<const float*>(d_workspace + 2 * channels + 1),
                                                 static_cast<uint4*>(d_dx), (int)(dx_cstride / 8), static_cast<uint4*>(d_dres),
                                                 (int)(dres_cstride / 8), total);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_channel_sum_nhwc(const void* d_in, int64_t pixels, int64_t channels, int64_t in_cstride, float* d_out,
                                    double* d_workspace, void* stream) {
  if (!d_in || !d_out || !d_workspace) return fail(SNB_E_INVALID, "snb_channel_sum_nhwc: null argument");
  if (pixels < 1 || channels <= 0 || channels % 8 || channels / 8 > kBnMaxCV || in_cstride % 8 || in_cstride < channels ||
      (reinterpret_cast<uintptr_t>(d_in) & 15))
    return fail(SNB_E_INVALID, "bad shape or alignment");
  cudaStream_t st = as_stream(stream);
  int threads, grid;
  const int cv = bn_launch_shape(pixels, channels, &threads, &grid);
  BnFinalize fin{};
  fin.mode = 2; fin.C = (int)channels; fin.out0 = d_out;
  bn_stats_nhwc_kernel<false><<<grid, threads, 0, st>>>(static_cast<const uint4*>(d_in), cv, (int)(in_cstride / 8), pixels, d_workspace, fin);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
