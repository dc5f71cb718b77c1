// Single-pass loss / metric reductions (lib/losses.py:18-101, lib/metrics.py:9-40, lib/train_utils.py:92-125).
// HBM-bound: 4 B logit + 8/1/4 B target per element, read once with 16-byte loads; partial sums stay in
// registers and are folded with warp shuffles.  Every reduction is ONE launch: a block stores its partial in a slot of the
// caller's workspace, takes a ticket, and the block that draws the last ticket adds the slots in slot order (so the result
// is deterministic for a given grid), writes the final values and puts the workspace back to its "zero at rest" state.
// No memset nodes, no finalize launch, no process-global scratch (two streams use two workspaces).
#include <cstdint>

#include "snb_internal.h"

namespace snb {

// ---- workspace layout (snb_reduce_workspace_bytes; allocated zeroed once by the caller, one per stream in flight)
constexpr int kWsSlots = 2048;                     // >= the largest grid any reduction launches
constexpr int kMaxThr = 1024;
struct alignas(64) WsSlot { double f[5]; unsigned int c[4]; unsigned int pad[2]; };
struct ReduceWs {
  unsigned int ticket;                             // 0 at rest
  unsigned int pad[15];
  WsSlot slot[kWsSlots];
  unsigned long long pr_hist[2][kMaxThr + 1];      // PR-curve histogram, all zero at rest
};
static_assert(sizeof(WsSlot) == 64, "slot size");

__device__ __forceinline__ float sigmoid_f32(float x) { return 1.f / (1.f + expf(-x)); }

template <int DT>
__device__ __forceinline__ float load_target(const void* t, int64_t i) {
  if (DT == SNB_DT_I64) return (float)static_cast<const long long*>(t)[i];
  if (DT == SNB_DT_U8) return (float)static_cast<const uint8_t*>(t)[i];
  return static_cast<const float*>(t)[i];
}

// byte / small integer -> float on the FMA pipe: 0x4B000000 | v is the float 2^23 + v (an I2F would go through the
// 16-per-clock conversion unit, the pipe the exp / log / reciprocal of the loss already keep busy)
__device__ __forceinline__ float small_uint_to_float(uint32_t v) { return __uint_as_float(0x4B000000u | v) - 8388608.f; }

// 4 targets of element group i4 as floats.  `big` is set when an int64 target is outside [0, 2^23) (then the exact
// conversion is redone on the slow path)
template <int DT>
__device__ __forceinline__ void load_target4(const void* t, int64_t i4, float (&tv)[4], bool& big) {
  if (DT == SNB_DT_I64) {
    const longlong2* p = static_cast<const longlong2*>(t) + i4 * 2;
    const longlong2 a = __ldg(p), b = __ldg(p + 1);
    const long long v[4] = {a.x, a.y, b.x, b.y};
    const unsigned long long any = (unsigned long long)(a.x | a.y | b.x | b.y);
#pragma unroll
    for (int k = 0; k < 4; ++k) tv[k] = small_uint_to_float((uint32_t)v[k] & 0x7FFFFFu);
    big = (any >> 23) != 0;
    if (big) {
#pragma unroll
      for (int k = 0; k < 4; ++k) tv[k] = (float)v[k];
    }
  } else if (DT == SNB_DT_U8) {
    const uint32_t w = __ldg(static_cast<const uint32_t*>(t) + i4);
    tv[0] = small_uint_to_float(w & 255u); tv[1] = small_uint_to_float((w >> 8) & 255u);
    tv[2] = small_uint_to_float((w >> 16) & 255u); tv[3] = small_uint_to_float(w >> 24);
  } else {
    const float4 a = __ldg(static_cast<const float4*>(t) + i4);
    tv[0] = a.x; tv[1] = a.y; tv[2] = a.z; tv[3] = a.w;
  }
}

template <int DT>
__device__ __forceinline__ void load_target4(const void* t, int64_t i4, float (&tv)[4]) {
  bool big;
  load_target4<DT>(t, i4, tv, big);
}

// bce is accumulated as ln2 * sum log2(arg) - sum t * min(x, 0) (one multiply per block instead of one per element)
struct LossAcc {
  float lg, lin, pt, p, t, focal;
  uint32_t n_pred, n_truth, n_tp;   // tp = n_tp, fp = n_pred - n_tp, fn = n_truth - n_tp, tn = n - n_pred - n_truth + n_tp
};

constexpr float kLn2 = 0.6931471805599453f, kLog2e = 1.4426950408889634f;

// The integer decision must be exactly torch's `sigmoid(x) > 0.5` in float32.  For |x| > 1e-6 that is `x > 0`
// (1 + exp(-x) rounds strictly below / above 2); only inside that sliver the rounded quotient decides, and there the
// IEEE expression is evaluated.
// One exponential serves everything: with e = exp(-|x|), d = 1 + e
//   p = sigmoid(x)     = (x >= 0 ? 1 : e) / d
//   z = logsigmoid(x)  = min(x, 0) - log(d)
//   BCE-with-logits(z, t) = (1 - t) z - logsigmoid(z) = -t z + log(1 + exp(z)) = -t z + log(1 + p)     (z <= 0, exp(z) = p)
// which is the reference's double squash (lib/losses.py:51-53).  With 1 + p = num / d, num = x >= 0 ? 2 + e : 1 + 2e:
//   t == 0: bce = log(1 + p)                 t == 1: bce = log(num) - min(x, 0)
// so a binary target costs ONE logarithm (3 SFU operations per element: ex2, rcp, lg2).  Everything unusual -- a target that
// is not 0 or 1, a logit inside the |x| <= 1e-6 sliver -- is handled by ONE out-of-line fix-up per group of four elements
// (never taken on real data), which keeps the main loop free of divergent regions.  The float sums use the SFU
// approximations (~2^-22 relative): they are tolerance-checked (rel 5e-6; the reference itself sums in float32).
// raw SFU instructions (the __expf / __log2f intrinsics wrap them in range checks and denormal rescaling: 5 and 4
// instructions per call; here every argument is in range by construction: -|x| * log2(e) <= 0, log arguments in [1, 3])
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// the unusual element: general target value and / or a logit inside the |x| <= 1e-6 sliver (exact decision)
__device__ __forceinline__ void loss_slow_element(float x, float t, float& lg, float& lin, bool& pred, bool& truth) {
  truth = ((int)t & 0xff) != 0;                // target.byte() (metrics.py:33)
  const float e = __expf(-fabsf(x));
  const float d = 1.f + e;
  const float p = __fdividef(x >= 0.f ? 1.f : e, d);
  const float b = __logf(1.f + p) - t * (fminf(x, 0.f) - __logf(d));
  lg = b * kLog2e;
  lin = 0.f;
  pred = fabsf(x) > 1e-6f ? x > 0.f : (1.f / (1.f + expf(-x)) > 0.5f);
}

// FocalLossBinary element (lib/losses.py:90-96): logpt = -bce, pt = exp(logpt), loss = (1 - pt)^gamma * bce
__device__ __forceinline__ float focal_element(float b, float gamma) {
  const float om = 1.f - __expf(-b);
  const float w = gamma == 2.f ? om * om : (gamma == 0.f ? 1.f : __powf(fmaxf(om, 0.f), gamma));
  return w * b;
}

// four elements; bout[k] = the per-element BCE when NEEDB
template <bool FOCAL, bool NEEDB>
__device__ __forceinline__ void loss_accumulate4(LossAcc& a, const float (&x)[4], const float (&t)[4], float gamma,
                                                 float (&bout)[4]) {
  float lg[4], lin[4], p[4];
  bool pred[4], truth[4];
  bool rare = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float e = ex2_approx(fabsf(x[k]) * -kLog2e);
    const float r = rcp_approx(1.f + e);
    const bool pos = x[k] >= 0.f;
    p[k] = (pos ? 1.f : e) * r;
    const float num = pos ? 2.f + e : fmaf(2.f, e, 1.f);
    const bool one = t[k] == 1.f;
    lg[k] = lg2_approx(one ? num : 1.f + p[k]);
    lin[k] = t[k] * fminf(x[k], 0.f);
    pred[k] = x[k] > 0.f;                       // metrics.py:31 outside the sliver
    truth[k] = t[k] != 0.f;                     // target.byte() != 0 for a target that is 0 or 1
    rare |= (fabsf(x[k]) <= 1e-6f) | (!one & (t[k] != 0.f));
  }
  if (rare) {                                   // one cold region per four elements
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if ((fabsf(x[k]) <= 1e-6f) | ((t[k] != 1.f) & (t[k] != 0.f))) loss_slow_element(x[k], t[k], lg[k], lin[k], pred[k], truth[k]);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    a.lg += lg[k];
    a.lin += lin[k];
    a.pt = fmaf(p[k], t[k], a.pt);
    a.p += p[k];
    a.t += t[k];
    if (FOCAL || NEEDB) {
      const float b = fmaf(lg[k], kLn2, -lin[k]);
      bout[k] = b;
      if (FOCAL) a.focal += focal_element(b, gamma);
    }
    a.n_pred += pred[k];
    a.n_truth += truth[k];
    a.n_tp += pred[k] && truth[k];
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block partial -> workspace slot -> ticket.  Returns true in every thread of the block that drew the last ticket, after
// which the slots of all blocks are visible to it (release by __threadfence + atomic, acquire by the fence after it).
__device__ __forceinline__ bool publish_and_elect(ReduceWs* ws) {
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(&ws->ticket, 1u);
    s_last = ticket == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

// Deterministic sum of the first `n_slots` slots by a 256-thread block: thread j adds slots j, j + 256, ... in order, then a
// fixed-shape shuffle butterfly per warp and the 8 warp totals in warp order.  NF float sums + NC counts.
template <int NF, int NC>
__device__ __forceinline__ void sum_slots(const ReduceWs* ws, int n_slots, double (&f)[NF], unsigned long long (&c)[NC]) {
  __shared__ double s_f[8][NF];
  __shared__ unsigned long long s_c[8][NC];
  double lf[NF];
  unsigned long long lc[NC];
#pragma unroll
  for (int k = 0; k < NF; ++k) lf[k] = 0.0;
#pragma unroll
  for (int k = 0; k < NC; ++k) lc[k] = 0;
  for (int j = threadIdx.x; j < n_slots; j += 256) {
    const WsSlot* sl = &ws->slot[j];
#pragma unroll
    for (int k = 0; k < NF; ++k) lf[k] += __ldcg(&sl->f[k]);
#pragma unroll
    for (int k = 0; k < NC; ++k) lc[k] += __ldcg(&sl->c[k]);
  }
  // fixed-shape butterfly inside each warp (the same order every run), then the 8 warp totals in warp order
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < NF; ++k) lf[k] += __shfl_xor_sync(0xffffffffu, lf[k], o);
#pragma unroll
    for (int k = 0; k < NC; ++k) lc[k] += __shfl_xor_sync(0xffffffffu, lc[k], o);
  }
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < NF; ++k) s_f[warp][k] = lf[k];
#pragma unroll
    for (int k = 0; k < NC; ++k) s_c[warp][k] = lc[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NF; ++k) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_f[w][k];
    f[k] = t;
  }
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    unsigned long long t = 0;
    for (int w = 0; w < 8; ++w) t += s_c[w][k];
    c[k] = t;
  }
}

// sums[0..4] = sum bce, sum p t, sum p, sum t, sum focal (0 unless FOCAL); counts = tp, fp, fn, tn; ELEM also writes
// the per-element BCE (BCEWithSigmoidLoss(reduce=False), lib/losses.py:46-53).
template <int DT, bool FOCAL, bool ELEM>
__global__ void __launch_bounds__(256) loss_reduce_kernel(const float* __restrict__ logits, const void* __restrict__ targets,
                                                          int64_t n, float gamma, float* __restrict__ elem,
                                                          ReduceWs* __restrict__ ws, double* __restrict__ sums,
                                                          long long* __restrict__ counts) {
  LossAcc a{};
  const uint32_t n4 = (uint32_t)(n >> 2);                    // the host checks n < 2^33
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 xv = __ldg(reinterpret_cast<const float4*>(logits) + i);
    float tv[4], b[4];
    load_target4<DT>(targets, i, tv);
    const float x[4] = {xv.x, xv.y, xv.z, xv.w};
    loss_accumulate4<FOCAL, ELEM>(a, x, tv, gamma, b);
    if (ELEM) reinterpret_cast<float4*>(elem)[i] = make_float4(b[0], b[1], b[2], b[3]);
  }
  // ragged tail (< 4 elements): the first threads of block 0, one element each, through the general (slow-path) formulas
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = ((int64_t)n4 << 2) + threadIdx.x;
    const float x = logits[i], t = load_target<DT>(targets, i);
    float lg, lin;
    bool pred, truth;
    loss_slow_element(x, t, lg, lin, pred, truth);
    const float e = __expf(-fabsf(x));
    const float p = __fdividef(x >= 0.f ? 1.f : e, 1.f + e);
    const float b = lg * kLn2;
    a.lg += lg; a.lin += lin; a.pt = fmaf(p, t, a.pt); a.p += p; a.t += t;
    if (FOCAL) a.focal += focal_element(b, gamma);
    if (ELEM) elem[i] = b;
    a.n_pred += pred; a.n_truth += truth; a.n_tp += pred && truth;
  }

  __shared__ double s_f[8][5];
  __shared__ uint32_t s_i[8][3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float f0 = fmaf(warp_sum(a.lg), kLn2, -warp_sum(a.lin)), f1 = warp_sum(a.pt), f2 = warp_sum(a.p), f3 = warp_sum(a.t);
  const float f4 = FOCAL ? warp_sum(a.focal) : 0.f;
  const uint32_t c0 = __reduce_add_sync(0xffffffffu, a.n_pred), c1 = __reduce_add_sync(0xffffffffu, a.n_truth);
  const uint32_t c2 = __reduce_add_sync(0xffffffffu, a.n_tp);
  if (lane == 0) {
    s_f[warp][0] = f0; s_f[warp][1] = f1; s_f[warp][2] = f2; s_f[warp][3] = f3; s_f[warp][4] = f4;
    s_i[warp][0] = c0; s_i[warp][1] = c1; s_i[warp][2] = c2;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double f = 0.0;
    for (int w = 0; w < 8; ++w) f += s_f[w][threadIdx.x];
    ws->slot[blockIdx.x].f[threadIdx.x] = f;
  } else if (threadIdx.x >= 32 && threadIdx.x < 35) {
    unsigned int c = 0;
    for (int w = 0; w < 8; ++w) c += s_i[w][threadIdx.x - 32];
    ws->slot[blockIdx.x].c[threadIdx.x - 32] = c;
  }
  if (!publish_and_elect(ws)) return;
  double f[5];
  unsigned long long c[3];
  sum_slots<5, 3>(ws, gridDim.x, f, c);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 5; ++k) sums[k] = f[k];
    const long long n_pred = (long long)c[0], n_truth = (long long)c[1], n_tp = (long long)c[2];
    counts[0] = n_tp;
    counts[1] = n_pred - n_tp;
    counts[2] = n_truth - n_tp;
    counts[3] = (long long)n - n_pred - n_truth + n_tp;
    ws->ticket = 0;                       // back to rest (stream order makes it visible to the next launch)
  }
}

// truth bits (target.byte() != 0) of the four targets of element group i4, straight from the stored type
template <int DT>
__device__ __forceinline__ uint32_t load_truth4(const void* t, uint32_t i4) {
  if (DT == SNB_DT_U8) {
    const uint32_t w = __ldg(static_cast<const uint32_t*>(t) + i4);
    return ((w & 0xffu) != 0) | (((w & 0xff00u) != 0) << 1) | (((w & 0xff0000u) != 0) << 2) | (((w >> 24) != 0) << 3);
  } else if (DT == SNB_DT_I64) {
    const longlong2* p = static_cast<const longlong2*>(t) + (size_t)i4 * 2;
    const longlong2 a = __ldg(p), b = __ldg(p + 1);
    return ((a.x & 0xff) != 0) | (((a.y & 0xff) != 0) << 1) | (((b.x & 0xff) != 0) << 2) | (((b.y & 0xff) != 0) << 3);
  } else {
    const float4 a = __ldg(static_cast<const float4*>(t) + i4);
    return ((((int)a.x) & 0xff) != 0) | (((((int)a.y) & 0xff) != 0) << 1) | (((((int)a.z) & 0xff) != 0) << 2) |
           (((((int)a.w) & 0xff) != 0) << 3);
  }
}

template <int DT>
__global__ void __launch_bounds__(256) confusion_kernel(const float* __restrict__ probs, const void* __restrict__ targets,
                                                        int64_t n, float thr, ReduceWs* __restrict__ ws,
                                                        long long* __restrict__ counts) {
  // n_pred, n_truth, n_tp (tp / fp / fn / tn follow from them and n)
  uint32_t n_pred = 0, n_truth = 0, n_tp = 0;
  auto add4 = [&](const float4& x, uint32_t tb) {
    const uint32_t pb = (x.x > thr) | ((x.y > thr) << 1) | ((x.z > thr) << 2) | ((x.w > thr) << 3);
    n_pred += __popc(pb);
    n_truth += __popc(tb);
    n_tp += __popc(pb & tb);
  };
  const uint32_t n4 = (uint32_t)(n >> 2);
  const uint32_t stride = gridDim.x * blockDim.x;
  // four 16-byte probability loads (and their targets) in flight per thread and trip
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 x[4];
    uint32_t tb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      x[u] = __ldg(reinterpret_cast<const float4*>(probs) + i + u * stride);
      tb[u] = load_truth4<DT>(targets, i + u * stride);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) add4(x[u], tb[u]);
  }
  for (; i < n4; i += stride) add4(__ldg(reinterpret_cast<const float4*>(probs) + i), load_truth4<DT>(targets, i));
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t j = ((int64_t)n4 << 2) + threadIdx.x;
    const bool pred = probs[j] > thr;
    const bool truth = ((int)load_target<DT>(targets, j) & 0xff) != 0;
    n_pred += pred; n_truth += truth; n_tp += pred && truth;
  }
  const uint32_t c[4] = {n_tp, n_pred - n_tp, n_truth - n_tp, 0};
  __shared__ uint32_t s_i[8][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t v = __reduce_add_sync(0xffffffffu, c[k]);
    if (lane == 0) s_i[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    unsigned int t = 0;
    for (int w = 0; w < 8; ++w) t += s_i[w][threadIdx.x];
    ws->slot[blockIdx.x].c[threadIdx.x] = t;
  }
  if (!publish_and_elect(ws)) return;
  double f[1];
  unsigned long long tot[4];
  sum_slots<0 + 1, 4>(ws, gridDim.x, f, tot);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) counts[k] = (long long)tot[k];
    counts[3] = (long long)n - (long long)(tot[0] + tot[1] + tot[2]);
    ws->ticket = 0;
  }
}

// ------------------------------------------------------------------------------------------- PR curve
// hist[truth][idx], idx = number of thresholds strictly below sigmoid(x), lives in the caller's workspace (zero at rest)
template <int DT>
__global__ void __launch_bounds__(256) pr_hist_kernel(const float* __restrict__ logits, const void* __restrict__ targets,
                                                      int64_t n, const float* __restrict__ thr, int n_thr,
                                                      ReduceWs* __restrict__ ws, unsigned long long* __restrict__ tp,
                                                      unsigned long long* __restrict__ tn, unsigned long long* __restrict__ fp,
                                                      unsigned long long* __restrict__ fn) {
  extern __shared__ uint32_t sh[];            // [8 warps][2][n_thr+1] then thresholds
  const int bins = n_thr + 1;
  uint32_t* hist = sh;
  float* s_thr = reinterpret_cast<float*>(sh + 8 * 2 * bins) + 1;   // s_thr[-1] = -inf, s_thr[n_thr] = +inf (sentinels)
  for (int i = threadIdx.x; i < 8 * 2 * bins; i += blockDim.x) hist[i] = 0;
  for (int i = threadIdx.x; i < n_thr; i += blockDim.x) s_thr[i] = thr[i];
  if (threadIdx.x == 0) {
    s_thr[-1] = -INFINITY;
    s_thr[n_thr] = INFINITY;
  }
  __syncthreads();
  uint32_t* my = hist + (threadIdx.x >> 5) * 2 * bins;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // idx = #{k : p > thr[k]} for ascending thresholds.  The reference's thresholds are (nearly) uniform (arange(0, 1, 1/127),
  // lib/train_utils.py:97), so a linear guess lands within one bin and two short fix-up loops make it exact for ANY
  // ascending table (they replace a 7-step binary search per element; the result is the same integer)
  const float t0 = s_thr[0];
  const float span = s_thr[n_thr - 1] - t0;
  const float gscale = span > 0.f ? (float)(n_thr - 1) / span : 0.f;
  auto bin_of = [&](float p) {
    int lo = min(n_thr, max(0, __float2int_rd((p - t0) * gscale) + 1));   // NaN -> 0
    while (lo < n_thr && p > s_thr[lo]) ++lo;
    while (lo > 0 && !(p > s_thr[lo - 1])) --lo;
    return lo;
  };
  // Fast path: the bin of sigmoid_f32(x) (the exact float the reference compares, expf + IEEE division, ~40 instructions)
  // is decided from an SFU approximation pa (ex2.approx + rcp.approx, 4 instructions) whenever pa is farther than kEps from
  // both neighbouring thresholds: |pa - sigmoid(x)| <= 1e-6 for |x| <= 20 (ex2.approx 2^-22 relative, its argument rounding
  // |x| 2^-24 log2(e) p (1 - p), rcp.approx 1 ulp) and |sigmoid_f32(x) - sigmoid(x)| <= 4 ulp <= 2.4e-7, so both floats lie
  // in the same open interval between two thresholds.  Everything else -- within kEps of a threshold (~5e-4 of uniform
  // probabilities), |x| > 20, NaN -- takes the exact evaluation through ONE cold branch.  Same integers, 3x fewer instructions.
  constexpr float kEps = 2e-6f;
  auto bin_fast = [&](float x, int& lo) {
    float e, pa;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(pa) : "f"(1.f + e));
    lo = min(n_thr, max(0, __float2int_rd((pa - t0) * gscale) + 1));
    return fabsf(x) <= 20.f && pa - s_thr[lo - 1] > kEps && s_thr[lo] - pa > kEps;   // (NaN fails every comparison)
  };
  const bool vec = (n & 3) == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0 && (reinterpret_cast<uintptr_t>(targets) & 15) == 0;
  if (vec) {
    for (int64_t i4 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i4 < n / 4; i4 += stride) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(logits) + i4);
      float tv[4];
      load_target4<DT>(targets, i4, tv);
      const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int truth = ((int)tv[e]) != 0;                         // astype(int32) then k*true with k=2
        int lo;
        if (!bin_fast(xs[e], lo)) lo = bin_of(sigmoid_f32(xs[e]));
        atomicAdd(&my[truth * bins + lo], 1u);
      }
    }
  } else {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
      const float p = sigmoid_f32(__ldg(logits + i));
      const int truth = ((int)load_target<DT>(targets, i)) != 0;
      atomicAdd(&my[truth * bins + bin_of(p)], 1u);
    }
  }
  __syncthreads();
  unsigned long long* g_hist = &ws->pr_hist[0][0];
  for (int i = threadIdx.x; i < 2 * bins; i += blockDim.x) {
    unsigned long long t = 0;
    for (int w = 0; w < 8; ++w) t += hist[w * 2 * bins + i];
    if (t) atomicAdd(&g_hist[(i / bins) * (kMaxThr + 1) + i % bins], t);
  }
  if (!publish_and_elect(ws)) return;
  // last block: suffix sums.  pred_k = idx > k: tp += sum_{idx>k} h[1][idx], fp += sum_{idx>k} h[0][idx], fn / tn the
  // complements; then the histogram goes back to zero.  The shared histogram area is reused: [2][bins] u64 totals.
  unsigned long long* s_h = reinterpret_cast<unsigned long long*>(sh);     // 2 * bins * 8 bytes <= 8 * 2 * bins * 4
  for (int i = threadIdx.x; i < 2 * bins; i += blockDim.x) s_h[i] = __ldcg(&g_hist[(i / bins) * (kMaxThr + 1) + i % bins]);
  __syncthreads();
  for (int k = threadIdx.x; k < n_thr; k += blockDim.x) {
    unsigned long long pos1 = 0, pos0 = 0, neg1 = 0, neg0 = 0;
    for (int idx = 0; idx <= n_thr; ++idx) {
      if (idx > k) { pos1 += s_h[bins + idx]; pos0 += s_h[idx]; }
      else { neg1 += s_h[bins + idx]; neg0 += s_h[idx]; }
    }
    tp[k] += pos1; fp[k] += pos0; fn[k] += neg1; tn[k] += neg0;
  }
  for (int i = threadIdx.x; i < 2 * bins; i += blockDim.x) g_hist[(i / bins) * (kMaxThr + 1) + i % bins] = 0;
  if (threadIdx.x == 0) ws->ticket = 0;
}

static int reduce_grid(int64_t n_items) {
  const int64_t need = (n_items + 255) / 256;
  int64_t cap = (int64_t)sm_count() * 8;
  if (cap > kWsSlots) cap = kWsSlots;
  return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

// Gradient with respect to the logits of
//   c_bce * sum_i bce_i + c_focal * sum_i focal_i + c_jac * (1 - A / D),   A = sum p t + smooth_num,
//                                                                          D = sum p + sum t - sum p t + smooth_den
// (lib/losses.py:18-101 under autograd):
//   bce_i   = BCE-with-logits(z, t), z = logsigmoid(x):  d/dx = (sigmoid(z) - t) * dz/dx = (p / (1 + p) - t) * (1 - p)
//   focal_i = (1 - e^-b)^gamma b, b = bce_i:             d/db = gamma (1 - e^-b)^(gamma-1) e^-b b + (1 - e^-b)^gamma
//   jaccard:  d/dp_i = (A (1 - t_i) - t_i D) / D^2,  dp/dx = p (1 - p)
// sums = the output of snb_loss_iou_reduce on the same tensors; grad_out = upstream gradient on the device: NULL = 1, a
// scalar, or (PER_ELEM) one value per element (BCEWithSigmoidLoss(reduce=False)), so no host synchronisation is needed.
template <int DT, bool FOCAL, bool PER_ELEM>
__global__ void __launch_bounds__(256) loss_grad_kernel(const float* __restrict__ logits, const void* __restrict__ targets,
                                                        int64_t n, const double* __restrict__ sums,
                                                        const float* __restrict__ grad_out, float c_bce, float c_focal,
                                                        float gamma, float c_jac, float smooth_num, float smooth_den,
                                                        float* __restrict__ grad) {
  float jb = 0.f, jt = 0.f;
  if (c_jac != 0.f) {
    const double A = sums[1] + smooth_num, D = sums[2] + sums[3] - sums[1] + smooth_den;
    jb = (float)(A / (D * D)) * c_jac;        // coefficient of (1 - t)
    jt = (float)(1.0 / D) * c_jac;            // coefficient of t
  }
  const float g0 = (grad_out && !PER_ELEM) ? __ldg(grad_out) : 1.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = __ldg(logits + i);
    const float t = load_target<DT>(targets, i);
    const float p = sigmoid_f32(x);
    const float q = 1.f - p;
    float cb = c_bce;
    if (FOCAL) {
      // b from the accurate functions: the gradient is tolerance-checked element by element
      const float z = fminf(x, 0.f) - log1pf(expf(-fabsf(x)));
      const float b = log1pf(p) - t * z;
      const float pt = expf(-b), om = fmaxf(1.f - pt, 0.f);
      float w, dw;          // om^gamma and gamma * om^(gamma-1)
      if (gamma == 2.f) { w = om * om; dw = 2.f * om; }
      else if (gamma == 0.f) { w = 1.f; dw = 0.f; }
      else { w = powf(om, gamma); dw = om > 0.f ? gamma * powf(om, gamma - 1.f) : 0.f; }
      cb += c_focal * (dw * pt * b + w);
    }
    const float g = PER_ELEM ? __ldg(grad_out + i) : g0;
    grad[i] = g * q * (cb * (p / (1.f + p) - t) + p * (jb * (1.f - t) - jt * t));
  }
}

}  // namespace snb

using namespace snb;

static int check_targets(const void* d_targets, int dt, int64_t n, bool vec) {
  if (dt != SNB_DT_I64 && dt != SNB_DT_U8 && dt != SNB_DT_F32) return fail(SNB_E_INVALID, "target dtype %d unsupported", dt);
  if (n < 0) return fail(SNB_E_INVALID, "negative length");
  const uintptr_t a = reinterpret_cast<uintptr_t>(d_targets);
  if (vec && ((dt == SNB_DT_I64 && (a & 15)) || (dt == SNB_DT_F32 && (a & 15)) || (dt == SNB_DT_U8 && (a & 3))))
    return fail(SNB_E_INVALID, "targets must be 16-byte aligned (4-byte for u8)");
  return SNB_OK;
}

static int check_workspace(const void* ws) {
  if (!ws) return fail(SNB_E_INVALID, "null reduction workspace (snb_reduce_workspace_bytes() zeroed bytes)");
  if (reinterpret_cast<uintptr_t>(ws) & 63) return fail(SNB_E_INVALID, "the reduction workspace must be 64-byte aligned");
  return SNB_OK;
}

extern "C" int64_t snb_reduce_workspace_bytes(void) { return (int64_t)sizeof(ReduceWs); }

extern "C" int snb_loss_iou_reduce(const float* d_logits, const void* d_targets, int target_dtype, int64_t n,
                                   float focal_gamma, float* d_elem_bce, double* d_sums, int64_t* d_counts,
                                   void* d_workspace, void* stream) {
  if (!d_logits || !d_targets || !d_sums || !d_counts) return fail(SNB_E_INVALID, "snb_loss_iou_reduce: null argument");
  if (int rc = check_workspace(d_workspace)) return rc;
  if (int rc = check_targets(d_targets, target_dtype, n, true)) return rc;
  if (reinterpret_cast<uintptr_t>(d_logits) & 15) return fail(SNB_E_INVALID, "logits must be 16-byte aligned");
  if (n >= (int64_t(1) << 33)) return fail(SNB_E_UNSUPPORTED, "more than 2^33 elements");
  if (d_elem_bce && (reinterpret_cast<uintptr_t>(d_elem_bce) & 15)) return fail(SNB_E_INVALID, "d_elem_bce must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  if (n == 0) {
    SNB_CUDA_CHECK(cudaMemsetAsync(d_sums, 0, 5 * sizeof(double), st));
    SNB_CUDA_CHECK(cudaMemsetAsync(d_counts, 0, 4 * sizeof(int64_t), st));
    return SNB_OK;
  }
  const int grid = reduce_grid((n + 3) / 4);
  ReduceWs* ws = static_cast<ReduceWs*>(d_workspace);
  long long* cnt = reinterpret_cast<long long*>(d_counts);
  const bool focal = focal_gamma >= 0.f, elem = d_elem_bce != nullptr;
#define SNB_LOSS(DT)                                                                                                       \
  do {                                                                                                                     \
    if (focal && elem) loss_reduce_kernel<DT, true, true><<<grid, 256, 0, st>>>(d_logits, d_targets, n, focal_gamma, d_elem_bce, ws, d_sums, cnt);   \
    else if (focal) loss_reduce_kernel<DT, true, false><<<grid, 256, 0, st>>>(d_logits, d_targets, n, focal_gamma, d_elem_bce, ws, d_sums, cnt);     \
    else if (elem) loss_reduce_kernel<DT, false, true><<<grid, 256, 0, st>>>(d_logits, d_targets, n, focal_gamma, d_elem_bce, ws, d_sums, cnt);      \
    else loss_reduce_kernel<DT, false, false><<<grid, 256, 0, st>>>(d_logits, d_targets, n, focal_gamma, d_elem_bce, ws, d_sums, cnt);               \
  } while (0)
  if (target_dtype == SNB_DT_I64) SNB_LOSS(SNB_DT_I64);
  else if (target_dtype == SNB_DT_U8) SNB_LOSS(SNB_DT_U8);
  else SNB_LOSS(SNB_DT_F32);
#undef SNB_LOSS
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_loss_grad(const float* d_logits, const void* d_targets, int target_dtype, int64_t n,
                             const double* d_sums, const float* d_grad_out, int grad_out_per_element, float c_bce,
                             float c_focal, float focal_gamma, float c_jac, float smooth_num, float smooth_den,
                             float* d_grad_logits, void* stream) {
  if (!d_logits || !d_targets || !d_sums || !d_grad_logits) return fail(SNB_E_INVALID, "snb_loss_grad: null argument");
  if (grad_out_per_element && !d_grad_out) return fail(SNB_E_INVALID, "per-element upstream gradient is null");
  if (int rc = check_targets(d_targets, target_dtype, n, false)) return rc;
  if (n == 0) return SNB_OK;
  cudaStream_t st = as_stream(stream);
  const int grid = reduce_grid(n);
  const bool focal = c_focal != 0.f, pe = grad_out_per_element != 0;
#define SNB_GRAD(DT)                                                                                                      \
  do {                                                                                                                    \
    if (focal && pe) loss_grad_kernel<DT, true, true><<<grid, 256, 0, st>>>(d_logits, d_targets, n, d_sums, d_grad_out, c_bce, c_focal, focal_gamma, c_jac, smooth_num, smooth_den, d_grad_logits);  \
    else if (focal) loss_grad_kernel<DT, true, false><<<grid, 256, 0, st>>>(d_logits, d_targets, n, d_sums, d_grad_out, c_bce, c_focal, focal_gamma, c_jac, smooth_num, smooth_den, d_grad_logits);   \
    else if (pe) loss_grad_kernel<DT, false, true><<<grid, 256, 0, st>>>(d_logits, d_targets, n, d_sums, d_grad_out, c_bce, c_focal, focal_gamma, c_jac, smooth_num, smooth_den, d_grad_logits);      \
    else loss_grad_kernel<DT, false, false><<<grid, 256, 0, st>>>(d_logits, d_targets, n, d_sums, d_grad_out, c_bce, c_focal, focal_gamma, c_jac, smooth_num, smooth_den, d_grad_logits);             \
  } while (0)
  if (target_dtype == SNB_DT_I64) SNB_GRAD(SNB_DT_I64);
  else if (target_dtype == SNB_DT_U8) SNB_GRAD(SNB_DT_U8);
  else SNB_GRAD(SNB_DT_F32);
#undef SNB_GRAD
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_confusion_counts(const float* d_probs, const void* d_targets, int target_dtype, int64_t n, float thr,
                                    int64_t* d_counts, void* d_workspace, void* stream) {
  if (!d_probs || !d_targets || !d_counts) return fail(SNB_E_INVALID, "snb_confusion_counts: null argument");
  if (int rc = check_workspace(d_workspace)) return rc;
  if (int rc = check_targets(d_targets, target_dtype, n, true)) return rc;
  if (reinterpret_cast<uintptr_t>(d_probs) & 15) return fail(SNB_E_INVALID, "probs must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  if (n == 0) {
    SNB_CUDA_CHECK(cudaMemsetAsync(d_counts, 0, 4 * sizeof(int64_t), st));
    return SNB_OK;
  }
  if (n >= (int64_t(1) << 33)) return fail(SNB_E_UNSUPPORTED, "more than 2^33 elements");
  const int grid = reduce_grid((n + 3) / 4);
  ReduceWs* ws = static_cast<ReduceWs*>(d_workspace);
  long long* cnt = reinterpret_cast<long long*>(d_counts);
  if (target_dtype == SNB_DT_I64) confusion_kernel<SNB_DT_I64><<<grid, 256, 0, st>>>(d_probs, d_targets, n, thr, ws, cnt);
  else if (target_dtype == SNB_DT_U8) confusion_kernel<SNB_DT_U8><<<grid, 256, 0, st>>>(d_probs, d_targets, n, thr, ws, cnt);
  else confusion_kernel<SNB_DT_F32><<<grid, 256, 0, st>>>(d_probs, d_targets, n, thr, ws, cnt);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_pr_curve_update(const float* d_logits, const void* d_targets, int target_dtype, int64_t n,
                                   const float* d_thresholds, int64_t n_thr, uint64_t* d_tp, uint64_t* d_tn,
                                   uint64_t* d_fp, uint64_t* d_fn, void* d_workspace, void* stream) {
  if (!d_logits || !d_targets || !d_thresholds || !d_tp || !d_tn || !d_fp || !d_fn)
    return fail(SNB_E_INVALID, "snb_pr_curve_update: null argument");
  if (int rc = check_workspace(d_workspace)) return rc;
  if (n_thr < 1 || n_thr > kMaxThr) return fail(SNB_E_INVALID, "n_thr=%lld not in [1, %d]", (long long)n_thr, kMaxThr);
  if (int rc = check_targets(d_targets, target_dtype, n, false)) return rc;
  if (n == 0) return SNB_OK;
  cudaStream_t st = as_stream(stream);
  const int bins = (int)n_thr + 1;
  const size_t smem = (size_t)(8 * 2 * bins) * sizeof(uint32_t) + (size_t)(n_thr + 2) * sizeof(float);   // + two sentinels
  const int grid = reduce_grid(n);
  ReduceWs* ws = static_cast<ReduceWs*>(d_workspace);
#define SNB_PR(DT)                                                                                         \
  do {                                                                                                     \
    if (smem > 48 * 1024)                                                                                  \
      SNB_CUDA_CHECK(cudaFuncSetAttribute(pr_hist_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)smem));                                                     \
    pr_hist_kernel<DT><<<grid, 256, smem, st>>>(                                                            \
        d_logits, d_targets, n, d_thresholds, (int)n_thr, ws, reinterpret_cast<unsigned long long*>(d_tp),  \
        reinterpret_cast<unsigned long long*>(d_tn), reinterpret_cast<unsigned long long*>(d_fp),           \
        reinterpret_cast<unsigned long long*>(d_fn));                                                       \
  } while (0)
  if (target_dtype == SNB_DT_I64) SNB_PR(SNB_DT_I64);
  else if (target_dtype == SNB_DT_U8) SNB_PR(SNB_DT_U8);
  else SNB_PR(SNB_DT_F32);
#undef SNB_PR
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
