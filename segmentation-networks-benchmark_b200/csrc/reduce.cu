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

template <int DT>
__device__ __forceinline__ void load_target4(const void* t, int64_t i4, float (&tv)[4]) {
  if (DT == SNB_DT_I64) {
    const longlong2* p = static_cast<const longlong2*>(t) + i4 * 2;
    const longlong2 a = __ldg(p), b = __ldg(p + 1);
    tv[0] = (float)a.x; tv[1] = (float)a.y; tv[2] = (float)b.x; tv[3] = (float)b.y;
  } else if (DT == SNB_DT_U8) {
    const uchar4 a = __ldg(static_cast<const uchar4*>(t) + i4);
    tv[0] = a.x; tv[1] = a.y; tv[2] = a.z; tv[3] = a.w;
  } else {
    const float4 a = __ldg(static_cast<const float4*>(t) + i4);
    tv[0] = a.x; tv[1] = a.y; tv[2] = a.z; tv[3] = a.w;
  }
}

struct LossAcc {
  float bce, pt, p, t, focal;
  uint32_t n_pred, n_truth, n_tp;   // tp = n_tp, fp = n_pred - n_tp, fn = n_truth - n_tp, tn = n - n_pred - n_truth + n_tp
};

// The integer decision must be exactly torch's `sigmoid(x) > 0.5` in float32.  For |x| > 1e-6 that is `x > 0`
// (1 + exp(-x) rounds strictly below / above 2); only inside that sliver the rounded quotient decides, and there the
// IEEE expression is evaluated (never taken on real data, warp-uniform skip otherwise).
__device__ __forceinline__ bool sigmoid_gt_half(float x) {
  if (fabsf(x) > 1e-6f) return x > 0.f;
  return 1.f / (1.f + expf(-x)) > 0.5f;
}

// One exponential serves everything: with e = exp(-|x|), d = 1 + e
//   p = sigmoid(x)     = (x >= 0 ? 1 : e) / d
//   z = logsigmoid(x)  = min(x, 0) - log(d)
//   BCE-with-logits(z, t) = (1 - t) z - logsigmoid(z) = -t z + log(1 + exp(z)) = -t z + log(1 + p)     (z <= 0, exp(z) = p)
// which is the reference's double squash (lib/losses.py:51-53).  With 1 + p = num / d, num = x >= 0 ? 2 + e : 1 + 2e:
//   t == 0: bce = log(1 + p)                 t == 1: bce = log(num) - min(x, 0)
// so a binary target costs ONE logarithm (3 SFU operations per element: ex2, rcp, lg2 -- with 5 bytes per element for
// uint8 targets the SFU pipe, 16 results per clock per SM, is the next bound after HBM); any other target value takes
// the general two-logarithm form.  The float sums use the SFU approximations (~2^-22 relative): they are
// tolerance-checked (rel 5e-6; the reference itself sums in float32).
__device__ __forceinline__ float bce_element(float x, float t, float& p_out) {
  const float e = __expf(-fabsf(x));
  const float d = 1.f + e;
  const bool pos = x >= 0.f;
  const float p = __fdividef(pos ? 1.f : e, d);
  p_out = p;
  const float xm = fminf(x, 0.f);
  // branch-free for binary targets (a select, not a divergent branch: neighbouring lanes hold different targets)
  const float arg = t == 1.f ? (pos ? 2.f + e : fmaf(2.f, e, 1.f)) : 1.f + p;
  float b = __logf(arg) - t * xm;
  if (t != 0.f && t != 1.f) b = __logf(1.f + p) - t * (xm - __logf(d));     // soft / non-binary targets: rare, whole warps skip it
  return b;
}

// FocalLossBinary element (lib/losses.py:90-96): logpt = -bce, pt = exp(logpt), loss = (1 - pt)^gamma * bce
__device__ __forceinline__ float focal_element(float b, float gamma) {
  const float om = 1.f - __expf(-b);
  const float w = gamma == 2.f ? om * om : (gamma == 0.f ? 1.f : __powf(fmaxf(om, 0.f), gamma));
  return w * b;
}

template <bool FOCAL>
__device__ __forceinline__ float loss_accumulate(LossAcc& a, float x, float t, float gamma) {
  float p;
  const float b = bce_element(x, t, p);
  a.bce += b;
  a.pt += p * t;
  a.p += p;
  a.t += t;
  if (FOCAL) a.focal += focal_element(b, gamma);
  const bool pred = sigmoid_gt_half(x);     // metrics.py:31
  const bool truth = ((int)t & 0xff) != 0;  // target.byte()
  a.n_pred += pred;
  a.n_truth += truth;
  a.n_tp += pred && truth;
  return b;
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
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(logits) + i);
    float tv[4];
    load_target4<DT>(targets, i, tv);
    float4 b;
    b.x = loss_accumulate<FOCAL>(a, x.x, tv[0], gamma);
    b.y = loss_accumulate<FOCAL>(a, x.y, tv[1], gamma);
    b.z = loss_accumulate<FOCAL>(a, x.z, tv[2], gamma);
    b.w = loss_accumulate<FOCAL>(a, x.w, tv[3], gamma);
    if (ELEM) reinterpret_cast<float4*>(elem)[i] = b;
  }
  // ragged tail (< 4 elements) handled by the first threads of block 0
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    const float b = loss_accumulate<FOCAL>(a, logits[i], load_target<DT>(targets, i), gamma);
    if (ELEM) elem[i] = b;
  }

  __shared__ double s_f[8][5];
  __shared__ uint32_t s_i[8][3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float f0 = warp_sum(a.bce), f1 = warp_sum(a.pt), f2 = warp_sum(a.p), f3 = warp_sum(a.t);
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

template <int DT>
__global__ void __launch_bounds__(256) confusion_kernel(const float* __restrict__ probs, const void* __restrict__ targets,
                                                        int64_t n, float thr, ReduceWs* __restrict__ ws,
                                                        long long* __restrict__ counts) {
  uint32_t c[4] = {0, 0, 0, 0};
  auto add = [&](float p, float t) {
    const bool pred = p > thr;
    const bool truth = ((int)t & 0xff) != 0;
    c[0] += pred && truth; c[1] += pred && !truth; c[2] += !pred && truth; c[3] += !pred && !truth;
  };
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // two 16-byte probability loads in flight per thread and iteration (the counting itself is a handful of integer ops)
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (; i + stride < n4; i += 2 * stride) {
    const float4 x0 = __ldg(reinterpret_cast<const float4*>(probs) + i);
    const float4 x1 = __ldg(reinterpret_cast<const float4*>(probs) + i + stride);
    float t0[4], t1[4];
    load_target4<DT>(targets, i, t0);
    load_target4<DT>(targets, i + stride, t1);
    add(x0.x, t0[0]); add(x0.y, t0[1]); add(x0.z, t0[2]); add(x0.w, t0[3]);
    add(x1.x, t1[0]); add(x1.y, t1[1]); add(x1.z, t1[2]); add(x1.w, t1[3]);
  }
  if (i < n4) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(probs) + i);
    float tv[4];
    load_target4<DT>(targets, i, tv);
    add(x.x, tv[0]); add(x.y, tv[1]); add(x.z, tv[2]); add(x.w, tv[3]);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t j = (n4 << 2) + threadIdx.x;
    add(probs[j], load_target<DT>(targets, j));
  }
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
    for (int k = 0; k < 4; ++k) counts[k] = (long long)tot[k];
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
  float* s_thr = reinterpret_cast<float*>(sh + 8 * 2 * bins);
  for (int i = threadIdx.x; i < 8 * 2 * bins; i += blockDim.x) hist[i] = 0;
  for (int i = threadIdx.x; i < n_thr; i += blockDim.x) s_thr[i] = thr[i];
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
        atomicAdd(&my[truth * bins + bin_of(sigmoid_f32(xs[e]))], 1u);
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
  const size_t smem = (size_t)(8 * 2 * bins) * sizeof(uint32_t) + (size_t)n_thr * sizeof(float);
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
