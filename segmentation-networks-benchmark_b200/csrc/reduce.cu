// Single-pass loss / metric reductions (lib/losses.py:31-75, lib/metrics.py:9-40, lib/train_utils.py:92-125).
// HBM-bound: 4 B logit + 8/1/4 B target per element, read once with 16-byte loads; partial sums stay in
// registers, are folded with warp shuffles, then one atomic per block (fp64 for the float sums so the result
// does not depend on the grid shape to fp32 precision; u64 for the integer counts, which are order independent).
#include <cstdint>

#include "snb_internal.h"

namespace snb {

__device__ __forceinline__ float sigmoid_f32(float x) { return 1.f / (1.f + expf(-x)); }

// F.logsigmoid: min(x, 0) - log1p(exp(-|x|))
__device__ __forceinline__ float logsigmoid_f32(float x) { return fminf(x, 0.f) - log1pf(expf(-fabsf(x))); }

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
  float bce, pt, p, t;
  uint32_t n_pred, n_truth, n_tp;   // tp = n_tp, fp = n_pred - n_tp, fn = n_truth - n_tp, tn = n - n_pred - n_truth + n_tp
};

// The integer decision must be exactly torch's `sigmoid(x) > 0.5` in float32.  For |x| > 1e-6 that is `x > 0`
// (1 + exp(-x) rounds strictly below / above 2); only inside that sliver the rounded quotient decides, and there the
// IEEE expression is evaluated (never taken on real data, warp-uniform skip otherwise).
__device__ __forceinline__ bool sigmoid_gt_half(float x) {
  if (fabsf(x) > 1e-6f) return x > 0.f;
  return 1.f / (1.f + expf(-x)) > 0.5f;
}

// One exponential serves everything: with e = exp(-|x|)
//   p = sigmoid(x)     = (x >= 0 ? 1 : e) / (1 + e)
//   z = logsigmoid(x)  = min(x, 0) - log(1 + e)
//   BCE-with-logits(z, t) = (1 - t) z - logsigmoid(z) = -t z + log(1 + exp(z)) = -t z + log(1 + p)     (z <= 0, exp(z) = p)
// which is the reference's double squash (lib/losses.py:51-53) with 1 exp, 2 log, 1 divide instead of 3 exp, 3 log1p.
// The float sums use the SFU approximations (ex2 / lg2 / rcp, ~2^-22 relative): the sums are tolerance-checked
// (rel 5e-6, the reference itself sums in float32), and it is what makes the kernel HBM-bound instead of issue-bound.
__device__ __forceinline__ void loss_accumulate(LossAcc& a, float x, float t) {
  const float e = __expf(-fabsf(x));
  const float d = 1.f + e;
  const float p = __fdividef(x >= 0.f ? 1.f : e, d);
  const float z = fminf(x, 0.f) - __logf(d);
  a.bce += __logf(1.f + p) - t * z;
  a.pt += p * t;
  a.p += p;
  a.t += t;
  const bool pred = sigmoid_gt_half(x);     // metrics.py:31
  const bool truth = ((int)t & 0xff) != 0;  // target.byte()
  a.n_pred += pred;
  a.n_truth += truth;
  a.n_tp += pred && truth;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int DT>
__global__ void __launch_bounds__(256) loss_iou_kernel(const float* __restrict__ logits, const void* __restrict__ targets,
                                                       int64_t n, double* __restrict__ sums,
                                                       unsigned long long* __restrict__ counts) {
  LossAcc a{};
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(logits) + i);
    float tv[4];
    load_target4<DT>(targets, i, tv);
    loss_accumulate(a, x.x, tv[0]);
    loss_accumulate(a, x.y, tv[1]);
    loss_accumulate(a, x.z, tv[2]);
    loss_accumulate(a, x.w, tv[3]);
  }
  // ragged tail (< 4 elements) handled by the first threads of block 0
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    loss_accumulate(a, logits[i], load_target<DT>(targets, i));
  }

  __shared__ double s_f[8][4];
  __shared__ uint32_t s_i[8][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float f0 = warp_sum(a.bce), f1 = warp_sum(a.pt), f2 = warp_sum(a.p), f3 = warp_sum(a.t);
  const uint32_t c0 = __reduce_add_sync(0xffffffffu, a.n_pred), c1 = __reduce_add_sync(0xffffffffu, a.n_truth);
  const uint32_t c2 = __reduce_add_sync(0xffffffffu, a.n_tp), c3 = 0;
  if (lane == 0) {
    s_f[warp][0] = f0; s_f[warp][1] = f1; s_f[warp][2] = f2; s_f[warp][3] = f3;
    s_i[warp][0] = c0; s_i[warp][1] = c1; s_i[warp][2] = c2; s_i[warp][3] = c3;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double f = 0.0;
    unsigned long long c = 0;
    for (int w = 0; w < 8; ++w) { f += s_f[w][threadIdx.x]; c += s_i[w][threadIdx.x]; }
    atomicAdd(&sums[threadIdx.x], f);
    atomicAdd(&counts[threadIdx.x], c);   // raw {n_pred, n_truth, n_tp, 0}; loss_iou_finalize turns them into tp/fp/fn/tn
  }
}

__global__ void loss_iou_finalize(unsigned long long* counts, unsigned long long n) {
  const unsigned long long n_pred = counts[0], n_truth = counts[1], n_tp = counts[2];
  counts[0] = n_tp;
  counts[1] = n_pred - n_tp;
  counts[2] = n_truth - n_tp;
  counts[3] = n - n_pred - n_truth + n_tp;
}

template <int DT>
__global__ void __launch_bounds__(256) confusion_kernel(const float* __restrict__ probs, const void* __restrict__ targets,
                                                        int64_t n, float thr, unsigned long long* __restrict__ counts) {
  uint32_t c[4] = {0, 0, 0, 0};
  auto add = [&](float p, float t) {
    const bool pred = p > thr;
    const bool truth = ((int)t & 0xff) != 0;
    c[0] += pred && truth; c[1] += pred && !truth; c[2] += !pred && truth; c[3] += !pred && !truth;
  };
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(probs) + i);
    float tv[4];
    load_target4<DT>(targets, i, tv);
    add(x.x, tv[0]); add(x.y, tv[1]); add(x.z, tv[2]); add(x.w, tv[3]);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    add(probs[i], load_target<DT>(targets, i));
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
    unsigned long long t = 0;
    for (int w = 0; w < 8; ++w) t += s_i[w][threadIdx.x];
    atomicAdd(&counts[threadIdx.x], t);
  }
}

// ------------------------------------------------------------------------------------------- PR curve
constexpr int kMaxThr = 1024;
// hist[truth][idx], idx = number of thresholds strictly below sigmoid(x); stream-ordered scratch
__device__ unsigned long long g_pr_hist[2][kMaxThr + 1];

template <int DT>
__global__ void __launch_bounds__(256) pr_hist_kernel(const float* __restrict__ logits, const void* __restrict__ targets,
                                                      int64_t n, const float* __restrict__ thr, int n_thr) {
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
  for (int i = threadIdx.x; i < 2 * bins; i += blockDim.x) {
    unsigned long long t = 0;
    for (int w = 0; w < 8; ++w) t += hist[w * 2 * bins + i];
    if (t) atomicAdd(&g_pr_hist[i / bins][i % bins], t);
  }
}

__global__ void pr_clear_kernel(int n_thr) {
  for (int i = threadIdx.x; i < 2 * (kMaxThr + 1); i += blockDim.x) (&g_pr_hist[0][0])[i] = 0;
}

// pred_k = idx > k.  tp += sum_{idx>k} h[1][idx], fp += sum_{idx>k} h[0][idx], fn/tn the complements.
__global__ void pr_finalize_kernel(int n_thr, unsigned long long* tp, unsigned long long* tn, unsigned long long* fp,
                                   unsigned long long* fn) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_thr) return;
  unsigned long long pos1 = 0, pos0 = 0, neg1 = 0, neg0 = 0;
  for (int idx = 0; idx <= n_thr; ++idx) {
    if (idx > k) { pos1 += g_pr_hist[1][idx]; pos0 += g_pr_hist[0][idx]; }
    else { neg1 += g_pr_hist[1][idx]; neg0 += g_pr_hist[0][idx]; }
  }
  tp[k] += pos1; fp[k] += pos0; fn[k] += neg1; tn[k] += neg0;
}

static int reduce_grid(int64_t n_items) {
  const int64_t need = (n_items + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

// Gradient of c_bce * sum_i bce_i + c_jac * jaccard with respect to the logits (lib/losses.py:31-75 under autograd):
//   bce_i = BCE-with-logits(z, t), z = logsigmoid(x):  d/dx = (sigmoid(z) - t) * dz/dx = (p / (1 + p) - t) * (1 - p)
//   jaccard = 1 - A / D, A = sum p t + smooth, D = sum p + sum t - sum p t + smooth:
//             d/dp_i = (A (1 - t_i) - t_i D) / D^2,  dp/dx = p (1 - p)
// sums = {sum bce, sum p t, sum p, sum t} from snb_loss_iou_reduce of the same tensors; grad_out = upstream scalar
// gradient on the device (NULL = 1), so no host synchronisation is needed.  4 B + target read, 4 B written per element.
template <int DT>
__global__ void __launch_bounds__(256) loss_grad_kernel(const float* __restrict__ logits, const void* __restrict__ targets,
                                                        int64_t n, const double* __restrict__ sums,
                                                        const float* __restrict__ grad_out, float c_bce, float c_jac,
                                                        float smooth, float* __restrict__ grad) {
  const double A = sums[1] + smooth, D = sums[2] + sums[3] - sums[1] + smooth;
  const float g = grad_out ? __ldg(grad_out) : 1.f;
  const float jb = (float)(A / (D * D)) * c_jac * g;        // coefficient of (1 - t)
  const float jt = (float)(1.0 / D) * c_jac * g;            // coefficient of t
  const float cb = c_bce * g;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = __ldg(logits + i);
    const float t = load_target<DT>(targets, i);
    const float p = sigmoid_f32(x);
    const float q = 1.f - p;
    grad[i] = q * (cb * (p / (1.f + p) - t) + p * (jb * (1.f - t) - jt * t));
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

extern "C" int snb_loss_iou_reduce(const float* d_logits, const void* d_targets, int target_dtype, int64_t n,
                                   double* d_sums, int64_t* d_counts, void* stream) {
  if (!d_logits || !d_targets || !d_sums || !d_counts) return fail(SNB_E_INVALID, "snb_loss_iou_reduce: null argument");
  if (int rc = check_targets(d_targets, target_dtype, n, true)) return rc;
  if (reinterpret_cast<uintptr_t>(d_logits) & 15) return fail(SNB_E_INVALID, "logits must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  SNB_CUDA_CHECK(cudaMemsetAsync(d_sums, 0, 4 * sizeof(double), st));
  SNB_CUDA_CHECK(cudaMemsetAsync(d_counts, 0, 4 * sizeof(int64_t), st));
  if (n == 0) return SNB_OK;
  const int grid = reduce_grid((n + 3) / 4);
  unsigned long long* cnt = reinterpret_cast<unsigned long long*>(d_counts);
  if (target_dtype == SNB_DT_I64) loss_iou_kernel<SNB_DT_I64><<<grid, 256, 0, st>>>(d_logits, d_targets, n, d_sums, cnt);
  else if (target_dtype == SNB_DT_U8) loss_iou_kernel<SNB_DT_U8><<<grid, 256, 0, st>>>(d_logits, d_targets, n, d_sums, cnt);
  else loss_iou_kernel<SNB_DT_F32><<<grid, 256, 0, st>>>(d_logits, d_targets, n, d_sums, cnt);
  SNB_LAUNCH_CHECK();
  loss_iou_finalize<<<1, 1, 0, st>>>(cnt, (unsigned long long)n);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_loss_grad(const float* d_logits, const void* d_targets, int target_dtype, int64_t n,
                             const double* d_sums, const float* d_grad_out, float c_bce, float c_jac, float smooth,
                             float* d_grad_logits, void* stream) {
  if (!d_logits || !d_targets || !d_sums || !d_grad_logits) return fail(SNB_E_INVALID, "snb_loss_grad: null argument");
  if (int rc = check_targets(d_targets, target_dtype, n, false)) return rc;
  if (n == 0) return SNB_OK;
  cudaStream_t st = as_stream(stream);
  const int grid = reduce_grid(n);
  if (target_dtype == SNB_DT_I64)
    loss_grad_kernel<SNB_DT_I64><<<grid, 256, 0, st>>>(d_logits, d_targets, n, d_sums, d_grad_out, c_bce, c_jac, smooth, d_grad_logits);
  else if (target_dtype == SNB_DT_U8)
    loss_grad_kernel<SNB_DT_U8><<<grid, 256, 0, st>>>(d_logits, d_targets, n, d_sums, d_grad_out, c_bce, c_jac, smooth, d_grad_logits);
  else
    loss_grad_kernel<SNB_DT_F32><<<grid, 256, 0, st>>>(d_logits, d_targets, n, d_sums, d_grad_out, c_bce, c_jac, smooth, d_grad_logits);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_confusion_counts(const float* d_probs, const void* d_targets, int target_dtype, int64_t n, float thr,
                                    int64_t* d_counts, void* stream) {
  if (!d_probs || !d_targets || !d_counts) return fail(SNB_E_INVALID, "snb_confusion_counts: null argument");
  if (int rc = check_targets(d_targets, target_dtype, n, true)) return rc;
  if (reinterpret_cast<uintptr_t>(d_probs) & 15) return fail(SNB_E_INVALID, "probs must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  SNB_CUDA_CHECK(cudaMemsetAsync(d_counts, 0, 4 * sizeof(int64_t), st));
  if (n == 0) return SNB_OK;
  const int grid = reduce_grid((n + 3) / 4);
  unsigned long long* cnt = reinterpret_cast<unsigned long long*>(d_counts);
  if (target_dtype == SNB_DT_I64) confusion_kernel<SNB_DT_I64><<<grid, 256, 0, st>>>(d_probs, d_targets, n, thr, cnt);
  else if (target_dtype == SNB_DT_U8) confusion_kernel<SNB_DT_U8><<<grid, 256, 0, st>>>(d_probs, d_targets, n, thr, cnt);
  else confusion_kernel<SNB_DT_F32><<<grid, 256, 0, st>>>(d_probs, d_targets, n, thr, cnt);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_pr_curve_update(const float* d_logits, const void* d_targets, int target_dtype, int64_t n,
                                   const float* d_thresholds, int64_t n_thr, uint64_t* d_tp, uint64_t* d_tn,
                                   uint64_t* d_fp, uint64_t* d_fn, void* stream) {
  if (!d_logits || !d_targets || !d_thresholds || !d_tp || !d_tn || !d_fp || !d_fn)
    return fail(SNB_E_INVALID, "snb_pr_curve_update: null argument");
  if (n_thr < 1 || n_thr > kMaxThr) return fail(SNB_E_INVALID, "n_thr=%lld not in [1, %d]", (long long)n_thr, kMaxThr);
  if (int rc = check_targets(d_targets, target_dtype, n, false)) return rc;
  if (n == 0) return SNB_OK;
  cudaStream_t st = as_stream(stream);
  const int bins = (int)n_thr + 1;
  const size_t smem = (size_t)(8 * 2 * bins) * sizeof(uint32_t) + (size_t)n_thr * sizeof(float);
  pr_clear_kernel<<<1, 256, 0, st>>>((int)n_thr);
  const int grid = reduce_grid(n);
#define SNB_PR(DT)                                                                                         \
  do {                                                                                                     \
    if (smem > 48 * 1024)                                                                                  \
      SNB_CUDA_CHECK(cudaFuncSetAttribute(pr_hist_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)smem));                                                     \
    pr_hist_kernel<DT><<<grid, 256, smem, st>>>(d_logits, d_targets, n, d_thresholds, (int)n_thr);          \
  } while (0)
  if (target_dtype == SNB_DT_I64) SNB_PR(SNB_DT_I64);
  else if (target_dtype == SNB_DT_U8) SNB_PR(SNB_DT_U8);
  else SNB_PR(SNB_DT_F32);
#undef SNB_PR
  SNB_LAUNCH_CHECK();
  pr_finalize_kernel<<<((int)n_thr + 127) / 128, 128, 0, st>>>(
      (int)n_thr, reinterpret_cast<unsigned long long*>(d_tp), reinterpret_cast<unsigned long long*>(d_tn),
      reinterpret_cast<unsigned long long*>(d_fp), reinterpret_cast<unsigned long long*>(d_fn));
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
