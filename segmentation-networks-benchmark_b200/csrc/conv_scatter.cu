// conv3x3 (stride 1, padding 1) with a NARROW output (Cout = 16: FCDenseNet's growth-rate layers,
// lib/models/tiramisu.py:9-19) as ONE wide GEMM per tile instead of nine narrow ones.
//
// The tap-list kernels of conv_tcgen05.cu issue, per K chunk, one MMA group per tap with N = Cout.  A tcgen05.mma
// M=128 x K=16 retires in max(N/2, 32 + N/4) cycles (tools/micro/mma_rate.cu), so at N = 32 (16 real channels) every
// MMA pays 40 cycles for 8 cycles of useful work and the dense layers run at ~10 % of the tensor peak.  Here the taps
// move into the N dimension:
//     P[p][tap*16 + co] = sum_ci  X[p][ci] * W[co][ci][tap]          one GEMM, M = 128 input pixels, N = 144, K = Cin
//     out[q][co]        = bias[co] + sum_tap P[q + (dy,dx)(tap)][tap*16 + co]
// i.e. every input pixel is multiplied ONCE with all nine filter taps (N = 144 -> 72 cycles per MMA, 4x the useful work
// per cycle), and the epilogue adds the nine shifted partial planes.  A tile is an 8 x 16 patch of INPUT pixels (plain
// TMA box, out-of-bounds zero fill = the conv padding); its 6 x 14 interior pixels are the outputs it owns, so tiles
// step by (6, 14) and 34 % of the MMA rows are halo recompute -- still ~3.3x fewer tensor cycles per output.
//
// Warp roles (256 threads; 512 with the pre-activation prologue): 0 TMA producer (activation ring + weights: resident
// when all K chunks fit next to the pipeline, else streamed through their own ring), 1 MMA issuer, 2 TMEM allocator,
// 4-7 epilogue (tcgen05.ld -> fp32 partial planes in shared memory -> shifted 9-term sum + bias -> bf16 -> 32-byte
// stores at the slab's channel offset), 8-15 prologue: y = relu(x * scale[c] + shift[c]) rewritten into the staged
// activation tile (FCDenseNet's per-consumer BatchNorm + ReLU), zero outside the image.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>

#include "sm100_ptx.cuh"
#include "snb_internal.h"

namespace snb {

int encode_map(CUtensorMap* map, void* base, int rank, const uint64_t* dims, const uint64_t* strides, const uint32_t* box,
               int swizzle_bytes, int elem_bytes);

constexpr int kScM = 128;                 // MMA rows = input pixels of the 8 x 16 patch
constexpr int kScPW = 8, kScPH = 16;      // patch
constexpr int kScOW = 6, kScOH = 14;      // outputs owned by a tile
constexpr int kScCout = 16;
constexpr int kScN = 9 * kScCout;         // 144 GEMM columns
constexpr int kScAcc = 256;               // TMEM columns per accumulator stage (144 used)
constexpr int kScMaxStages = 8;
constexpr int kScPBytes = 9 * kScM * kScCout * 4;   // fp32 partial planes [tap][channel quad][pixel] of float4
constexpr int kScSmemBudget = 227 * 1024;

struct alignas(64) ScatterParams {
  CUtensorMap map_a;   // activations (C, W, H, N), box (BK, 8, 16, 1)
  CUtensorMap map_b;   // weights (Cin, 144, 1), box (BK, 144, 1)
  int32_t k_chunks, tiles_x, tiles_y, n_img, total_tiles;
  int32_t w, h;
  int32_t a_stages, b_stages, bres;
  const float* bias;
  const float* pre_scale;
  const float* pre_shift;
  __nv_bfloat16* out;
  int64_t out_cstride;
  const __nv_bfloat16* src;   // TS variant: the activation slab itself (read with plain loads by the prologue warps)
  int64_t src_cstride;
};

struct ScTile {
  int x0, y0, img;   // origin of the OUTPUT block; the input patch starts at (x0 - 1, y0 - 1)
};

__device__ __forceinline__ ScTile sc_decode(const ScatterParams& p, int t) {
  ScTile c;
  c.x0 = (t % p.tiles_x) * kScOW;
  t /= p.tiles_x;
  c.y0 = (t % p.tiles_y) * kScOH;
  c.img = t / p.tiles_y;
  return c;
}

__device__ __forceinline__ uint32_t sc_pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// smem: [a_stages x A_BYTES][B: b_stages (or all k_chunks when resident) x B_BYTES][P: 73,728 B][ctrl]
// -DSNB_CONV_PROFILE (profiling builds only, tools/build_rev.py --profile): wait cycles of the roles, as in conv_tcgen05.cu.
// [0] issuer: activation stage, [1] prologue warp 8: raw stage landed, [2] issuer: accumulator stage, [3] issuer loop,
// [4] epilogue warp 4: accumulator full, [5] epilogue loop, [6] prologue loop, [7] tiles.
#ifdef SNB_CONV_PROFILE
__device__ unsigned long long g_scatter_prof[12];   // [8] TS prologue: TMEM stage free, [9] TS prologue: tcgen05.st + wait + arrive
#define SC_PROF_DECL long long prof_t = 0; unsigned long long prof_c[4] = {0, 0, 0, 0}; (void)prof_t;
#define SC_PROF_T0 prof_t = clock64();
#define SC_PROF_ADD(i) prof_c[i] += static_cast<unsigned long long>(clock64() - prof_t);
#define SC_PROF_FLUSH(dst, i) atomicAdd(&g_scatter_prof[dst], prof_c[i]);
#else
#define SC_PROF_DECL
#define SC_PROF_T0
#define SC_PROF_ADD(i)
#define SC_PROF_FLUSH(dst, i)
#endif

template <int BK, bool PRE>
__global__ void __launch_bounds__(PRE ? 512 : 256, 1) conv_scatter_kernel(const __grid_constant__ ScatterParams p) {
  constexpr int SWZ = BK * 2;
  constexpr int A_BYTES = kScM * SWZ;
  constexpr int B_BYTES = kScN * SWZ;            // 144 rows: a multiple of 1024 for SWZ = 128 and 64
  constexpr uint32_t IDESC = make_idesc(kScM, kScN, 1u);
  static_assert(B_BYTES % 1024 == 0 && A_BYTES % 1024 == 0, "swizzle atom alignment");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int n_b_slots = p.bres ? p.k_chunks : p.b_stages;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + p.a_stages * A_BYTES;
  float* smem_p = reinterpret_cast<float*>(smem_b + n_b_slots * B_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(smem_p) + kScPBytes);
  uint64_t* a_full = bars;                          // [kScMaxStages]
  uint64_t* a_empty = a_full + kScMaxStages;
  uint64_t* b_full = a_empty + kScMaxStages;        // b_full[0] doubles as the resident-weights barrier
  uint64_t* b_empty = b_full + kScMaxStages;
  uint64_t* a_ready = b_empty + kScMaxStages;       // stage rewritten by the prologue warps (PRE)
  uint64_t* tmem_full = a_ready + kScMaxStages;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;             // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_a);
    tma_prefetch_desc(&p.map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kScMaxStages; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
      mbar_init(&a_ready[i], 8);   // one arrival per prologue warp
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, 2 * kScAcc);
    tmem_relinquish();
  }
  tc05_fence_before();
  __syncthreads();
  tc05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      if (p.bres) {
        mbar_arrive_expect_tx(&b_full[0], static_cast<uint32_t>(p.k_chunks) * B_BYTES);
        for (int kc = 0; kc < p.k_chunks; ++kc)
          tma_load_3d(&p.map_b, &b_full[0], smem_b + kc * B_BYTES, kc * BK, 0, 0);
      }
      uint32_t sa = 0, pa = 1, sb = 0, pb = 1;   // waiting on parity 1 of a fresh barrier returns immediately
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const ScTile tc = sc_decode(p, t);
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          mbar_wait(&a_empty[sa], pa);
          mbar_arrive_expect_tx(&a_full[sa], A_BYTES);
          tma_load_4d(&p.map_a, &a_full[sa], smem_a + sa * A_BYTES, kc * BK, tc.x0 - 1, tc.y0 - 1, tc.img);
          if (++sa == static_cast<uint32_t>(p.a_stages)) { sa = 0; pa ^= 1; }
          if (!p.bres) {
            mbar_wait(&b_empty[sb], pb);
            mbar_arrive_expect_tx(&b_full[sb], B_BYTES);
            tma_load_3d(&p.map_b, &b_full[sb], smem_b + sb * B_BYTES, kc * BK, 0, 0);
            if (++sb == static_cast<uint32_t>(p.b_stages)) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (single thread)
    if (elect_one()) {
      const bool bres = p.bres != 0;
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0, local_tile = 0;
      if (bres) {
        mbar_wait(&b_full[0], 0);
        tc05_fence_after();
      }
      SC_PROF_DECL
#ifdef SNB_CONV_PROFILE
      const long long prof_loop0 = clock64();
#endif
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++local_tile) {
        const uint32_t acc = local_tile & 1;
        SC_PROF_T0
        mbar_wait(&tmem_empty[acc], ((local_tile >> 1) & 1) ^ 1);
        SC_PROF_ADD(2)
        tc05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kScAcc;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          SC_PROF_T0
          mbar_wait(PRE ? &a_ready[sa] : &a_full[sa], pa);
          SC_PROF_ADD(0)
          if (!bres) mbar_wait(&b_full[sb], pb);
          tc05_fence_after();
          const uint64_t adesc = make_kmajor_desc<SWZ>(smem_u32(smem_a + sa * A_BYTES), 8 * SWZ);
          const uint64_t bdesc = make_kmajor_desc<SWZ>(smem_u32(smem_b + (bres ? kc : (int)sb) * B_BYTES), 8 * SWZ);
#pragma unroll
          for (int k = 0; k < SWZ / 32; ++k)
            umma_ss<2>(adesc + 2 * k, bdesc + 2 * k, d_tmem, IDESC, (kc > 0 || k > 0) ? 1u : 0u);
          umma_commit(&a_empty[sa]);
          if (++sa == static_cast<uint32_t>(p.a_stages)) { sa = 0; pa ^= 1; }
          if (!bres) {
            umma_commit(&b_empty[sb]);
            if (++sb == static_cast<uint32_t>(p.b_stages)) { sb = 0; pb ^= 1; }
          }
        }
        umma_commit(&tmem_full[acc]);
      }
#ifdef SNB_CONV_PROFILE
      prof_c[3] = static_cast<unsigned long long>(clock64() - prof_loop0);
      SC_PROF_FLUSH(0, 0) SC_PROF_FLUSH(2, 2) SC_PROF_FLUSH(3, 3)
      atomicAdd(&g_scatter_prof[7], static_cast<unsigned long long>(local_tile));
#endif
    }
  } else if (PRE && warp >= 8) {
    // ------------------------------------------------------------------ prologue: pre-activation of the A operand
    constexpr int CPR = SWZ / 16;            // 16-byte chunks (8 channels) per operand row
    constexpr int RSTEP = 256 / CPR;         // rows covered by the 256 prologue threads per sweep
    const int tt = threadIdx.x - 256;
    const int jl = tt % CPR;                 // logical chunk = channels [8 jl, 8 jl + 8) of the K chunk
    const int r0 = tt / CPR;
    uint32_t s = 0, par = 0;
    SC_PROF_DECL
#ifdef SNB_CONV_PROFILE
    const long long prof_loop0 = clock64();
#endif
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const ScTile tc = sc_decode(p, t);
      for (int kc = 0; kc < p.k_chunks; ++kc) {
        const float4* sc4 = reinterpret_cast<const float4*>(p.pre_scale + kc * BK + jl * 8);
        const float4* sh4 = reinterpret_cast<const float4*>(p.pre_shift + kc * BK + jl * 8);
        const float4 s0 = __ldg(sc4), s1 = __ldg(sc4 + 1), b0 = __ldg(sh4), b1 = __ldg(sh4 + 1);
        SC_PROF_T0
        mbar_wait(&a_full[s], par);
        SC_PROF_ADD(1)
        uint8_t* base = smem_a + s * A_BYTES;
#pragma unroll 2
        for (int r = r0; r < kScM; r += RSTEP) {
          const int hy = r / kScPW, hx = r - hy * kScPW;
          const bool inside = static_cast<unsigned>(tc.x0 - 1 + hx) < static_cast<unsigned>(p.w) &&
                              static_cast<unsigned>(tc.y0 - 1 + hy) < static_cast<unsigned>(p.h);
          const int phys = jl ^ (SWZ == 128 ? (r & 7) : ((r >> 1) & 3));   // the swizzle TMA applied to this row
          uint4* ptr = reinterpret_cast<uint4*>(base + r * SWZ + (phys << 4));
          uint4 o = make_uint4(0u, 0u, 0u, 0u);
          if (inside) {
            const uint4 v = *ptr;
            const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&v);
            const float2 x0 = __bfloat1622float2(pv[0]), x1 = __bfloat1622float2(pv[1]);
            const float2 x2 = __bfloat1622float2(pv[2]), x3 = __bfloat1622float2(pv[3]);
            o.x = sc_pack_bf16x2(fmaxf(fmaf(x0.x, s0.x, b0.x), 0.f), fmaxf(fmaf(x0.y, s0.y, b0.y), 0.f));
            o.y = sc_pack_bf16x2(fmaxf(fmaf(x1.x, s0.z, b0.z), 0.f), fmaxf(fmaf(x1.y, s0.w, b0.w), 0.f));
            o.z = sc_pack_bf16x2(fmaxf(fmaf(x2.x, s1.x, b1.x), 0.f), fmaxf(fmaf(x2.y, s1.y, b1.y), 0.f));
            o.w = sc_pack_bf16x2(fmaxf(fmaf(x3.x, s1.z, b1.z), 0.f), fmaxf(fmaf(x3.y, s1.w, b1.w), 0.f));
          }
          *ptr = o;
        }
        fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[s]);
        if (++s == static_cast<uint32_t>(p.a_stages)) { s = 0; par ^= 1; }
      }
    }
#ifdef SNB_CONV_PROFILE
    if (threadIdx.x == 256) {
      prof_c[2] = static_cast<unsigned long long>(clock64() - prof_loop0);
      SC_PROF_FLUSH(1, 1) SC_PROF_FLUSH(6, 2)
    }
#endif
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------------------------------------------ epilogue (128 threads = 128 patch pixels)
    const int q = warp & 3;
    const int row = q * 32 + lane;                     // patch pixel (row % 8, row / 8)
    const int px = row % kScPW, py = row / kScPW;
    const bool interior = px >= 1 && px <= kScOW && py >= 1 && py <= kScOH;
    float bias[kScCout];
#pragma unroll
    for (int c = 0; c < kScCout; c += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c));
      bias[c] = b.x; bias[c + 1] = b.y; bias[c + 2] = b.z; bias[c + 3] = b.w;
    }
    uint32_t local_tile = 0;
    SC_PROF_DECL
#ifdef SNB_CONV_PROFILE
    const long long prof_loop0 = clock64();
#endif
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++local_tile) {
      const ScTile tc = sc_decode(p, t);
      const uint32_t acc = local_tile & 1;
      SC_PROF_T0
      mbar_wait(&tmem_full[acc], (local_tile >> 1) & 1);
      SC_PROF_ADD(0)
      tc05_fence_after();
      const uint32_t t_addr = tmem_base + acc * kScAcc + (static_cast<uint32_t>(q * 32) << 16);
      // partial planes: P[tap][channel quad j][pixel] as float4 (fp32): consecutive pixels (= lanes) are 16 bytes
      // apart, so the stores here and the shifted loads below are bank-conflict free
      float4* p4 = reinterpret_cast<float4*>(smem_p);
#pragma unroll 3
      for (int tap = 0; tap < 9; ++tap) {
        uint32_t v[16];
        tmem_ld_32x16(t_addr + tap * kScCout, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          p4[(tap * 4 + j) * kScM + row] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                      __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
      }
      // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp
      tc05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      SC_PROF_T0
      named_bar_sync(1, 128);
      SC_PROF_ADD(2)
      const int ox = tc.x0 + px - 1, oy = tc.y0 + py - 1;
      if (interior && ox < p.w && oy < p.h) {
        float o[kScCout];
#pragma unroll
        for (int c = 0; c < kScCout; ++c) o[c] = bias[c];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          // out(q) += P[tap][q + (ky-1, kx-1)]: the partial product of the input pixel this tap reads
          const int src = row + (tap / 3 - 1) * kScPW + (tap % 3 - 1);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 a = p4[(tap * 4 + j) * kScM + src];
            o[4 * j] += a.x; o[4 * j + 1] += a.y; o[4 * j + 2] += a.z; o[4 * j + 3] += a.w;
          }
        }
        __nv_bfloat16* dst = p.out + ((static_cast<int64_t>(tc.img) * p.h + oy) * p.w + ox) * p.out_cstride;
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        d4[0] = make_uint4(sc_pack_bf16x2(o[0], o[1]), sc_pack_bf16x2(o[2], o[3]), sc_pack_bf16x2(o[4], o[5]),
                           sc_pack_bf16x2(o[6], o[7]));
        d4[1] = make_uint4(sc_pack_bf16x2(o[8], o[9]), sc_pack_bf16x2(o[10], o[11]), sc_pack_bf16x2(o[12], o[13]),
                           sc_pack_bf16x2(o[14], o[15]));
      }
      SC_PROF_T0
      named_bar_sync(1, 128);   // the planes may be overwritten by the next tile
      SC_PROF_ADD(2)
    }
#ifdef SNB_CONV_PROFILE
    if (warp == 4 && lane == 0) {
      prof_c[1] = static_cast<unsigned long long>(clock64() - prof_loop0);
      SC_PROF_FLUSH(4, 0) SC_PROF_FLUSH(5, 1)
    }
#endif
  }

  tc05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc05_fence_after();
    tmem_dealloc(tmem_base, 2 * kScAcc);
  }
}

// ================================================================================== TS variant (fused pre-activation)
// The role profile of the kernel above (tools/conv_wait_profile.py) shows the issuer waiting on the prologue warps 60 % of
// its time, and the prologue busy ~1400 cycles per 16 KB chunk: per tile, TMA fill (64 KB) + prologue read and rewrite
// (128 KB) + MMA operand reads (138 KB) + partial planes (122 KB) = 452 KB through a 128 B/clk shared memory -- the wall is
// the shared-memory bandwidth, and two thirds of it is the activation operand going through shared memory three times.
// Here the operand makes ONE pass through shared memory: TMA stages the raw activation tile (deep, coalesced, asynchronous
// prefetch); the prologue warps (one thread per patch pixel) read their row into registers, apply y = relu(x * scale +
// shift) and write the bf16 row into TENSOR MEMORY with tcgen05.st; the MMA takes its A operand from there (tcgen05.mma
// with A in TMEM, the FlashAttention "P @ V" form), so neither the rewrite nor the MMA's A reads touch shared memory
// (452 -> 324 KB per tile).  (A first version loaded the rows straight from global memory into registers: 32 lanes x 16
// bytes 608 bytes apart is 32 L1 wavefronts per load instruction, 4600 per tile -- slower than the kernel above.)
// TMEM: accumulator stages at columns 0 / 192 (144 used), four activation stages of BK / 2 columns from column 384.  Two
// prologue groups (warps 8-11 / 12-15) take alternate chunks.
constexpr int kTsAccStride = 192;
constexpr int kTsACol0 = 384;
constexpr int kTsNA = 4;

template <int BK>
__global__ void __launch_bounds__(512, 1) conv_scatter_ts_kernel(const __grid_constant__ ScatterParams p) {
  constexpr int SWZ = BK * 2;
  constexpr int B_BYTES = kScN * SWZ;
  constexpr int ACOLS = BK / 2;                    // 32-bit TMEM columns per activation stage
  constexpr int NV = BK / 8;                       // 16-byte vectors per pixel and chunk
  constexpr uint32_t IDESC = make_idesc(kScM, kScN, 1u);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int A_BYTES = kScM * SWZ;
  const int n_b_slots = p.bres ? p.k_chunks : p.b_stages;
  uint8_t* smem_a = smem;                           // a_stages raw activation tiles (TMA, swizzled)
  uint8_t* smem_b = smem_a + p.a_stages * A_BYTES;
  float* smem_p = reinterpret_cast<float*>(smem_b + n_b_slots * B_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(smem_p) + kScPBytes);
  uint64_t* b_full = bars;                          // [kScMaxStages]; b_full[0] doubles as the resident-weights barrier
  uint64_t* b_empty = b_full + kScMaxStages;
  uint64_t* s_full = b_empty + kScMaxStages;        // [kScMaxStages] raw tile landed in shared memory
  uint64_t* s_empty = s_full + kScMaxStages;        // [kScMaxStages] ... and read into registers by its prologue group
  uint64_t* a_ready = s_empty + kScMaxStages;       // [kTsNA] activation stage written to TMEM
  uint64_t* a_empty = a_ready + kTsNA;              // [kTsNA] ... and consumed by the MMAs
  uint64_t* tmem_full = a_empty + kTsNA;            // [2]
  uint64_t* tmem_empty = tmem_full + 2;             // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_pre = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 1024);   // [2][k_chunks * BK] scale, shift

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nc = p.k_chunks * BK;
  for (int i = threadIdx.x; i < nc; i += blockDim.x) {
    s_pre[i] = __ldg(p.pre_scale + i);
    s_pre[nc + i] = __ldg(p.pre_shift + i);
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_a);
    tma_prefetch_desc(&p.map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kScMaxStages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);   // one arrival per warp of the prologue group that read the stage
    }
    for (int i = 0; i < kTsNA; ++i) {
      mbar_init(&a_ready[i], 4);   // one arrival per warp of the prologue group that wrote the stage
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc05_fence_before();
  __syncthreads();
  tc05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ raw activation tiles (+ resident weights)
    if (elect_one()) {
      if (p.bres) {
        mbar_arrive_expect_tx(&b_full[0], static_cast<uint32_t>(p.k_chunks) * B_BYTES);
        for (int kc = 0; kc < p.k_chunks; ++kc) tma_load_3d(&p.map_b, &b_full[0], smem_b + kc * B_BYTES, kc * BK, 0, 0);
      }
      uint32_t ss = 0, ps = 1;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const ScTile tc = sc_decode(p, t);
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          mbar_wait(&s_empty[ss], ps);
          mbar_arrive_expect_tx(&s_full[ss], A_BYTES);
          tma_load_4d(&p.map_a, &s_full[ss], smem_a + ss * A_BYTES, kc * BK, tc.x0 - 1, tc.y0 - 1, tc.img);
          if (++ss == static_cast<uint32_t>(p.a_stages)) { ss = 0; ps ^= 1; }
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ streamed weights
    if (!p.bres && elect_one()) {
      uint32_t sb = 0, pb = 1;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          mbar_wait(&b_empty[sb], pb);
          mbar_arrive_expect_tx(&b_full[sb], B_BYTES);
          tma_load_3d(&p.map_b, &b_full[sb], smem_b + sb * B_BYTES, kc * BK, 0, 0);
          if (++sb == static_cast<uint32_t>(p.b_stages)) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: A from tensor memory, B from shared memory
    if (elect_one()) {
      const bool bres = p.bres != 0;
      uint32_t sb = 0, pb = 0, local_tile = 0, n = 0;
      if (bres) {
        mbar_wait(&b_full[0], 0);
        tc05_fence_after();
      }
      SC_PROF_DECL
#ifdef SNB_CONV_PROFILE
      const long long prof_loop0 = clock64();
#endif
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++local_tile) {
        const uint32_t acc = local_tile & 1;
        SC_PROF_T0
        mbar_wait(&tmem_empty[acc], ((local_tile >> 1) & 1) ^ 1);
        SC_PROF_ADD(2)
        tc05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kTsAccStride;
        for (int kc = 0; kc < p.k_chunks; ++kc, ++n) {
          const uint32_t sa = n % kTsNA;
          SC_PROF_T0
          mbar_wait(&a_ready[sa], (n / kTsNA) & 1);
          SC_PROF_ADD(0)
          if (!bres) mbar_wait(&b_full[sb], pb);
          tc05_fence_after();
          const uint32_t a_tmem = tmem_base + kTsACol0 + sa * ACOLS;
          const uint64_t bdesc = make_kmajor_desc<SWZ>(smem_u32(smem_b + (bres ? kc : (int)sb) * B_BYTES), 8 * SWZ);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16_ts(a_tmem + 8 * k, bdesc + 2 * k, d_tmem, IDESC, (kc > 0 || k > 0) ? 1u : 0u);
          umma_commit(&a_empty[sa]);
          if (!bres) {
            umma_commit(&b_empty[sb]);
            if (++sb == static_cast<uint32_t>(p.b_stages)) { sb = 0; pb ^= 1; }
          }
        }
        umma_commit(&tmem_full[acc]);
      }
#ifdef SNB_CONV_PROFILE
      prof_c[3] = static_cast<unsigned long long>(clock64() - prof_loop0);
      SC_PROF_FLUSH(0, 0) SC_PROF_FLUSH(2, 2) SC_PROF_FLUSH(3, 3)
      atomicAdd(&g_scatter_prof[7], static_cast<unsigned long long>(local_tile));
#endif
    }
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ prologue: shared memory -> registers -> BN + ReLU -> TMEM
    const int grp = (warp - 8) >> 2;                 // chunks n = grp, grp + 2, ... of this CTA's (tile, chunk) sequence
    const int q = warp & 3;                          // TMEM lane quarter this warp may write
    const int row = q * 32 + lane;                   // patch pixel (row % 8, row / 8)
    const int px = row % kScPW, py = row / kScPW;
    const int sw = SWZ == 128 ? (row & 7) : ((row >> 1) & 3);   // the swizzle TMA applied to this row
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t a_stages = p.a_stages;
    int kc = grp;
    // ring positions advance by two chunks with compare-and-wrap, the tile is decoded once per tile: a division per chunk
    // (sc_decode, n % a_stages) is a ~100-cycle dependent chain in front of every stage
    uint32_t ss = grp % a_stages, ps = 0;            // shared-memory stage / parity of chunk n (a_stages is even)
    uint32_t sa = grp, pa = 1;                       // TMEM stage / parity of its "free" barrier (kTsNA = 4)
    SC_PROF_DECL
#ifdef SNB_CONV_PROFILE
    const long long prof_loop0 = clock64();
#endif
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, kc -= p.k_chunks) {
     const ScTile tc = sc_decode(p, t);
     const bool inside = static_cast<unsigned>(tc.x0 - 1 + px) < static_cast<unsigned>(p.w) &&
                         static_cast<unsigned>(tc.y0 - 1 + py) < static_cast<unsigned>(p.h);
     for (; kc < p.k_chunks; kc += 2) {
      SC_PROF_T0
      mbar_wait(&s_full[ss], ps);
      SC_PROF_ADD(1)
      uint4 cur[NV];
      const uint8_t* base = smem_a + ss * A_BYTES + row * SWZ;
#pragma unroll
      for (int j = 0; j < NV; ++j) cur[j] = *reinterpret_cast<const uint4*>(base + ((j ^ sw) << 4));
      uint32_t o[BK / 2];
      const float4* sc4 = reinterpret_cast<const float4*>(s_pre + kc * BK);
      const float4* sh4 = reinterpret_cast<const float4*>(s_pre + nc + kc * BK);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 s0 = sc4[2 * j], s1 = sc4[2 * j + 1], b0 = sh4[2 * j], b1 = sh4[2 * j + 1];
        const uint32_t w[4] = {cur[j].x, cur[j].y, cur[j].z, cur[j].w};
        const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float lo = fmaf(__uint_as_float(w[e] << 16), sv[2 * e], bv[2 * e]);
          const float hi = fmaf(__uint_as_float(w[e] & 0xffff0000u), sv[2 * e + 1], bv[2 * e + 1]);
          uint32_t d;
          asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));   // ReLU folded into the conversion
          o[4 * j + e] = inside ? d : 0u;           // outside the image: the convolution's zero padding, AFTER the activation
        }
      }
      // The raw stage goes back to TMA only now: o[] depends on every loaded vector, so the shared-memory reads have
      // completed.  (With the arrive right behind the loads the next TMA fill overtook them now and then: the parity tests
      // caught it as a doubled error on one run in three.)  The proxy fence orders these generic-proxy reads before the
      // async-proxy write that the producer issues after it has seen the arrive.
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[ss]);
      SC_PROF_T0
      mbar_wait(&a_empty[sa], pa);
      SC_PROF_ADD(0)
      tc05_fence_after();
      const uint32_t a_tmem = tmem_base + kTsACol0 + sa * ACOLS + lane_addr;
      SC_PROF_T0
      if constexpr (BK == 64) tmem_st_32x32(a_tmem, o);
      else tmem_st_32x16(a_tmem, o);
      tmem_st_wait();
      tc05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_ready[sa]);
      SC_PROF_ADD(3)
      ss += 2;
      if (ss >= a_stages) { ss -= a_stages; ps ^= 1; }
      sa += 2;
      if (sa >= kTsNA) { sa -= kTsNA; pa ^= 1; }
     }
    }
#ifdef SNB_CONV_PROFILE
    if (threadIdx.x == 256) {
      prof_c[2] = static_cast<unsigned long long>(clock64() - prof_loop0);
      SC_PROF_FLUSH(1, 1) SC_PROF_FLUSH(6, 2) SC_PROF_FLUSH(8, 0) SC_PROF_FLUSH(9, 3)
    }
#endif
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------------------------------------------ epilogue: shifted sum of the nine partial planes
    // out(q) = bias + sum_{dy,dx} P[dy,dx](q + 8 dy + dx).  The kernel above sends all nine fp32 planes through shared
    // memory (73 KB written + read per tile, the largest shared-memory stream of the kernel).  Here the dx shifts are warp
    // shuffles (a warp holds 4 patch rows of 8 pixels: the x neighbours of an interior pixel are lanes +-1) and only the
    // two row sums S(-1), S(+1) that have to move by one patch ROW (8 lanes: across warps at the warp's first / last row)
    // go through shared memory: 2 x 8 KB.
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int px = row % kScPW, py = row / kScPW;
    const bool interior = px >= 1 && px <= kScOW && py >= 1 && py <= kScOH;
    float bias[kScCout];
#pragma unroll
    for (int c = 0; c < kScCout; c += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c));
      bias[c] = b.x; bias[c + 1] = b.y; bias[c + 2] = b.z; bias[c + 3] = b.w;
    }
    float4* p4 = reinterpret_cast<float4*>(smem_p);      // [2 planes][4 channel quads][128 pixels] of float4
    uint32_t local_tile = 0;
    SC_PROF_DECL
#ifdef SNB_CONV_PROFILE
    const long long prof_loop0 = clock64();
#endif
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++local_tile) {
      const ScTile tc = sc_decode(p, t);
      const uint32_t acc = local_tile & 1;
      SC_PROF_T0
      mbar_wait(&tmem_full[acc], (local_tile >> 1) & 1);
      SC_PROF_ADD(0)
      tc05_fence_after();
      const uint32_t t_addr = tmem_base + acc * kTsAccStride + (static_cast<uint32_t>(q * 32) << 16);
      float o[kScCout];
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        uint32_t vl[16], vc[16], vr[16];
        tmem_ld_32x16(t_addr + (dy * 3 + 0) * kScCout, vl);     // tap (dy, dx = -1): wanted from the pixel on the left
        tmem_ld_32x16(t_addr + (dy * 3 + 1) * kScCout, vc);
        tmem_ld_32x16(t_addr + (dy * 3 + 2) * kScCout, vr);     // tap (dy, dx = +1): from the pixel on the right
        tmem_ld_wait();
        float sdy[kScCout];
#pragma unroll
        for (int c = 0; c < kScCout; ++c) {
          const float l = __shfl_up_sync(0xffffffffu, __uint_as_float(vl[c]), 1);
          const float r = __shfl_down_sync(0xffffffffu, __uint_as_float(vr[c]), 1);
          sdy[c] = (l + __uint_as_float(vc[c])) + r;            // (garbage at px = 0 / 7: those pixels are not outputs)
        }
        if (dy == 1) {
#pragma unroll
          for (int c = 0; c < kScCout; ++c) o[c] = bias[c] + sdy[c];
        } else {
          // S(dy = -1) is wanted by the pixel one patch row BELOW, S(+1) by the one above: plane 0 / 1
          float4* pl = p4 + (dy == 0 ? 0 : 4 * kScM);
#pragma unroll
          for (int j = 0; j < 4; ++j) pl[j * kScM + row] = make_float4(sdy[4 * j], sdy[4 * j + 1], sdy[4 * j + 2], sdy[4 * j + 3]);
        }
      }
      // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp
      tc05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      named_bar_sync(1, 128);
      const int ox = tc.x0 + px - 1, oy = tc.y0 + py - 1;
      if (interior && ox < p.w && oy < p.h) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 a = p4[j * kScM + row - kScPW];               // S(-1) of the pixel above
          const float4 b = p4[(4 + j) * kScM + row + kScPW];         // S(+1) of the pixel below
          o[4 * j] += a.x + b.x; o[4 * j + 1] += a.y + b.y; o[4 * j + 2] += a.z + b.z; o[4 * j + 3] += a.w + b.w;
        }
        __nv_bfloat16* dst = p.out + ((static_cast<int64_t>(tc.img) * p.h + oy) * p.w + ox) * p.out_cstride;
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        d4[0] = make_uint4(sc_pack_bf16x2(o[0], o[1]), sc_pack_bf16x2(o[2], o[3]), sc_pack_bf16x2(o[4], o[5]),
                           sc_pack_bf16x2(o[6], o[7]));
        d4[1] = make_uint4(sc_pack_bf16x2(o[8], o[9]), sc_pack_bf16x2(o[10], o[11]), sc_pack_bf16x2(o[12], o[13]),
                           sc_pack_bf16x2(o[14], o[15]));
      }
      named_bar_sync(1, 128);   // the planes may be overwritten by the next tile
    }
#ifdef SNB_CONV_PROFILE
    if (warp == 4 && lane == 0) {
      prof_c[1] = static_cast<unsigned long long>(clock64() - prof_loop0);
      SC_PROF_FLUSH(4, 0) SC_PROF_FLUSH(5, 1)
    }
#endif
  }

  tc05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace snb

struct snb_conv_scatter {
  snb::ScatterParams params;
  const void* fn;
  int smem, grid, threads;
  double flops;
};

using namespace snb;

extern "C" int snb_conv_scatter_create(const void* d_in, int64_t n, int64_t h, int64_t w, int64_t cin, int64_t in_cstride,
                                       const void* d_weight, const float* d_bias, const float* d_pre_scale,
                                       const float* d_pre_shift, void* d_out, int64_t out_cstride,
                                       snb_conv_scatter** out) {
  if (!out) return fail(SNB_E_INVALID, "snb_conv_scatter_create: null output");
  *out = nullptr;
  if (!d_in || !d_weight || !d_bias || !d_out) return fail(SNB_E_INVALID, "null tensor pointer");
  if (n <= 0 || h <= 0 || w <= 0) return fail(SNB_E_INVALID, "bad input shape");
  if (cin <= 0 || cin % 32 != 0) return fail(SNB_E_INVALID, "cin=%lld must be a positive multiple of 32", (long long)cin);
  if (in_cstride < cin || in_cstride % 8 != 0 || out_cstride < kScCout || out_cstride % 8 != 0)
    return fail(SNB_E_INVALID, "bad channel strides");
  if ((d_pre_scale == nullptr) != (d_pre_shift == nullptr)) return fail(SNB_E_INVALID, "pre_scale and pre_shift come together");
  if ((reinterpret_cast<uintptr_t>(d_in) & 15) || (reinterpret_cast<uintptr_t>(d_out) & 31) ||
      (reinterpret_cast<uintptr_t>(d_weight) & 15) || (reinterpret_cast<uintptr_t>(d_bias) & 15) ||
      (reinterpret_cast<uintptr_t>(d_pre_scale) & 15) || (reinterpret_cast<uintptr_t>(d_pre_shift) & 15) ||
      (out_cstride * 2) % 32 != 0)
    return fail(SNB_E_INVALID, "pointers must be 16-byte aligned (the 16-channel output slot 32-byte aligned)");
  const int sms = sm_count();
  if (sms <= 0) return fail(SNB_E_CUDA, "no CUDA device");
  const bool pre = d_pre_scale != nullptr;
  const int bk = cin % 64 == 0 ? 64 : 32;
  const int swz = bk * 2;
  const int a_bytes = kScM * swz, b_bytes = kScN * swz;

  snb_conv_scatter* c = new (std::nothrow) snb_conv_scatter();
  if (!c) return fail(SNB_E_INVALID, "out of host memory");
  ScatterParams& p = c->params;
  std::memset(&p, 0, sizeof(p));
  p.k_chunks = static_cast<int32_t>(cin / bk);
  p.tiles_x = static_cast<int32_t>((w + kScOW - 1) / kScOW);
  p.tiles_y = static_cast<int32_t>((h + kScOH - 1) / kScOH);
  p.n_img = static_cast<int32_t>(n);
  const int64_t total = (int64_t)p.n_img * p.tiles_y * p.tiles_x;
  if (total > INT32_MAX) { delete c; return fail(SNB_E_UNSUPPORTED, "too many tiles"); }
  p.total_tiles = static_cast<int32_t>(total);
  p.w = static_cast<int32_t>(w);
  p.h = static_cast<int32_t>(h);
  p.bias = d_bias;
  p.pre_scale = d_pre_scale;
  p.pre_shift = d_pre_shift;
  p.out = static_cast<__nv_bfloat16*>(d_out);
  p.out_cstride = out_cstride;
  p.src = static_cast<const __nv_bfloat16*>(d_in);
  p.src_cstride = in_cstride;

  // TS variant (fused pre-activation only): activations go global -> registers -> tensor memory, shared memory holds the
  // weights (resident when all K chunks fit, else a ring) and the partial planes.  SNB_SCATTER_TS=0 selects the older kernel.
  bool ts = pre;
  if (const char* e = std::getenv("SNB_SCATTER_TS")) ts = ts && std::atoi(e) != 0;
  if (ts) {
    const int pre_bytes = (int)(((2 * cin * 4) + 127) & ~127);
    const int fixed_ts = 1024 /*align*/ + 1024 /*ctrl*/ + kScPBytes + pre_bytes;
    const int64_t w_all = (int64_t)p.k_chunks * b_bytes;
    // raw activation stages: an even count (the two prologue groups take alternate chunks, so a stage then always belongs to
    // the same group), at least 2; the weights are resident when they fit next to them
    const bool res = fixed_ts + 2 * a_bytes + w_all <= kScSmemBudget;
    int a_st, slots;
    if (res) {
      slots = p.k_chunks;
      a_st = (int)std::min<int64_t>(kScMaxStages, (kScSmemBudget - fixed_ts - w_all) / a_bytes) & ~1;
    } else {
      a_st = 2;
      slots = std::min<int>(kScMaxStages, (kScSmemBudget - fixed_ts - a_st * a_bytes) / b_bytes);
      if (slots >= 6 && fixed_ts + 4 * a_bytes + (slots - 2) * b_bytes <= kScSmemBudget) { a_st = 4; slots -= 2; }
    }
    if (a_st >= 2 && (slots >= 2 || res)) {
      p.a_stages = a_st;
      p.b_stages = res ? 0 : slots;
      p.bres = res ? 1 : 0;
      c->smem = fixed_ts + a_st * a_bytes + slots * b_bytes;
      {
        uint64_t dims[4] = {(uint64_t)cin, (uint64_t)w, (uint64_t)h, (uint64_t)n};
        uint64_t str[3] = {(uint64_t)in_cstride * 2, (uint64_t)w * in_cstride * 2, (uint64_t)h * w * in_cstride * 2};
        uint32_t box[4] = {(uint32_t)bk, (uint32_t)kScPW, (uint32_t)kScPH, 1};
        int rc = encode_map(&p.map_a, const_cast<void*>(d_in), 4, dims, str, box, swz, 2);
        if (rc) { delete c; return rc; }
      }
      uint64_t dims[3] = {(uint64_t)cin, (uint64_t)kScN, 1};
      uint64_t str[2] = {(uint64_t)cin * 2, (uint64_t)kScN * cin * 2};
      uint32_t box[3] = {(uint32_t)bk, (uint32_t)kScN, 1};
      int rc = encode_map(&p.map_b, const_cast<void*>(d_weight), 3, dims, str, box, swz, 2);
      if (rc) { delete c; return rc; }
      c->fn = bk == 64 ? (const void*)&conv_scatter_ts_kernel<64> : (const void*)&conv_scatter_ts_kernel<32>;
      c->threads = 512;
      cudaError_t e = cudaFuncSetAttribute(c->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kScSmemBudget);
      if (e != cudaSuccess) {
        delete c;
        return fail(SNB_E_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      }
      c->grid = std::min<int>(p.total_tiles, sms);
      c->flops = 2.0 * (double)n * h * w * (double)cin * kScCout * 9.0;
      *out = c;
      return SNB_OK;
    }
  }

  // pipeline shape: resident weights when every K chunk fits next to >= 2 activation stages
  const int fixed = 1024 /*align*/ + 1024 /*ctrl*/ + kScPBytes;
  const int64_t w_bytes = (int64_t)p.k_chunks * b_bytes;
  int a_stages, b_stages = 0;
  bool bres = fixed + 2 * a_bytes + w_bytes <= kScSmemBudget;
  if (const char* e = std::getenv("SNB_SCATTER_BRES")) bres = bres && std::atoi(e) != 0;
  if (bres) {
    a_stages = static_cast<int>(std::min<int64_t>(kScMaxStages, (kScSmemBudget - fixed - w_bytes) / a_bytes));
    c->smem = fixed + a_stages * a_bytes + (int)w_bytes;
  } else {
    // one activation stage per weight stage
    a_stages = b_stages = std::min(kScMaxStages, (kScSmemBudget - fixed) / (a_bytes + b_bytes));
    if (a_stages < 2) { delete c; return fail(SNB_E_UNSUPPORTED, "scatter pipeline does not fit in shared memory"); }
    c->smem = fixed + a_stages * (a_bytes + b_bytes);
  }
  p.a_stages = a_stages;
  p.b_stages = b_stages;
  p.bres = bres ? 1 : 0;

  int rc;
  {
    uint64_t dims[4] = {(uint64_t)cin, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    uint64_t str[3] = {(uint64_t)in_cstride * 2, (uint64_t)w * in_cstride * 2, (uint64_t)h * w * in_cstride * 2};
    uint32_t box[4] = {(uint32_t)bk, (uint32_t)kScPW, (uint32_t)kScPH, 1};
    rc = encode_map(&p.map_a, const_cast<void*>(d_in), 4, dims, str, box, swz, 2);
    if (rc) { delete c; return rc; }
  }
  {
    uint64_t dims[3] = {(uint64_t)cin, (uint64_t)kScN, 1};
    uint64_t str[2] = {(uint64_t)cin * 2, (uint64_t)kScN * cin * 2};
    uint32_t box[3] = {(uint32_t)bk, (uint32_t)kScN, 1};
    rc = encode_map(&p.map_b, const_cast<void*>(d_weight), 3, dims, str, box, swz, 2);
    if (rc) { delete c; return rc; }
  }
  if (bk == 64) c->fn = pre ? (const void*)&conv_scatter_kernel<64, true> : (const void*)&conv_scatter_kernel<64, false>;
  else c->fn = pre ? (const void*)&conv_scatter_kernel<32, true> : (const void*)&conv_scatter_kernel<32, false>;
  c->threads = pre ? 512 : 256;
  cudaError_t e = cudaFuncSetAttribute(c->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kScSmemBudget);
  if (e != cudaSuccess) {
    delete c;
    return fail(SNB_E_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
  }
  c->grid = std::min<int>(p.total_tiles, sms);
  c->flops = 2.0 * (double)n * h * w * (double)cin * kScCout * 9.0;
  *out = c;
  return SNB_OK;
}

extern "C" int snb_conv_scatter_launch(const snb_conv_scatter* c, void* stream) {
  if (!c) return fail(SNB_E_INVALID, "snb_conv_scatter_launch: null handle");
  void* args[1] = {const_cast<ScatterParams*>(&c->params)};
  cudaError_t e = cudaLaunchKernel(c->fn, dim3(c->grid), dim3(c->threads), args, c->smem, as_stream(stream));
  if (e != cudaSuccess) return fail(SNB_E_CUDA, "conv scatter launch failed: %s", cudaGetErrorString(e));
  return SNB_OK;
}

extern "C" void snb_conv_scatter_destroy(snb_conv_scatter* c) { delete c; }

extern "C" double snb_conv_scatter_flops(const snb_conv_scatter* c) { return c ? c->flops : 0.0; }

#ifdef SNB_CONV_PROFILE
extern "C" __attribute__((visibility("default"))) int snb_debug_scatter_profile(unsigned long long* out8, int reset) {
  if (out8 && cudaMemcpyFromSymbol(out8, snb::g_scatter_prof, sizeof(unsigned long long) * 12) != cudaSuccess) return 1;   // 12 values
  if (reset) {
    unsigned long long z[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpyToSymbol(snb::g_scatter_prof, z, sizeof(z)) != cudaSuccess) return 1;
  }
  return 0;
}
#endif
