// Tap-list implicit-GEMM convolution on the sm_100a tensor cores (tcgen05 + TMEM), fed by TMA.
//
// One kernel family covers what UNet16 / UNet11 execute (lib/models/unet16.py:8-49,113-131):
//   conv3x3 s1 p1      = 1 phase  x 9 taps  (dy,dx in -1..1)
//   conv1x1            = 1 phase  x 1 tap
//   ConvTranspose k4s2 = 4 phases x 4 taps  (sub-pixel decomposition; phase (py,px) writes out[2y+py][2x+px])
// GEMM view per phase:  D[pixel][cout] = sum_{tap,cin} A[pixel + (dy,dx)][cin] * W[tap][cout][cin]
//   M tile = 128 output pixels of one image, N tile = BN output channels, K step = one tap x BK input channels
//   (BK*2 bytes = one swizzle span).  TMA boxes over the NHWC slab give the conv zero padding and ragged edges
//   for free through out-of-bounds zero fill; no im2col buffer exists.
//
// Two main-loop variants share the epilogue:
//   conv_igemm_kernel (v1, "tap" mode): 16x8 pixel patch; every K step loads its own shifted A box + B box.
//   conv_halo_kernel  (v2, "halo" mode): 8-wide x 16-tall patch; ONE (8+2)x(16+2) halo box per input-channel chunk
//     feeds all taps: tap (dy,dx) is just a different start row of the UMMA descriptor (rows are 10 halo pixels
//     apart: SBO = 10 rows), so the activation traffic L2->smem drops ~6x; weights either stream through their
//     own ring or, when the whole [phase][tap][Cout][Cin] slab fits, stay resident in shared memory ("bres").
//
// Persistent CTAs (<= 1 per SM), warp-specialised:
//   warp 0  : TMA producer for A (and for the resident weights)      warp 3: TMA producer for B (v2)
//   warp 1  : MMA issuer (one thread: tcgen05.mma kind::f16, fp32 accumulators in TMEM, 2 accumulator stages)
//   warp 2  : TMEM allocator
//   warps 4-7: epilogue (tcgen05.ld -> bias/ReLU -> bf16 -> swizzled smem -> TMA store into the channel slab at
//              its concat offset; or the fused 1x1 head + sigmoid of the last layer)
// so the epilogue of tile i overlaps the main loop of tile i+1.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>

#include "sm100_ptx.cuh"
#include "snb_internal.h"

namespace snb {

constexpr int kBM = 128;  // UMMA M = pixels per tile
constexpr int kMaxPhases = 4;
constexpr int kMaxTaps = 9;
constexpr int kSmemBudget = 227 * 1024;
constexpr int kHaloW = 10, kHaloH = 18;           // halo box of the 8 x 16 patch
constexpr int kHaloRows = kHaloW * kHaloH;        // 180 smem rows
constexpr int kMaxAStages = 8, kMaxBStages = 8;

struct alignas(64) ConvParams {
  CUtensorMap map_a;               // activations (C, W, H, N)
  CUtensorMap map_b;               // weights (Cin, Cout, phase*taps), box (BK, BN, 1)
  CUtensorMap map_d[kMaxPhases];   // outputs per phase (C, W, H, N)
  CUtensorMap map_p;               // fused 2x2 max-pool output (C, W/2, H/2, N), box (CW, TW/2, TH/2, 1)
  CUtensorMap map_bh;              // weights with a half-height box (BK, BN/2, 1): multicast halves in CTA pairs
  int32_t n_phases, taps;          // taps per phase
  int32_t k_chunks;                // Cin / BK
  int32_t n_tiles;                 // Cout / BN
  int32_t tiles_x, tiles_y, n_img;
  int32_t total_tiles;
  int32_t relu;
  int32_t head_sigmoid;
  int32_t out_w, out_h;            // head output bounds
  int32_t a_stages, b_stages, bres;  // v2 pipeline shape
  int32_t dual_mma;                  // two MMA-issuing warps: 1 = alternating tiles, 2 = phases 0-1 / 2-3 of every tile
  int32_t epi_groups;                // 2 = warps 8-11 are a second epilogue group (kernels launched with 384 threads)
  int32_t out_bufs;                  // 1 = single output staging buffer (the room goes to the weight ring); else 2
  int32_t pool;                      // also write the 2x2 max-pooled tile through map_p
  int32_t m_pairs;                   // CTA-pair kernels: spatial tiles per phase / 2
  int32_t up2x;                      // store every tile through all 4 map_d (nearest 2x upsample of the output)
  float head_b;
  const float* bias;
  const float* head_w;
  float* head_out;
  const void* residual;            // optional tensor added in the epilogue (same grid as the output), or null
  int32_t res_cstride;             // its pixel stride in channels
  int32_t res_after_act;           // 0: act(conv + bias + res) (ResNet block); 1: act(conv + bias) + res (LinkNet skip)
  int32_t load_dx;                 // offset added to the halo-box origin (+1 for a 'valid' conv3x3: the tile grid is
                                   // the output grid and tap (dy,dx) reads input pixel (x+1+dx, y+1+dy))
  int32_t in_w, in_h;              // input extent (pre-activation prologue masks with it)
  float act_slope;                 // leaky-ReLU slope of the activation (0 = ReLU); only used when relu != 0
  int32_t ext;                     // residual operand or leaky slope present: run the EXT epilogue body
  const float* pre_scale;          // optional pre-activation y = relu(x * scale[c] + shift[c]) applied to the A operand
  const float* pre_shift;
  int8_t tap_dy[kMaxPhases][kMaxTaps];
  int8_t tap_dx[kMaxPhases][kMaxTaps];
};

// EB = element bytes of activations / weights: 2 = bf16, 4 = fp32 storage consumed as TF32.  BK counts elements, so a
// K chunk is always one swizzle span of SWZ = BK * EB bytes (128 or 64) and one MMA consumes 32 bytes of it.
template <int BN, int BK, int EB = 2>
struct ConvCfg {
  static constexpr int SWZ = BK * EB;                   // bytes per operand row = swizzle span
  static constexpr int A_BYTES = kBM * SWZ;
  static constexpr int B_BYTES = BN * SWZ;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int CW = BN < 128 / EB ? BN : 128 / EB;   // channels per output store chunk (<= 128 bytes)
  static constexpr int OUT_SWZ = CW * EB;
  static constexpr int OUT_BYTES = kBM * OUT_SWZ;
  static constexpr int CTRL_BYTES = 1024;               // barriers + tmem pointer
  static constexpr int RAW_STAGES = (kSmemBudget - 1024 /*align slack*/ - 2 * OUT_BYTES - CTRL_BYTES) / STAGE_BYTES;
  static constexpr int NSTAGES = RAW_STAGES > 8 ? 8 : RAW_STAGES;
  static constexpr int SMEM_BYTES = 1024 + NSTAGES * STAGE_BYTES + 2 * OUT_BYTES + CTRL_BYTES;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;  // two accumulator stages (power of two)
  // v2
  static constexpr int HALO_BOX_BYTES = kHaloRows * SWZ;
  static constexpr int HALO_STAGE_BYTES = (HALO_BOX_BYTES + 1023) / 1024 * 1024;
  static constexpr int POOL_BYTES = (kBM / 4) * OUT_SWZ;   // staging of the pooled tile (32 pixels)
  static_assert(NSTAGES >= 3, "pipeline too shallow");
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0 && OUT_BYTES % 1024 == 0, "swizzle atom alignment");
  static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS <= 512, "TMEM columns");
};

struct TileCoord {
  int nt, x0, y0, img, ph;
};

template <int TW, int TH>
__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int t) {
  TileCoord c;
  c.nt = t % p.n_tiles;
  t /= p.n_tiles;
  c.x0 = (t % p.tiles_x) * TW;
  t /= p.tiles_x;
  c.y0 = (t % p.tiles_y) * TH;
  t /= p.tiles_y;
  c.img = t % p.n_img;
  c.ph = t / p.n_img;
  return c;
}

// CTA-pair order: consecutive tile ids (2j, 2j+1) are two spatial tiles with the same N tile and phase, so the two
// CTAs of a cluster walk identical (chunk, tap) sequences and can share every weight box.
template <int TW, int TH>
__device__ __forceinline__ TileCoord decode_tile_pair(const ConvParams& p, int u) {
  TileCoord c;
  const int r = u & 1;
  int v = u >> 1;
  c.nt = v % p.n_tiles;
  v /= p.n_tiles;
  int m = (v % p.m_pairs) * 2 + r;
  c.ph = v / p.m_pairs;
  c.x0 = (m % p.tiles_x) * TW;
  m /= p.tiles_x;
  c.y0 = (m % p.tiles_y) * TH;
  c.img = m / p.tiles_y;
  return c;
}

// bf16x2 pack with the ReLU folded into the conversion (negative -> +0): one instruction instead of two FMNMX + F2F
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// two fp32 adds in one instruction (sm_100 packed fp32 pipe)
__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tadd.rn.f32x2 ra, ra, rb;\n\t"
      "mov.b64 {%0, %1}, ra;\n\t}"
      : "+f"(a0), "+f"(a1) : "f"(b0), "f"(b1));
}

__device__ __forceinline__ uint4 shfl_xor_u4(uint4 v, int m) {
  return make_uint4(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m),
                    __shfl_xor_sync(0xffffffffu, v.z, m), __shfl_xor_sync(0xffffffffu, v.w, m));
}

__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

__device__ __forceinline__ uint4 bf16x8_max(uint4 a, uint4 b) {
  return make_uint4(bf16x2_max(a.x, b.x), bf16x2_max(a.y, b.y), bf16x2_max(a.z, b.z), bf16x2_max(a.w, b.w));
}

__device__ __forceinline__ uint4 f32x4_max(uint4 a, uint4 b) {
  return make_uint4(__float_as_uint(fmaxf(__uint_as_float(a.x), __uint_as_float(b.x))),
                    __float_as_uint(fmaxf(__uint_as_float(a.y), __uint_as_float(b.y))),
                    __float_as_uint(fmaxf(__uint_as_float(a.z), __uint_as_float(b.z))),
                    __float_as_uint(fmaxf(__uint_as_float(a.w), __uint_as_float(b.w))));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// Epilogue of one accumulator tile, executed by the 4 epilogue warps (128 threads = 128 TMEM lanes = 128 pixels).
// TW = patch width in pixels (row r of the tile is pixel (r % TW, r / TW)).
// EXT = the ResNet / LinkNet extras (residual operand, leaky-ReLU slope); the VGG / UNet layers run the lean EXT = false
// body: the high-resolution layers (K = 32..576 per 128 x 64 outputs) are bound by this epilogue, where the extra
// selects and predicated adds cost 10-35 %.
// EG = 2: two epilogue groups of 4 warps take alternate tiles (group = TMEM stage); each group then owns ONE staging
// buffer (index grp) and its own named barrier instead of double-buffering inside one group.
template <int BN, bool HEAD, int TW, int EB, bool EXT, int EG = 1>
__device__ __forceinline__ void epilogue_tile(const ConvParams& p, const TileCoord& tc, uint32_t t_addr,
                                              uint64_t* tmem_empty_bar, uint8_t* smem_out, uint32_t& n_store,
                                              int row, int lane, int epi_tid, bool release = true, int grp = 0) {
  constexpr int CW = BN < 128 / EB ? BN : 128 / EB;
  constexpr int OUT_SWZ = CW * EB;
  constexpr int OUT_BYTES = kBM * OUT_SWZ;
  if constexpr (HEAD) {
    static_assert(!HEAD || BN == 32, "fused head needs all channels of a pixel in one thread");
    uint32_t v[32];
    tmem_ld_32x32(t_addr, v);
    tmem_ld_wait();
    tc05_fence_before();
    __syncwarp();
    if (release && lane == 0) mbar_arrive(tmem_empty_bar);
    float dot = p.head_b;
    const float4* b4 = reinterpret_cast<const float4*>(p.bias);
    const float4* w4 = reinterpret_cast<const float4*>(p.head_w);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 b = __ldg(b4 + i), w = __ldg(w4 + i);
      float x0 = __uint_as_float(v[4 * i + 0]) + b.x, x1 = __uint_as_float(v[4 * i + 1]) + b.y;
      float x2 = __uint_as_float(v[4 * i + 2]) + b.z, x3 = __uint_as_float(v[4 * i + 3]) + b.w;
      if (p.relu) {
        if constexpr (EXT) {
          x0 = x0 > 0.f ? x0 : x0 * p.act_slope; x1 = x1 > 0.f ? x1 : x1 * p.act_slope;
          x2 = x2 > 0.f ? x2 : x2 * p.act_slope; x3 = x3 > 0.f ? x3 : x3 * p.act_slope;
        } else {
          x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f);
        }
      }
      dot = fmaf(x0, w.x, dot); dot = fmaf(x1, w.y, dot); dot = fmaf(x2, w.z, dot); dot = fmaf(x3, w.w, dot);
    }
    if (p.head_sigmoid) dot = 1.f / (1.f + expf(-dot));
    const int ox = tc.x0 + (row % TW);
    const int oy = tc.y0 + (row / TW);
    if (ox < p.out_w && oy < p.out_h)
      p.head_out[(static_cast<int64_t>(tc.img) * p.out_h + oy) * p.out_w + ox] = dot;
  } else {
    constexpr int NCHUNK = BN / CW;
    constexpr int POOL_BYTES = (kBM / 4) * OUT_SWZ;
    // 2x2 max-pool partners of pixel (x, y) sit 1 and TW lanes away; lanes with even x and even y keep the result
    const bool pool_keep = ((lane & 1) | (lane & TW)) == 0;
    const int prow = ((row / TW) >> 1) * (TW / 2) + ((row % TW) >> 1);
    // residual operand: same pixel grid as the stored output (only used by single-phase convs)
    const int rx = tc.x0 + (row % TW), ry = tc.y0 + (row / TW);
    const bool res_ok = rx < p.out_w && ry < p.out_h;
    const int64_t res_pix = ((static_cast<int64_t>(tc.img) * p.out_h + ry) * p.out_w + rx) * p.res_cstride;
#pragma unroll 1
    for (int c = 0; c < NCHUNK; ++c, ++n_store) {
      const bool two = EG == 2 && p.epi_groups == 2;
      const bool single = !two && p.out_bufs == 1;
      const uint32_t buf = two ? static_cast<uint32_t>(grp) : (single ? 0u : (n_store & 1));
      uint8_t* sout = smem_out + buf * OUT_BYTES;
      uint8_t* spool = smem_out + (single ? 1 : 2) * OUT_BYTES + buf * POOL_BYTES;   // (only allocated when p.pool)
      // the TMA store that last read this buffer (two chunks ago; the previous chunk with one buffer per group) must be done
      if (epi_tid == 0) {
        if (two || single) tma_store_wait_read<0>();
        else tma_store_wait_read<1>();
      }
      named_bar_sync(1 + grp, 128);
#pragma unroll
      for (int g = 0; g < CW / 32; ++g) {
        uint32_t v[32];
        tmem_ld_32x32(t_addr + c * CW + g * 32, v);
        tmem_ld_wait();
        const float4* bias4 = reinterpret_cast<const float4*>(p.bias + tc.nt * BN + c * CW + g * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // 4 x 16-byte chunks of 8 channels
          const float4 b0 = __ldg(bias4 + 2 * j), b1 = __ldg(bias4 + 2 * j + 1);
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j * 8 + e]);
          add2(f[0], f[1], b0.x, b0.y);
          add2(f[2], f[3], b0.z, b0.w);
          add2(f[4], f[5], b1.x, b1.y);
          add2(f[6], f[7], b1.z, b1.w);
          if constexpr (EXT) {
            float res[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (p.residual != nullptr && res_ok) {
              const int64_t ro = res_pix + tc.nt * BN + c * CW + g * 32 + j * 8;
              if constexpr (EB == 2) {
                const uint4 rv = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.residual) + ro));
                const __nv_bfloat162* pr = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 t2 = __bfloat1622float2(pr[e]);
                  res[2 * e] = t2.x;
                  res[2 * e + 1] = t2.y;
                }
              } else {
                const float4 r0 = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(p.residual) + ro));
                const float4 r1 = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(p.residual) + ro) + 1);
                res[0] = r0.x; res[1] = r0.y; res[2] = r0.z; res[3] = r0.w;
                res[4] = r1.x; res[5] = r1.y; res[6] = r1.z; res[7] = r1.w;
              }
            }
            if (!p.res_after_act) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] += res[e];
            }
            if (p.relu) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = f[e] > 0.f ? f[e] : f[e] * p.act_slope;
            }
            if (p.res_after_act) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] += res[e];
            }
          } else {
            if (EB != 2 && p.relu) {   // bf16 storage folds the ReLU into the conversion below
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
            }
          }
          const int sw = OUT_SWZ == 128 ? (row & 7) : ((row >> 1) & 3);
          const int swp = OUT_SWZ == 128 ? (prow & 7) : ((prow >> 1) & 3);
          if constexpr (EB == 2) {
            const int chunk = g * 4 + j;
            uint4* dst = reinterpret_cast<uint4*>(sout + row * OUT_SWZ + ((chunk ^ sw) << 4));
            uint4 pk;
            if (!EXT && p.relu)
              pk = make_uint4(pack_bf16x2_relu(f[0], f[1]), pack_bf16x2_relu(f[2], f[3]), pack_bf16x2_relu(f[4], f[5]),
                              pack_bf16x2_relu(f[6], f[7]));
            else
              pk = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                              pack_bf16x2(f[6], f[7]));
            *dst = pk;   // (the 2x2 max-pool of a bf16 tile is taken from this staging buffer after the barrier)
          } else {
            // fp32 storage: 8 channels = two 16-byte chunks; values are rounded to TF32 so that the next layer's
            // tensor-core read (which drops the low mantissa bits) sees exactly what is stored
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int chunk = g * 8 + j * 2 + h;
              uint4 pk = make_uint4(__float_as_uint(round_tf32(f[4 * h + 0])), __float_as_uint(round_tf32(f[4 * h + 1])),
                                    __float_as_uint(round_tf32(f[4 * h + 2])), __float_as_uint(round_tf32(f[4 * h + 3])));
              *reinterpret_cast<uint4*>(sout + row * OUT_SWZ + ((chunk ^ sw) << 4)) = pk;
              if (p.pool) {
                pk = f32x4_max(pk, shfl_xor_u4(pk, 1));
                pk = f32x4_max(pk, shfl_xor_u4(pk, TW));
                if (pool_keep) *reinterpret_cast<uint4*>(spool + prow * OUT_SWZ + ((chunk ^ swp) << 4)) = pk;
              }
            }
          }
        }
      }
      if (release && c == NCHUNK - 1) {
        // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp
        tc05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty_bar);
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + grp, 128);
      if (epi_tid == 0) {
        if (p.up2x) {
#pragma unroll
          for (int rep = 0; rep < 4; ++rep)
            tma_store_4d(&p.map_d[rep], sout, tc.nt * BN + c * CW, tc.x0, tc.y0, tc.img);
        } else {
          tma_store_4d(&p.map_d[tc.ph], sout, tc.nt * BN + c * CW, tc.x0, tc.y0, tc.img);
        }
        if (EB != 2 && p.pool) tma_store_4d(&p.map_p, spool, tc.nt * BN + c * CW, tc.x0 >> 1, tc.y0 >> 1, tc.img);
        if (!(EB == 2 && p.pool)) tma_store_commit();   // (with the staged pool below: one group per chunk, closed there)
      }
      if constexpr (EB == 2) {
        if (p.pool) {
          // fused MaxPool2d(2, 2) of a bf16 tile, from the staged rows: one (pooled pixel, 16-byte chunk) item per thread
          // and pass = 4 LDS.128 + 12 HMNMX2 + 1 STS.128, against 8 shuffles + 8 HMNMX2 per chunk in registers
          constexpr int NCH = OUT_SWZ / 16;
#pragma unroll
          for (int it = 0; it < NCH / 4; ++it) {
            const int item = epi_tid + it * 128;
            const int pr = item / NCH, ch = item % NCH;
            const int r00 = (pr / (TW / 2)) * (2 * TW) + (pr % (TW / 2)) * 2;
            auto at = [&](int r) {
              const int swr = OUT_SWZ == 128 ? (r & 7) : ((r >> 1) & 3);
              return *reinterpret_cast<const uint4*>(sout + r * OUT_SWZ + ((ch ^ swr) << 4));
            };
            const uint4 m = bf16x8_max(bf16x8_max(at(r00), at(r00 + 1)), bf16x8_max(at(r00 + TW), at(r00 + TW + 1)));
            const int swq = OUT_SWZ == 128 ? (pr & 7) : ((pr >> 1) & 3);
            *reinterpret_cast<uint4*>(spool + pr * OUT_SWZ + ((ch ^ swq) << 4)) = m;
          }
          fence_proxy_async_smem();
          named_bar_sync(1 + grp, 128);
          if (epi_tid == 0) {
            tma_store_4d(&p.map_p, spool, tc.nt * BN + c * CW, tc.x0 >> 1, tc.y0 >> 1, tc.img);
            tma_store_commit();
          }
        }
      }
    }
  }
}

// =============================================================================================== v1: tap mode
template <int BN, int BK, bool HEAD, int EB>
__global__ void __launch_bounds__(256, 1) conv_igemm_kernel(const __grid_constant__ ConvParams p) {
  using Cfg = ConvCfg<BN, BK, EB>;
  constexpr int NSTAGES = Cfg::NSTAGES;
  constexpr uint32_t IDESC = make_idesc(kBM, BN, EB == 2 ? 1u : 2u);
  constexpr int TW = 16, TH = 8;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_stage = smem;                                  // NSTAGES x [A | B]
  uint8_t* smem_out = smem + NSTAGES * Cfg::STAGE_BYTES;       // 2 x OUT_BYTES
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_out + 2 * Cfg::OUT_BYTES);
  uint64_t* full_bar = bars;                  // [NSTAGES]
  uint64_t* empty_bar = bars + NSTAGES;       // [NSTAGES]
  uint64_t* tmem_full = bars + 2 * NSTAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_a);
    tma_prefetch_desc(&p.map_b);
    if (!HEAD)
      for (int i = 0; i < (p.up2x ? 4 : p.n_phases); ++i) tma_prefetch_desc(&p.map_d[i]);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NSTAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc05_fence_before();
  __syncthreads();
  tc05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int k_steps = p.taps * p.k_chunks;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile<TW, TH>(p, t);
        for (int tap = 0; tap < p.taps; ++tap) {
          const int ax = tc.x0 + p.tap_dx[tc.ph][tap];
          const int ay = tc.y0 + p.tap_dy[tc.ph][tap];
          const int wtap = tc.ph * p.taps + tap;
          for (int kc = 0; kc < p.k_chunks; ++kc, ++it) {
            const uint32_t s = it % NSTAGES;
            const uint32_t ph = (it / NSTAGES) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* sa = smem_stage + s * Cfg::STAGE_BYTES;
            uint8_t* sb = sa + Cfg::A_BYTES;
            mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
            tma_load_4d(&p.map_a, &full_bar[s], sa, kc * BK, ax, ay, tc.img);
            tma_load_3d(&p.map_b, &full_bar[s], sb, kc * BK, tc.nt * BN, wtap);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (single thread)
    if (elect_one()) {
      uint32_t it = 0;
      uint32_t local_tile = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++local_tile) {
        const uint32_t acc = local_tile & 1;
        const uint32_t acc_ph = (local_tile >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_ph ^ 1);
        tc05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int ks = 0; ks < k_steps; ++ks, ++it) {
          const uint32_t s = it % NSTAGES;
          const uint32_t ph = (it / NSTAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc05_fence_after();
          const uint32_t a_addr = smem_u32(smem_stage + s * Cfg::STAGE_BYTES);
          const uint64_t adesc = make_kmajor_desc<Cfg::SWZ>(a_addr, 8 * Cfg::SWZ);
          const uint64_t bdesc = make_kmajor_desc<Cfg::SWZ>(a_addr + Cfg::A_BYTES, 8 * Cfg::SWZ);
#pragma unroll
          for (int k = 0; k < Cfg::SWZ / 32; ++k) {
            // advance 32 bytes (16 bf16 / 8 tf32) along K inside the swizzle span: +2 in the (addr >> 4) field
            umma_ss<EB>(adesc + 2 * k, bdesc + 2 * k, d_tmem, IDESC, (ks > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);  // smem slot reusable once these MMAs have read it
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (128 threads)
    const int q = warp & 3;              // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;       // pixel index inside the patch
    const int epi_tid = threadIdx.x - 128;
    uint32_t local_tile = 0;
    uint32_t n_store = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++local_tile) {
      const TileCoord tc = decode_tile<TW, TH>(p, t);
      const uint32_t acc = local_tile & 1;
      const uint32_t acc_ph = (local_tile >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_ph);
      tc05_fence_after();
      const uint32_t t_addr = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
      if (p.ext) epilogue_tile<BN, HEAD, TW, EB, true>(p, tc, t_addr, &tmem_empty[acc], smem_out, n_store, row, lane, epi_tid);
      else epilogue_tile<BN, HEAD, TW, EB, false>(p, tc, t_addr, &tmem_empty[acc], smem_out, n_store, row, lane, epi_tid);
    }
    if (!HEAD && epi_tid == 0) tma_store_wait_all<0>();
  }

  tc05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc05_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ============================================================================================== v2: halo mode
// smem: [a_stages x HALO_STAGE][B: b_stages x B_BYTES, or all (phase, chunk, tap) weight boxes when bres][2 x OUT][ctrl]
//
// The MMA-issuing thread is the critical resource: tcgen05.mma M=128 x K=16 retires in max(N/2, ~40) cycles
// (measured, tools/micro/mma_rate.cu), so one thread has to issue an MMA every 64 cycles at N=128.  Everything it
// needs is therefore precomputed: tap -> descriptor offsets live in a shared-memory table, stage indices advance by
// compare-and-wrap (no division), descriptors are built from 32-bit halves with immediate K offsets, and the tap
// loop is fully unrolled (TAPS is a template parameter: 9 for conv3x3, 4 for one ConvTranspose phase).
// -DSNB_CONV_PROFILE (tools/build_rev.py --profile; never in the shipped library): cycles the roles of conv_halo_kernel
// spend waiting, summed over all CTAs.  [0] issuer: activation stage, [1] issuer: weight slot, [2] issuer: accumulator
// stage, [3] issuer: whole loop, [4] epilogue warp 4: accumulator full, [5] epilogue warp 4: whole loop, [6] weight
// producer: slot free, [7] tiles.
#ifdef SNB_CONV_PROFILE
__device__ unsigned long long g_conv_prof[8];
#define SNB_PROF_DECL long long prof_t = 0; unsigned long long prof_c[4] = {0, 0, 0, 0}; (void)prof_t;
#define SNB_PROF_T0 prof_t = clock64();
#define SNB_PROF_ADD(i) prof_c[i] += static_cast<unsigned long long>(clock64() - prof_t);
#define SNB_PROF_FLUSH(dst, i) atomicAdd(&g_conv_prof[dst], prof_c[i]);
#else
#define SNB_PROF_DECL
#define SNB_PROF_T0
#define SNB_PROF_ADD(i)
#define SNB_PROF_FLUSH(dst, i)
#endif

__device__ __forceinline__ uint64_t desc_from_halves(uint32_t lo, uint32_t hi) {
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

template <int SWZ>
__device__ __forceinline__ constexpr uint32_t desc_hi(uint32_t sbo_bytes) {
  // bits [32,46) SBO>>4, [46,48) version = 1, [61,64) layout type
  return (sbo_bytes >> 4) | (1u << 14) | ((SWZ == 128 ? 2u : (SWZ == 64 ? 4u : 6u)) << 29);
}

// NPH = 4 fuses the four sub-pixel phases of a ConvTranspose into one tile (four accumulators of BN columns): the
// halo box is then loaded once per spatial tile instead of once per phase.
// CL = 2 runs CTA pairs (thread-block cluster of 2) on two spatial tiles of the same N tile: each CTA fetches half of
// every weight box and TMA-multicasts it into both CTAs' shared memory, halving the L2->SM weight traffic that
// bounds the large layers (32 KB per tap per CTA at BN=256 otherwise).
//
// PRE = true adds eight "prologue" warps (8-15) that rewrite every halo stage in place after TMA lands it and before
// the MMA reads it: y = relu(x * scale[c] + shift[c]) per input channel, zero outside the image (the convolution's
// zero padding applies AFTER the pre-activation).  This is FCDenseNet's per-consumer BatchNorm+ReLU
// (lib/models/tiramisu.py:12-13) fused into the operand path instead of a separate pass over the slab.
template <int BN, int BK, bool HEAD, int TAPS, int NPH, int CL, int EB, bool PRE = false>
__global__ void __launch_bounds__(PRE ? 512 : (CL == 1 && BN <= 64 ? 384 : 256), 1)
conv_halo_kernel(const __grid_constant__ ConvParams p) {
  // narrow tiles are bound by the epilogue's dependent-instruction latency (one warp per scheduler): two epilogue groups
  // (not the fused 1x1 head: its epilogue is a dot product per pixel, that layer is bound by the MMA pipe)
  constexpr int EG = (!PRE && !HEAD && CL == 1 && BN <= 64) ? 2 : 1;
  static_assert(!PRE || (EB == 2 && NPH == 1 && CL == 1), "the fused pre-activation is built for bf16 conv3x3");
  static_assert(CL == 1 || (CL == 2 && NPH == 1 && !HEAD && BN >= 128), "CTA pairs are for the streamed-weight layers");
  using Cfg = ConvCfg<BN, BK, EB>;
  static_assert(NPH == 1 || (NPH == 4 && TAPS == 4 && !HEAD), "phase fusion is for ConvTranspose");
  // accumulator stages: narrow tiles have a short main loop (M) next to a long epilogue (E); with two stages a tile costs
  // (M + E) / 2 however many warps share the work, with four the issuers run ahead and it costs max(M, E / groups)
  constexpr int NACC = (!PRE && !HEAD && CL == 1 && NPH * BN <= 128) ? 4 : 2;
  constexpr int TCOLS = NACC * NPH * BN < 32 ? 32 : NACC * NPH * BN;
  static_assert(TCOLS <= 512 && (TCOLS & (TCOLS - 1)) == 0, "TMEM columns");
  constexpr uint32_t IDESC = make_idesc(kBM, BN, EB == 2 ? 1u : 2u);
  constexpr int TW = 8, TH = 16;
  constexpr int SWZ = Cfg::SWZ;
  constexpr uint32_t A_STAGE16 = Cfg::HALO_STAGE_BYTES >> 4;
  constexpr uint32_t B_BYTES16 = Cfg::B_BYTES >> 4;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int n_b_slots = p.bres ? p.n_phases * TAPS * p.k_chunks : p.b_stages;  // n_phases: weight phases
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + p.a_stages * Cfg::HALO_STAGE_BYTES;
  uint8_t* smem_out = smem_b + n_b_slots * Cfg::B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(
      smem_out + (HEAD ? 0 : (p.out_bufs == 1 ? 1 : 2) * (Cfg::OUT_BYTES + (p.pool ? Cfg::POOL_BYTES : 0))));
  uint64_t* a_full = bars;                          // [kMaxAStages]
  uint64_t* a_empty = a_full + kMaxAStages;         // [kMaxAStages]
  uint64_t* b_full = a_empty + kMaxAStages;         // [kMaxBStages]  (b_full[0] doubles as the bres barrier)
  uint64_t* b_empty = b_full + kMaxBStages;         // [kMaxBStages]
  uint64_t* tmem_full = b_empty + kMaxBStages;      // [4]
  uint64_t* tmem_empty = tmem_full + 4;             // [4]
  uint64_t* a_ready = tmem_empty + 4;               // [kMaxAStages] stage rewritten by the prologue warps (PRE)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(a_ready + kMaxAStages);
  uint32_t* s_aoff = tmem_ptr + 4;                  // [kMaxPhases][TAPS] descriptor start offsets (>>4) per tap

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_a);
    tma_prefetch_desc(&p.map_b);
    if (!HEAD)
      for (int i = 0; i < (p.up2x ? 4 : p.n_phases); ++i) tma_prefetch_desc(&p.map_d[i]);
    if (p.pool) tma_prefetch_desc(&p.map_p);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kMaxAStages; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], p.dual_mma == 2 ? 2 : 1);   // phase-split issuers both read every activation stage
      mbar_init(&a_ready[i], 8);    // one arrival per prologue warp
    }
    for (int i = 0; i < kMaxBStages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], CL);   // every CTA of the cluster must have consumed a multicast stage
    }
    for (int i = 0; i < NACC; ++i) {
      mbar_init(&tmem_full[i], p.dual_mma == 2 ? 2 : 1);
      mbar_init(&tmem_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, TCOLS);
    tmem_relinquish();
  }
  if (warp == 3 && lane < p.n_phases * TAPS) {
    // tap (dy,dx): the 128 operand rows start (dy+1) halo rows down and (dx+1) pixels right
    const int ph = lane / TAPS, tap = lane % TAPS;
    s_aoff[lane] = static_cast<uint32_t>(((p.tap_dy[ph][tap] + 1) * kHaloW + (p.tap_dx[ph][tap] + 1)) * SWZ) >> 4;
  }
  tc05_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // peer barriers are initialised before anything is multicast into them
  tc05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  auto decode = [&](int t) { return CL > 1 ? decode_tile_pair<TW, TH>(p, t) : decode_tile<TW, TH>(p, t); };

  if (warp == 0) {
    // ------------------------------------------------------------------ A producer (+ resident weights)
    if (elect_one()) {
      if (p.bres) {
        mbar_arrive_expect_tx(&b_full[0], static_cast<uint32_t>(n_b_slots) * Cfg::B_BYTES);
        uint8_t* dst = smem_b;   // slot order = the order the MMA loop walks: (phase, chunk, tap), or
                                 // (chunk, phase, tap) when the phases are fused into one tile
        if (NPH == 1) {
          for (int ph = 0; ph < p.n_phases; ++ph)
            for (int kc = 0; kc < p.k_chunks; ++kc)
              for (int tap = 0; tap < TAPS; ++tap, dst += Cfg::B_BYTES)
                tma_load_3d(&p.map_b, &b_full[0], dst, kc * BK, 0, ph * TAPS + tap);
        } else {
          for (int kc = 0; kc < p.k_chunks; ++kc)
            for (int ph = 0; ph < NPH; ++ph)
              for (int tap = 0; tap < TAPS; ++tap, dst += Cfg::B_BYTES)
                tma_load_3d(&p.map_b, &b_full[0], dst, kc * BK, 0, ph * TAPS + tap);
        }
      }
      uint32_t s = 0, par = 1;   // waiting on parity 1 of a fresh barrier returns immediately
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord tc = decode(t);
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          mbar_wait(&a_empty[s], par);
          mbar_arrive_expect_tx(&a_full[s], Cfg::HALO_BOX_BYTES);
          tma_load_4d(&p.map_a, &a_full[s], smem_a + s * Cfg::HALO_STAGE_BYTES, kc * BK, tc.x0 - 1 + p.load_dx,
                      tc.y0 - 1 + p.load_dx, tc.img);
          if (++s == static_cast<uint32_t>(p.a_stages)) { s = 0; par ^= 1; }
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ B producer (streamed weights)
    if (!p.bres && elect_one()) {
      const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0;
      uint32_t s = 0, par = 1;
      SNB_PROF_DECL
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord tc = decode(t);
        const int n0 = tc.nt * BN, w0 = tc.ph * TAPS;   // tc.ph == 0 when the phases are fused
        for (int kc = 0; kc < p.k_chunks; ++kc) {
#pragma unroll 1
          for (int wt = 0; wt < NPH * TAPS; ++wt) {
            SNB_PROF_T0
            mbar_wait(&b_empty[s], par);
            SNB_PROF_ADD(0)
            mbar_arrive_expect_tx(&b_full[s], Cfg::B_BYTES);
            if (CL == 1) {
              tma_load_3d(&p.map_b, &b_full[s], smem_b + s * Cfg::B_BYTES, kc * BK, n0, w0 + wt);
            } else {
              // my half of the box (BN/2 weight rows) goes to both CTAs; the peer sends the other half
              tma_load_3d_mc(&p.map_bh, &b_full[s], smem_b + s * Cfg::B_BYTES + cta_rank * (Cfg::B_BYTES / 2), kc * BK,
                             n0 + static_cast<int>(cta_rank) * (BN / 2), w0 + wt, static_cast<uint16_t>(0x3));
            }
            if (++s == static_cast<uint32_t>(p.b_stages)) { s = 0; par ^= 1; }
          }
        }
      }
      SNB_PROF_FLUSH(6, 0)
    }
  } else if (warp == 1 || (warp == 2 && p.dual_mma)) {
    // ------------------------------------------------------------------ MMA issuer (one thread per issuing warp)
    // A narrow-N MMA retires in 40-48 cycles but costs the issuing thread ~60 (descriptor moves into uniform registers),
    // so those layers were issue-bound (ncu: the issuing warp busy 85 % of the time, epilogue warps waiting).  With
    // dual_mma = 1 a second warp issues too: warp 1 owns the even local tiles (TMEM stage 0), warp 2 the odd ones (stage 1);
    // both walk the same in-order operand rings and skip the other warp's stages.  dual_mma = 2 (fused ConvTranspose
    // phases, resident weights): both warps work on EVERY tile, warp 1 on the accumulators of phases 0-1, warp 2 on 2-3.
    if (elect_one()) {
      constexpr uint32_t HI_A = desc_hi<SWZ>(kHaloW * SWZ);   // 8-row groups are one halo row (10 pixels) apart
      constexpr uint32_t HI_B = desc_hi<SWZ>(8 * SWZ);
      const uint32_t a_lo0 = (smem_u32(smem_a) & 0x3FFFFu) >> 4;
      const uint32_t b_lo0 = (smem_u32(smem_b) & 0x3FFFFu) >> 4;
      const bool bres = p.bres != 0;
      const uint32_t a_stages = p.a_stages, b_stages = p.b_stages;
      const bool dual = p.dual_mma == 1;
      const bool by_phase = NPH == 4 && p.dual_mma == 2;
      const uint32_t mine = warp == 1 ? 0u : 1u;
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
      uint32_t local_tile = 0;
      if (bres) {
        mbar_wait(&b_full[0], 0);
        tc05_fence_after();
      }
      SNB_PROF_DECL
#ifdef SNB_CONV_PROFILE
      const long long prof_loop0 = clock64();
#endif
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++local_tile) {
        if (dual && (local_tile & 1) != mine) {
          // the other issuer's tile: step over its ring slots.  The host makes a_stages a multiple of 2 * k_chunks in this
          // mode, so a given slot only ever holds chunks of one tile parity: each barrier is waited on by ONE issuer, which
          // sees every phase of it in order (a parity wait is unsound for a waiter that skipped a phase).
          sa += p.k_chunks;
          if (sa >= a_stages) { sa -= a_stages; pa ^= 1; }
          continue;
        }
        // phase is the slowest tile coordinate (and absent when the phases are fused)
        const int ph0 = NPH == 1 ? t / (p.total_tiles / p.n_phases) : 0;
        uint32_t aoff[NPH * TAPS];
#pragma unroll
        for (int i = 0; i < NPH * TAPS; ++i) aoff[i] = s_aoff[ph0 * TAPS + i];
        const uint32_t acc = local_tile % NACC;
        SNB_PROF_T0
        mbar_wait(&tmem_empty[acc], ((local_tile / NACC) & 1) ^ 1);
        SNB_PROF_ADD(2)
        tc05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (NPH * BN);
        uint32_t b_lo = b_lo0 + static_cast<uint32_t>(ph0 * p.k_chunks * TAPS) * B_BYTES16;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          SNB_PROF_T0
          mbar_wait(PRE ? &a_ready[sa] : &a_full[sa], pa);
          SNB_PROF_ADD(0)
          tc05_fence_after();
          const uint32_t a_lo = a_lo0 + sa * A_STAGE16;
#pragma unroll
          for (int wt = 0; wt < NPH * TAPS; ++wt) {
            const int tap = wt % TAPS;
            if (NPH == 4 && by_phase && static_cast<uint32_t>(wt / (2 * TAPS)) != mine) {   // the other issuer's phases
              if (!bres) {
                mbar_wait(&b_full[sb], pb);
                if (++sb == b_stages) { sb = 0; pb ^= 1; }
              } else {
                b_lo += B_BYTES16;
              }
              continue;
            }
            if (!bres) {
              SNB_PROF_T0
              mbar_wait(&b_full[sb], pb);
              SNB_PROF_ADD(1)
              tc05_fence_after();
              b_lo = b_lo0 + sb * B_BYTES16;
            }
#pragma unroll
            for (int k = 0; k < SWZ / 32; ++k)
              umma_ss<EB>(desc_from_halves(a_lo + aoff[wt] + 2 * k, HI_A), desc_from_halves(b_lo + 2 * k, HI_B),
                           d_tmem + (wt / TAPS) * BN, IDESC, (kc > 0 || tap > 0 || k > 0) ? 1u : 0u);
            if (!bres) {
              if (CL == 1) umma_commit(&b_empty[sb]);
              else umma_commit_mc(&b_empty[sb], static_cast<uint16_t>(0x3));
              if (++sb == b_stages) { sb = 0; pb ^= 1; }
            } else {
              b_lo += B_BYTES16;
            }
          }
          umma_commit(&a_empty[sa]);
          if (++sa == a_stages) { sa = 0; pa ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
      }
#ifdef SNB_CONV_PROFILE
      if (mine == 0) {
        prof_c[3] = static_cast<unsigned long long>(clock64() - prof_loop0);
        SNB_PROF_FLUSH(0, 0) SNB_PROF_FLUSH(1, 1) SNB_PROF_FLUSH(2, 2) SNB_PROF_FLUSH(3, 3)
        atomicAdd(&g_conv_prof[7], static_cast<unsigned long long>(local_tile));
      }
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
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const TileCoord tc = decode(t);
      for (int kc = 0; kc < p.k_chunks; ++kc) {
        const float4* sc4 = reinterpret_cast<const float4*>(p.pre_scale + kc * BK + jl * 8);
        const float4* sh4 = reinterpret_cast<const float4*>(p.pre_shift + kc * BK + jl * 8);
        const float4 s0 = __ldg(sc4), s1 = __ldg(sc4 + 1), b0 = __ldg(sh4), b1 = __ldg(sh4 + 1);
        mbar_wait(&a_full[s], par);
        uint8_t* base = smem_a + s * Cfg::HALO_STAGE_BYTES;
#pragma unroll 3
        for (int r = r0; r < kHaloRows; r += RSTEP) {
          const int hy = r / kHaloW, hx = r - hy * kHaloW;
          const bool inside = static_cast<unsigned>(tc.x0 - 1 + p.load_dx + hx) < static_cast<unsigned>(p.in_w) &&
                              static_cast<unsigned>(tc.y0 - 1 + p.load_dx + hy) < static_cast<unsigned>(p.in_h);
          const int phys = jl ^ (SWZ == 128 ? (r & 7) : ((r >> 1) & 3));   // the swizzle TMA applied to this row
          uint4* ptr = reinterpret_cast<uint4*>(base + r * SWZ + (phys << 4));
          uint4 o = make_uint4(0u, 0u, 0u, 0u);
          if (inside) {
            const uint4 v = *ptr;
            const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&v);
            const float2 x0 = __bfloat1622float2(pv[0]), x1 = __bfloat1622float2(pv[1]);
            const float2 x2 = __bfloat1622float2(pv[2]), x3 = __bfloat1622float2(pv[3]);
            o.x = pack_bf16x2(fmaxf(fmaf(x0.x, s0.x, b0.x), 0.f), fmaxf(fmaf(x0.y, s0.y, b0.y), 0.f));
            o.y = pack_bf16x2(fmaxf(fmaf(x1.x, s0.z, b0.z), 0.f), fmaxf(fmaf(x1.y, s0.w, b0.w), 0.f));
            o.z = pack_bf16x2(fmaxf(fmaf(x2.x, s1.x, b1.x), 0.f), fmaxf(fmaf(x2.y, s1.y, b1.y), 0.f));
            o.w = pack_bf16x2(fmaxf(fmaf(x3.x, s1.z, b1.z), 0.f), fmaxf(fmaf(x3.y, s1.w, b1.w), 0.f));
          }
          *ptr = o;
        }
        fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[s]);
        if (++s == static_cast<uint32_t>(p.a_stages)) { s = 0; par ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 4 + 4 * EG) {
    // ------------------------------------------------------------------ epilogue (EG groups of 128 threads)
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int grp = (warp - 4) >> 2;        // EG = 2: group g takes the local tiles of parity g = TMEM stage g
    const int row = q * 32 + lane;
    const int epi_tid = threadIdx.x - 128 - 128 * grp;
    uint32_t local_tile = 0;
    uint32_t n_store = 0;
    SNB_PROF_DECL
#ifdef SNB_CONV_PROFILE
    const long long prof_loop0 = clock64();
#endif
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++local_tile) {
      const uint32_t acc = local_tile % NACC;
      if (EG == 2 && (p.epi_groups == 2 ? (acc & 1) != static_cast<uint32_t>(grp) : grp != 0)) continue;
      TileCoord tc = decode(t);
      const uint32_t acc_ph = (local_tile / NACC) & 1;
      SNB_PROF_T0
      mbar_wait(&tmem_full[acc], acc_ph);
      SNB_PROF_ADD(0)
      tc05_fence_after();
      const uint32_t t_addr = tmem_base + acc * (NPH * BN) + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int ph = 0; ph < NPH; ++ph) {
        if (NPH > 1) tc.ph = ph;
        if (p.ext)
          epilogue_tile<BN, HEAD, TW, EB, true, EG>(p, tc, t_addr + ph * BN, &tmem_empty[acc], smem_out, n_store, row, lane,
                                                    epi_tid, ph == NPH - 1, grp);
        else
          epilogue_tile<BN, HEAD, TW, EB, false, EG>(p, tc, t_addr + ph * BN, &tmem_empty[acc], smem_out, n_store, row, lane,
                                                     epi_tid, ph == NPH - 1, grp);
      }
    }
    if (!HEAD && epi_tid == 0) tma_store_wait_all<0>();
#ifdef SNB_CONV_PROFILE
    if (warp == 4 && lane == 0) {
      prof_c[1] = static_cast<unsigned long long>(clock64() - prof_loop0);
      SNB_PROF_FLUSH(4, 0) SNB_PROF_FLUSH(5, 1)
    }
#endif
  }

  tc05_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // no CTA exits while its peer may still multicast into it
  if (warp == 2) {
    tc05_fence_after();
    tmem_dealloc(tmem_base, TCOLS);
  }
}

// ===================================================================================== first layer: conv3x3 over 3 channels
// conv1_1 of every network (Cin = 3) as a K = 32 GEMM whose operand rows are BUILT IN SHARED MEMORY: the input is the
// normalised tile as packed NHWC bf16 with 3 channels (6 bytes per pixel -- exactly the algorithmic size of SURVEY 8d, no
// 64-byte im2col row per pixel in HBM).  Per 16 x 8 patch one TMA box brings the (16+2) x (8+2) x 3 neighbourhood (zero
// filled outside the tile = the convolution's padding; widened to a 16-byte aligned start); four builder warps (one thread per pixel) assemble the 128 operand
// rows k = tap * 3 + c (27 values + 5 zeros, 64 bytes, written with the 64-byte swizzle the MMA descriptor expects); the
// MMA thread issues two K = 16 steps against the resident [Cout][32] weights; the epilogue is the common one.
// TMA fetches 16-byte granules: the innermost start coordinate must be a multiple of 16 bytes, i.e. of 8 packed pixels
// (48 bytes), so the box starts 8 pixels left of the patch and spans pixels x0 - 8 .. x0 + 18 (80 elements = 160 bytes)
constexpr int kFirstRawW = 80;
constexpr int kFirstRawLead = 8;                       // pixels between the box start and the patch origin
constexpr int kFirstRawRows = 10;
constexpr int kFirstRawStage = 1664;                   // 10 x 160 = 1600 bytes, 128-byte aligned stages
constexpr int kFirstRawStages = 4, kFirstAStages = 3;
constexpr int kFirstABytes = kBM * 64;                 // 128 rows x 32 bf16

// Warp roles (448 threads): 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2-5 = builders (one pixel per thread),
// 6-9 and 10-13 = TWO epilogue groups that take alternate tiles (accumulator stage 0 / 1): with K = 32 the MMA of a tile lasts
// ~100 cycles while its epilogue (bias, ReLU, bf16, swizzled staging, TMA store of 128 x Cout) lasts ~1000.  Measured with
// two builder warps (two pixels per thread): the issuer waited on the builders 998 of 1269 cycles per tile.
constexpr int kFirstThreads = 448;
template <int BN>
__global__ void __launch_bounds__(kFirstThreads, 1) conv_first_kernel(const __grid_constant__ ConvParams p) {
  using Cfg = ConvCfg<BN, 32, 2>;
  constexpr uint32_t IDESC = make_idesc(kBM, BN, 1u);
  constexpr int TW = 16, TH = 8;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                                   // kFirstAStages x 8 KB operand rows
  uint8_t* smem_b = smem_a + kFirstAStages * kFirstABytes;                  // BN x 64 B weights (resident)
  uint8_t* smem_out = smem_b + Cfg::B_BYTES;                                // 2 groups x 2 x OUT_BYTES
  uint8_t* smem_in = smem_out + 4 * Cfg::OUT_BYTES;                         // kFirstRawStages raw neighbourhoods
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_in + kFirstRawStages * kFirstRawStage);
  uint64_t* raw_full = bars;                          // [4]
  uint64_t* raw_empty = raw_full + kFirstRawStages;   // [4]
  uint64_t* a_ready = raw_empty + kFirstRawStages;    // [3]
  uint64_t* a_empty = a_ready + kFirstAStages;        // [3]
  uint64_t* b_full = a_empty + kFirstAStages;         // [1]
  uint64_t* tmem_full = b_full + 1;                   // [2]
  uint64_t* tmem_empty = tmem_full + 2;               // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_a);
    tma_prefetch_desc(&p.map_b);
    tma_prefetch_desc(&p.map_d[0]);
    for (int i = 0; i < kFirstRawStages; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], 4); }
    for (int i = 0; i < kFirstAStages; ++i) { mbar_init(&a_ready[i], 4); mbar_init(&a_empty[i], 1); }
    mbar_init(b_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc05_fence_before();
  __syncthreads();
  tc05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA: weights once, then raw neighbourhoods
    if (elect_one()) {
      mbar_arrive_expect_tx(b_full, Cfg::B_BYTES);
      tma_load_3d(&p.map_b, b_full, smem_b, 0, 0, 0);
      uint32_t s = 0, par = 1;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile<TW, TH>(p, t);
        mbar_wait(&raw_empty[s], par);
        mbar_arrive_expect_tx(&raw_full[s], kFirstRawRows * kFirstRawW * 2);
        tma_load_3d(&p.map_a, &raw_full[s], smem_in + s * kFirstRawStage, (tc.x0 - kFirstRawLead) * 3, tc.y0 - 1, tc.img);
        if (++s == kFirstRawStages) { s = 0; par ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      mbar_wait(b_full, 0);
      tc05_fence_after();
      const uint64_t bdesc = make_kmajor_desc<64>(smem_u32(smem_b), 8 * 64);
      uint32_t sa = 0, pa = 0, local_tile = 0;
      SNB_PROF_DECL
#ifdef SNB_CONV_PROFILE
      const long long prof_loop0 = clock64();
#endif
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++local_tile) {
        const uint32_t acc = local_tile & 1;
        SNB_PROF_T0
        mbar_wait(&tmem_empty[acc], ((local_tile >> 1) & 1) ^ 1);
        SNB_PROF_ADD(2)
        SNB_PROF_T0
        mbar_wait(&a_ready[sa], pa);
        SNB_PROF_ADD(0)
        tc05_fence_after();
        const uint64_t adesc = make_kmajor_desc<64>(smem_u32(smem_a + sa * kFirstABytes), 8 * 64);
        umma_bf16_ss(adesc, bdesc, tmem_base + acc * BN, IDESC, 0u);
        umma_bf16_ss(adesc + 2, bdesc + 2, tmem_base + acc * BN, IDESC, 1u);
        umma_commit(&a_empty[sa]);
        umma_commit(&tmem_full[acc]);
        if (++sa == kFirstAStages) { sa = 0; pa ^= 1; }
      }
#ifdef SNB_CONV_PROFILE
      prof_c[3] = static_cast<unsigned long long>(clock64() - prof_loop0);
      SNB_PROF_FLUSH(0, 0) SNB_PROF_FLUSH(2, 2) SNB_PROF_FLUSH(3, 3)
      atomicAdd(&g_conv_prof[7], static_cast<unsigned long long>(local_tile));
#endif
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ builders: one thread = one pixel = one 64-byte row
    const int r = threadIdx.x - 64;
    const int px = r & 15, py = r >> 4;
    uint32_t s = 0, ps = 0, sa = 0, pa = 1;
    SNB_PROF_DECL
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      SNB_PROF_T0
      mbar_wait(&raw_full[s], ps);
      SNB_PROF_ADD(1)
      const unsigned short* raw = reinterpret_cast<const unsigned short*>(smem_in + s * kFirstRawStage);
      uint4 rowv[4];
      {
        unsigned short el[32];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          // k = tap * 3 + c = dy * 9 + (dx * 3 + c): the nine values of one neighbourhood row are contiguous in the raw box
          const unsigned short* q = raw + (py + dy) * kFirstRawW + (px + kFirstRawLead - 1) * 3;
#pragma unroll
          for (int e = 0; e < 9; ++e) el[dy * 9 + e] = q[e];
        }
#pragma unroll
        for (int k = 27; k < 32; ++k) el[k] = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          rowv[c] = make_uint4(el[c * 8 + 0] | ((uint32_t)el[c * 8 + 1] << 16), el[c * 8 + 2] | ((uint32_t)el[c * 8 + 3] << 16),
                               el[c * 8 + 4] | ((uint32_t)el[c * 8 + 5] << 16), el[c * 8 + 6] | ((uint32_t)el[c * 8 + 7] << 16));
      }
      fence_proxy_async_smem();                        // generic-proxy reads before the async-proxy refill of the stage
      __syncwarp();
      if (lane == 0) mbar_arrive(&raw_empty[s]);       // the raw stage is in registers now (rowv depends on every load)
      if (++s == kFirstRawStages) { s = 0; ps ^= 1; }
      SNB_PROF_T0
      mbar_wait(&a_empty[sa], pa);
      SNB_PROF_ADD(0)
      {
        const int sw = (r >> 1) & 3;                   // 64-byte swizzle: 16-byte chunk index ^= (row >> 1) & 3
        uint8_t* row = smem_a + sa * kFirstABytes + r * 64;
#pragma unroll
        for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(row + ((c ^ sw) << 4)) = rowv[c];
      }
      fence_proxy_async_smem();                        // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_ready[sa]);
      if (++sa == kFirstAStages) { sa = 0; pa ^= 1; }
    }
#ifdef SNB_CONV_PROFILE
    if (threadIdx.x == 64) { SNB_PROF_FLUSH(6, 0) SNB_PROF_FLUSH(1, 1) }   // builder: [6] operand stage free, [1] raw box landed
#endif
  } else {
    // ------------------------------------------------------------------ two epilogue groups (128 threads each)
    const int grp = (warp - 6) >> 2;                   // 0: warps 6-9 (even tiles), 1: warps 10-13 (odd tiles)
    const int q = warp & 3;                            // TMEM lane quarter (any four consecutive warps cover all four)
    const int row = q * 32 + lane;
    const int epi_tid = (warp - 6 - grp * 4) * 32 + lane;
    uint8_t* my_out = smem_out + grp * 2 * Cfg::OUT_BYTES;
    constexpr int CW = BN < 64 ? BN : 64;
    constexpr int OUT_SWZ = CW * 2;
    uint32_t local_tile = 0, n_store = 0;
    SNB_PROF_DECL
#ifdef SNB_CONV_PROFILE
    const long long prof_loop0 = clock64();
#endif
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++local_tile) {
      if ((local_tile & 1) != static_cast<uint32_t>(grp)) continue;
      const TileCoord tc = decode_tile<TW, TH>(p, t);
      const uint32_t acc = grp;
      SNB_PROF_T0
      mbar_wait(&tmem_full[acc], (local_tile >> 1) & 1);
      SNB_PROF_ADD(0)
      tc05_fence_after();
      const uint32_t t_addr = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
      // bias, ReLU, bf16, swizzled staging, TMA store (the common epilogue with this group's own staging buffers and barrier)
      static_assert(BN <= 64, "one store chunk per tile");
      uint8_t* sout = my_out + (n_store & 1) * Cfg::OUT_BYTES;
      if (epi_tid == 0) tma_store_wait_read<1>();
      named_bar_sync(1 + grp, 128);
#pragma unroll
      for (int g = 0; g < CW / 32; ++g) {
        uint32_t v[32];
        tmem_ld_32x32(t_addr + g * 32, v);
        tmem_ld_wait();
        const float4* bias4 = reinterpret_cast<const float4*>(p.bias + g * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b0 = __ldg(bias4 + 2 * j), b1 = __ldg(bias4 + 2 * j + 1);
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j * 8 + e]);
          add2(f[0], f[1], b0.x, b0.y);
          add2(f[2], f[3], b0.z, b0.w);
          add2(f[4], f[5], b1.x, b1.y);
          add2(f[6], f[7], b1.z, b1.w);
          const int sw = OUT_SWZ == 128 ? (row & 7) : ((row >> 1) & 3);
          const int chunk = g * 4 + j;
          uint4 pk;
          if (p.relu)
            pk = make_uint4(pack_bf16x2_relu(f[0], f[1]), pack_bf16x2_relu(f[2], f[3]), pack_bf16x2_relu(f[4], f[5]),
                            pack_bf16x2_relu(f[6], f[7]));
          else
            pk = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
          *reinterpret_cast<uint4*>(sout + row * OUT_SWZ + ((chunk ^ sw) << 4)) = pk;
        }
      }
      tc05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);    // accumulator stage back to the MMA warp
      fence_proxy_async_smem();
      named_bar_sync(1 + grp, 128);
      if (epi_tid == 0) {
        tma_store_4d(&p.map_d[0], sout, 0, tc.x0, tc.y0, tc.img);
        tma_store_commit();
      }
      ++n_store;
    }
    if (epi_tid == 0) tma_store_wait_all<0>();
#ifdef SNB_CONV_PROFILE
    if (warp == 6 && lane == 0) {
      prof_c[1] = static_cast<unsigned long long>(clock64() - prof_loop0);
      SNB_PROF_FLUSH(4, 0) SNB_PROF_FLUSH(5, 1)
    }
#endif
  }

  tc05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc05_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// --------------------------------------------------------------------------------------------- host side
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

// bf16 / fp32 tensor map over up to 4 dims; strides in bytes for dims 1..rank-1 (shared with conv_scatter.cu)
int encode_map(CUtensorMap* map, void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                      const uint32_t* box, int swizzle_bytes, int elem_bytes = 2) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(SNB_E_CUDA, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 0  ? CU_TENSOR_MAP_SWIZZLE_NONE
                                                : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = fn(map, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SNB_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return SNB_OK;
}

struct KernelChoice {
  const void* fn;        // v1 kernel
  const void* fn_halo9;  // v2 kernel, 9 taps (conv3x3)
  const void* fn_halo4;  // v2 kernel, 4 taps per phase (ConvTranspose k4 s2)
  const void* fn_halo4f; // v2 kernel, ConvTranspose with the 4 phases fused into one tile (BN <= 64), or null
  const void* fn_pair9;  // v2 kernel in CTA pairs with multicast weights (BN >= 128), or null
  const void* fn_pair4;
  const void* fn_pre9;   // v2 kernel, conv3x3 with the fused pre-activation prologue (BN = 32, bf16), or null
  int smem;              // v1 dynamic smem
  int bn, bk;
  int halo_stage_bytes, b_bytes, out_bytes, pool_bytes;
};

template <int BN, int BK, bool HEAD, int EB>
static const void* fused_phase_kernel() {
  if constexpr (!HEAD && BN <= 64) return reinterpret_cast<const void*>(&conv_halo_kernel<BN, BK, false, 4, 4, 1, EB>);
  else return nullptr;
}

template <int BN, int BK, bool HEAD, int TAPS, int EB>
static const void* pair_kernel() {
  // CTA pairs are a measured no-gain variant: only built for the bf16 path
  if constexpr (!HEAD && BN >= 128 && EB == 2) return reinterpret_cast<const void*>(&conv_halo_kernel<BN, BK, false, TAPS, 1, 2, EB>);
  else return nullptr;
}

template <int BN, int BK, bool HEAD, int EB>
static const void* pre_kernel() {
  if constexpr (!HEAD && BN == 32 && EB == 2) return reinterpret_cast<const void*>(&conv_halo_kernel<BN, BK, false, 9, 1, 1, EB, true>);
  else return nullptr;
}

template <int BN, int BK, bool HEAD, int EB>
static KernelChoice choice() {
  using Cfg = ConvCfg<BN, BK, EB>;
  return KernelChoice{reinterpret_cast<const void*>(&conv_igemm_kernel<BN, BK, HEAD, EB>),
                      reinterpret_cast<const void*>(&conv_halo_kernel<BN, BK, HEAD, 9, 1, 1, EB>),
                      reinterpret_cast<const void*>(&conv_halo_kernel<BN, BK, HEAD, 4, 1, 1, EB>),
                      fused_phase_kernel<BN, BK, HEAD, EB>(), pair_kernel<BN, BK, HEAD, 9, EB>(),
                      pair_kernel<BN, BK, HEAD, 4, EB>(), pre_kernel<BN, BK, HEAD, EB>(), Cfg::SMEM_BYTES, BN, BK, Cfg::HALO_STAGE_BYTES, Cfg::B_BYTES,
                      HEAD ? 0 : Cfg::OUT_BYTES, HEAD ? 0 : Cfg::POOL_BYTES};
}

// bk in elements; eb = element bytes (2 = bf16, 4 = tf32): the row is bk * eb = 64 or 128 bytes either way
static bool pick_kernel(int bn, int bk, bool head, int eb, KernelChoice* out) {
#define SNB_PICK(BN_, BK_, EB_)                          \
  if (bn == BN_ && bk == BK_ && eb == EB_) {             \
    *out = head ? choice<32, BK_, true, EB_>() : choice<BN_, BK_, false, EB_>(); \
    return true;                                         \
  }
  if (head && bn != 32) return false;
  SNB_PICK(32, 32, 2)
  SNB_PICK(32, 64, 2)
  SNB_PICK(32, 16, 4)
  SNB_PICK(32, 32, 4)
  if (head) return false;
  SNB_PICK(64, 32, 2)
  SNB_PICK(64, 64, 2)
  SNB_PICK(128, 32, 2)
  SNB_PICK(128, 64, 2)
  SNB_PICK(256, 32, 2)
  SNB_PICK(256, 64, 2)
  SNB_PICK(64, 16, 4)
  SNB_PICK(64, 32, 4)
  SNB_PICK(128, 16, 4)
  SNB_PICK(128, 32, 4)
  SNB_PICK(256, 16, 4)
  SNB_PICK(256, 32, 4)
#undef SNB_PICK
  return false;
}

// SNB_CONV_MODE: 0 = tap mode everywhere, 1 = halo mode with streamed weights, 2 = + resident weights where they
// fit, 3 = + ConvTranspose phases fused into one tile (Cout <= 64) and N tile narrowed to 128 when the 256-wide tiling
// would leave the last wave mostly empty (default), 4 = + CTA pairs with multicast weights for the streamed-weight
// layers (measured: no gain, the large layers are power-bound, not L2-bound).  Read at snb_conv_create time.
static int conv_mode() {
  const char* e = std::getenv("SNB_CONV_MODE");
  if (!e || !*e) return 3;
  return std::atoi(e);
}

static double wave_efficiency(int64_t tiles, int sms) {
  const double waves = (double)tiles / sms;
  const double full = (double)((tiles + sms - 1) / sms);
  return full > 0 ? waves / full : 1.0;
}

int conv_tap_geometry(int kind, int valid, int64_t h, int64_t w, ConvTapGeom* g) {
  std::memset(g, 0, sizeof(*g));
  if (kind < SNB_CONV_3X3 || kind > SNB_CONV_2X2_ADJ) return fail(SNB_E_INVALID, "unknown conv kind %d", kind);
  const bool is_convt3 = kind == SNB_CONVT_3X3_S2 || kind == SNB_CONVT_3X3_S2_FULL;
  const bool is_convt = kind == SNB_CONVT_4X4_S2 || is_convt3;
  if (valid < 0 || valid > 2 || (valid == 2 && kind != SNB_CONV_3X3) ||
      (valid == 1 && !((kind == SNB_CONV_3X3 && h >= 3 && w >= 3) || kind == SNB_CONV_2X2 || (kind == SNB_CONV_2X2_ADJ && h >= 2 && w >= 2))))
    return fail(SNB_E_INVALID, "valid mode %d does not apply to conv kind %d at %lld x %lld", valid, kind, (long long)h, (long long)w);
  // conv2x2 (taps -1, 0): padded = (h+1) x (w+1) outputs, valid = h x w.  conv2x2-adjoint (taps 0, +1): h x w, valid = (h-1) x (w-1)
  int64_t grow = 0;
  if ((kind == SNB_CONV_2X2 && !valid) || kind == SNB_CONVT_3X3_S2_FULL) grow = 1;
  if (kind == SNB_CONV_2X2_ADJ && valid) grow = -1;
  if (kind == SNB_CONV_3X3) grow = valid == 1 ? -2 : (valid == 2 ? 2 : 0);
  g->load_off = kind == SNB_CONV_3X3 ? (valid == 1 ? 1 : (valid == 2 ? -1 : 0)) : 0;
  g->grid_h = h + grow;
  g->grid_w = w + grow;
  const int64_t osc = is_convt ? 2 : 1;   // output pixels per grid pixel and axis (before out_upsample2x)
  g->out_h = kind == SNB_CONVT_3X3_S2_FULL ? 2 * h + 1 : g->grid_h * osc;
  g->out_w = kind == SNB_CONVT_3X3_S2_FULL ? 2 * w + 1 : g->grid_w * osc;
  g->n_phases = is_convt ? 4 : 1;
  g->taps = kind == SNB_CONV_3X3 ? 9 : (kind == SNB_CONV_1X1 ? 1 : 4);
  g->flop_taps = is_convt3 ? 9.0 : (is_convt ? 16.0 : (double)g->taps);
  if (kind == SNB_CONV_3X3) {
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        g->tap_dy[0][ky * 3 + kx] = static_cast<int8_t>(ky - 1);
        g->tap_dx[0][ky * 3 + kx] = static_cast<int8_t>(kx - 1);
      }
  } else if (kind == SNB_CONV_2X2 || kind == SNB_CONV_2X2_ADJ) {
    // k2 s1 p1: out[y][x] = sum in[y+ky-1][x+kx-1] * W[ky][kx]; tap = ky*2 + kx; its adjoint reads in[y+ky][x+kx]
    const int o = kind == SNB_CONV_2X2 ? -1 : 0;
    for (int ky = 0; ky < 2; ++ky)
      for (int kx = 0; kx < 2; ++kx) {
        g->tap_dy[0][ky * 2 + kx] = static_cast<int8_t>(ky + o);
        g->tap_dx[0][ky * 2 + kx] = static_cast<int8_t>(kx + o);
      }
  } else if (is_convt) {
    // k4 s2 p1: out[2y+py] gathers in[y+dy] * W[ky]:  py=0: (dy=0,ky=1), (dy=-1,ky=3);  py=1: (dy=+1,ky=0), (dy=0,ky=2)
    // k3 s2 p0: out[2y+py] gathers                    py=0: (dy=0,ky=0), (dy=-1,ky=2);  py=1: (dy=0,ky=1), (unused slot)
    // the packed weight tap order is (ty, tx) with ty, tx in {0,1} following that list
    const int d4[2][2] = {{0, -1}, {1, 0}}, d3[2][2] = {{0, -1}, {0, 0}};
    const int (*dlist)[2] = kind == SNB_CONVT_4X4_S2 ? d4 : d3;   // the FULL variant only widens the tile grid
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px)
        for (int ty = 0; ty < 2; ++ty)
          for (int tx = 0; tx < 2; ++tx) {
            g->tap_dy[py * 2 + px][ty * 2 + tx] = static_cast<int8_t>(dlist[py][ty]);
            g->tap_dx[py * 2 + px][ty * 2 + tx] = static_cast<int8_t>(dlist[px][tx]);
          }
  }
  return SNB_OK;
}

}  // namespace snb

struct snb_conv {
  snb::ConvParams params;
  snb::KernelChoice kernel;
  const void* fn;
  int smem;
  int grid;
  int cluster;   // 1, or 2 for the CTA-pair kernels
  int threads;   // 256, or 512 with the pre-activation prologue warps
  double flops;
};

using namespace snb;

// conv1_1 from the packed 3-channel tile (SNB_CONV_FIRST_3X3): see conv_first_kernel
static int create_first(const snb_conv_desc* d, snb_conv** out) {
  if (d->dtype != SNB_CONV_BF16) return fail(SNB_E_UNSUPPORTED, "the 3-channel first-layer kernel is bf16 only");
  if (d->cin != 3 || d->in_cstride != 3) return fail(SNB_E_INVALID, "first-layer input must be packed NHWC with 3 channels");
  if (d->cout != 32 && d->cout != 64) return fail(SNB_E_UNSUPPORTED, "first-layer cout must be 32 or 64, got %lld", (long long)d->cout);
  if (d->w % 8) return fail(SNB_E_INVALID, "first-layer width must be a multiple of 8 (16-byte image rows)");
  if (d->valid || d->d_head_w || d->d_pool_out || d->out_upsample2x || d->d_pre_scale || d->d_residual || d->act_slope != 0.f)
    return fail(SNB_E_INVALID, "the first-layer kernel takes bias + optional ReLU only");
  if (!d->d_in || !d->d_weight || !d->d_bias || !d->d_out || d->out_cstride < d->cout || d->out_cstride % 8)
    return fail(SNB_E_INVALID, "bad first-layer tensors");
  if ((reinterpret_cast<uintptr_t>(d->d_in) & 15) || (reinterpret_cast<uintptr_t>(d->d_out) & 15) ||
      (reinterpret_cast<uintptr_t>(d->d_weight) & 15) || (reinterpret_cast<uintptr_t>(d->d_bias) & 15))
    return fail(SNB_E_INVALID, "tensor pointers must be 16-byte aligned");
  const int sms = sm_count();
  if (sms <= 0) return fail(SNB_E_CUDA, "no CUDA device");
  snb_conv* c = new (std::nothrow) snb_conv();
  if (!c) return fail(SNB_E_INVALID, "out of host memory");
  c->cluster = 1;
  c->threads = kFirstThreads;
  ConvParams& p = c->params;
  std::memset(&p, 0, sizeof(p));
  const int bn = (int)d->cout;
  p.n_phases = 1; p.taps = 1; p.k_chunks = 1; p.n_tiles = 1;
  p.tiles_x = static_cast<int32_t>((d->w + 15) / 16);
  p.tiles_y = static_cast<int32_t>((d->h + 7) / 8);
  p.n_img = static_cast<int32_t>(d->n);
  const int64_t total = (int64_t)p.n_img * p.tiles_y * p.tiles_x;
  if (total > INT32_MAX) { delete c; return fail(SNB_E_UNSUPPORTED, "too many tiles"); }
  p.total_tiles = static_cast<int32_t>(total);
  p.relu = d->relu;
  p.bias = d->d_bias;
  p.out_w = static_cast<int32_t>(d->w); p.out_h = static_cast<int32_t>(d->h);
  p.in_w = p.out_w; p.in_h = p.out_h;
  int rc;
  {
    uint64_t dims[3] = {(uint64_t)d->w * 3, (uint64_t)d->h, (uint64_t)d->n};
    uint64_t str[2] = {(uint64_t)d->w * 6, (uint64_t)d->h * d->w * 6};
    uint32_t box[3] = {kFirstRawW, kFirstRawRows, 1};
    rc = encode_map(&p.map_a, const_cast<void*>(d->d_in), 3, dims, str, box, 0, 2);
    if (rc) { delete c; return rc; }
  }
  {
    uint64_t dims[3] = {32, (uint64_t)d->cout, 1};
    uint64_t str[2] = {64, (uint64_t)d->cout * 64};
    uint32_t box[3] = {32, (uint32_t)bn, 1};
    rc = encode_map(&p.map_b, const_cast<void*>(d->d_weight), 3, dims, str, box, 64, 2);
    if (rc) { delete c; return rc; }
  }
  const int cw = bn < 64 ? bn : 64;
  {
    uint64_t dims[4] = {(uint64_t)d->cout, (uint64_t)d->w, (uint64_t)d->h, (uint64_t)d->n};
    uint64_t str[3] = {(uint64_t)d->out_cstride * 2, (uint64_t)d->w * d->out_cstride * 2, (uint64_t)d->h * d->w * d->out_cstride * 2};
    uint32_t box[4] = {(uint32_t)cw, 16, 8, 1};
    rc = encode_map(&p.map_d[0], d->d_out, 4, dims, str, box, cw * 2, 2);
    if (rc) { delete c; return rc; }
  }
  c->fn = bn == 64 ? reinterpret_cast<const void*>(&conv_first_kernel<64>) : reinterpret_cast<const void*>(&conv_first_kernel<32>);
  const int out_bytes = kBM * cw * 2;
  c->smem = 1024 + kFirstAStages * kFirstABytes + bn * 64 + 4 * out_bytes + kFirstRawStages * kFirstRawStage + 1024;
  cudaError_t e = cudaFuncSetAttribute(c->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
  if (e != cudaSuccess) { delete c; return fail(SNB_E_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); }
  c->grid = std::min<int>(p.total_tiles, sms);
  c->flops = 2.0 * (double)d->n * d->h * d->w * 27.0 * (double)d->cout;
  *out = c;
  return SNB_OK;
}

extern "C" int snb_conv_create(const snb_conv_desc* d, snb_conv** out) {
  if (!d || !out) return fail(SNB_E_INVALID, "snb_conv_create: null argument");
  *out = nullptr;
  if (d->kind == SNB_CONV_FIRST_3X3) return create_first(d, out);
  if (d->n <= 0 || d->h <= 0 || d->w <= 0) return fail(SNB_E_INVALID, "bad input shape");
  ConvTapGeom tgeo;
  if (int rc = conv_tap_geometry(d->kind, d->valid, d->h, d->w, &tgeo)) return rc;
  const bool is_convt3 = d->kind == SNB_CONVT_3X3_S2 || d->kind == SNB_CONVT_3X3_S2_FULL;
  const bool is_convt = d->kind == SNB_CONVT_4X4_S2 || is_convt3;
  const int valid = d->valid;
  // tile grid (where accumulators are computed), halo-origin offset and output extent per kind
  const int grid_pad = (tgeo.grid_h != d->h) ? 1 : 0;          // the tile grid differs from the input grid
  const int load_dx = tgeo.load_off;
  const int64_t grid_h = tgeo.grid_h, grid_w = tgeo.grid_w;
  const int64_t out_h = tgeo.out_h, out_w = tgeo.out_w;
  if (d->dtype != SNB_CONV_BF16 && d->dtype != SNB_CONV_TF32) return fail(SNB_E_INVALID, "unknown conv dtype %d", d->dtype);
  const int eb = d->dtype == SNB_CONV_TF32 ? 4 : 2;     // element bytes of activations and weights
  const int cmul = 64 / eb;                             // channel granularity: one 64-byte operand row
  const int calign = 16 / eb;                           // slab strides are whole 16-byte vectors
  if (d->cin <= 0 || d->cin % cmul != 0)
    return fail(SNB_E_INVALID, "cin=%lld must be a positive multiple of %d", (long long)d->cin, cmul);
  if (d->cout <= 0 || d->cout % 32 != 0) return fail(SNB_E_INVALID, "cout=%lld must be a positive multiple of 32", (long long)d->cout);
  if (d->in_cstride < d->cin || d->in_cstride % calign != 0) return fail(SNB_E_INVALID, "bad in_cstride");
  const bool head = d->d_head_w != nullptr;
  if (head && (d->cout != 32 || is_convt || !d->d_head_out))
    return fail(SNB_E_INVALID, "fused head needs cout == 32, a plain conv and an output pointer");
  if (!head && (!d->d_out || d->out_cstride < d->cout || d->out_cstride % calign != 0))
    return fail(SNB_E_INVALID, "bad output slab");
  if (!d->d_in || !d->d_weight || !d->d_bias) return fail(SNB_E_INVALID, "null tensor pointer");
  const bool up2x = d->out_upsample2x != 0;
  if (up2x && (head || is_convt))
    return fail(SNB_E_INVALID, "out_upsample2x applies to conv3x3 / conv1x1 without a fused head");
  const bool pre = d->d_pre_scale != nullptr;
  if (pre && (!d->d_pre_shift || d->kind != SNB_CONV_3X3 || head || eb != 2 || d->cout != 32 ||
              (reinterpret_cast<uintptr_t>(d->d_pre_scale) & 15) || (reinterpret_cast<uintptr_t>(d->d_pre_shift) & 15)))
    return fail(SNB_E_INVALID, "fused pre-activation needs a bf16 conv3x3 with cout == 32 and 16-byte aligned scale / shift");
  const bool has_res = d->d_residual != nullptr;
  if (has_res && (is_convt || head || up2x || d->res_cstride < d->cout || d->res_cstride % calign != 0 ||
                  (reinterpret_cast<uintptr_t>(d->d_residual) & 15)))
    return fail(SNB_E_INVALID, "the residual operand needs a plain conv without head and a 16-byte aligned slab");
  const bool pool = d->d_pool_out != nullptr;
  if (pool && (valid != 0 || head || d->kind != SNB_CONV_3X3 || (d->h & 1) || (d->w & 1) || d->pool_cstride < d->cout ||
               d->pool_cstride % calign != 0 || (reinterpret_cast<uintptr_t>(d->d_pool_out) & 15)))
    return fail(SNB_E_INVALID, "fused max-pool needs a conv3x3 without head, even h and w and a valid pooled slab");
  if ((reinterpret_cast<uintptr_t>(d->d_in) & 15) || (reinterpret_cast<uintptr_t>(d->d_out) & 15) ||
      (reinterpret_cast<uintptr_t>(d->d_weight) & 15) || (reinterpret_cast<uintptr_t>(d->d_bias) & 15))
    return fail(SNB_E_INVALID, "tensor pointers must be 16-byte aligned");

  const int mode = conv_mode();
  const int sms = sm_count();
  if (sms <= 0) return fail(SNB_E_CUDA, "no CUDA device");
  const int bk = (d->cin * eb) % 128 == 0 ? 128 / eb : 64 / eb;   // elements per K chunk: a 128- or 64-byte row
  int bn = 32;
  if (d->cout % 256 == 0) bn = 256;
  else if (d->cout % 128 == 0) bn = 128;
  else if (d->cout % 64 == 0) bn = 64;
  if (mode >= 3 && bn == 256) {
    // wave quantisation: with few tiles (deep, low-resolution layers) a 128-wide N tile fills the last wave better
    const int64_t m_tiles = (int64_t)(is_convt ? 4 : 1) * d->n * ((grid_h + 15) / 16) * ((grid_w + 7) / 8);
    const double e256 = wave_efficiency(m_tiles * (d->cout / 256), sms);
    const double e128 = wave_efficiency(m_tiles * (d->cout / 128), sms);
    if (e256 < 0.85 && e128 > e256 + 0.05) bn = 128;
  }
  KernelChoice kc;
  if (!pick_kernel(bn, bk, head, eb, &kc)) return fail(SNB_E_UNSUPPORTED, "no kernel for BN=%d BK=%d head=%d eb=%d", bn, bk, (int)head, eb);

  snb_conv* c = new (std::nothrow) snb_conv();
  if (!c) return fail(SNB_E_INVALID, "out of host memory");
  c->cluster = 1;
  c->threads = 256;
  ConvParams& p = c->params;
  std::memset(&p, 0, sizeof(p));
  c->kernel = kc;

  p.n_phases = tgeo.n_phases;
  p.taps = tgeo.taps;
  static_assert(sizeof(p.tap_dy) == sizeof(tgeo.tap_dy) && sizeof(p.tap_dx) == sizeof(tgeo.tap_dx), "tap tables");
  std::memcpy(p.tap_dy, tgeo.tap_dy, sizeof(p.tap_dy));
  std::memcpy(p.tap_dx, tgeo.tap_dx, sizeof(p.tap_dx));
  p.k_chunks = static_cast<int32_t>(d->cin / bk);
  p.n_tiles = static_cast<int32_t>(d->cout / bn);

  // ---- main-loop variant and pipeline shape
  const bool halo = mode >= 1 && d->kind != SNB_CONV_1X1;
  if ((valid != 0 || grid_pad || has_res) && !halo && d->kind != SNB_CONV_1X1) {
    delete c;
    return fail(SNB_E_UNSUPPORTED, "this layer shape needs halo mode (SNB_CONV_MODE >= 1)");
  }
  if (pre && (!halo || kc.fn_pre9 == nullptr)) {
    delete c;
    return fail(SNB_E_UNSUPPORTED, "fused pre-activation is only available in halo mode (SNB_CONV_MODE >= 1)");
  }
  if (pool && !halo) {
    delete c;
    return fail(SNB_E_UNSUPPORTED, "fused max-pool is only available in halo mode (SNB_CONV_MODE >= 1)");
  }
  const bool fuse_phases = halo && mode >= 3 && is_convt && kc.fn_halo4f != nullptr;
  int tile_w = 16, tile_h = 8;
  c->fn = kc.fn;
  c->smem = kc.smem;
  if (halo) {
    tile_w = 8;
    tile_h = 16;
    const int pool_bytes = pool ? kc.pool_bytes : 0;   // staging of the pooled tile only where a pool is fused
    int fixed = 1024 /*align*/ + 1024 /*ctrl*/ + 2 * kc.out_bytes + 2 * pool_bytes;
    const int64_t w_slots = (int64_t)p.n_phases * p.taps * p.k_chunks;
    const int64_t w_bytes = w_slots * kc.b_bytes;
    int a_stages = std::min<int>(3, std::max<int>(2, p.k_chunks + 1));
    bool bres = mode >= 2 && p.n_tiles == 1 && fixed + 2 * kc.halo_stage_bytes + w_bytes <= kSmemBudget;
    p.out_bufs = 2;
    if (std::getenv("SNB_OUT_BUFS1_BRES") && !bres && mode >= 2 && p.n_tiles == 1 && kc.bn > 64 && !pre &&
        fixed - kc.out_bytes - pool_bytes + 2 * kc.halo_stage_bytes + w_bytes <= kSmemBudget) {
      // experiment: resident weights at the price of the second staging buffer.  Measured on conv 64 -> 128 before the pool
      // staging became conditional: no gain, the single-buffered epilogue (3300 cycles per tile) becomes the bound.
      bres = true;
      p.out_bufs = 1;
      fixed -= kc.out_bytes + pool_bytes;
    }
    int b_stages = 0;
    if (bres) {
      while (a_stages > 2 && fixed + a_stages * kc.halo_stage_bytes + w_bytes > kSmemBudget) --a_stages;
      // resident weights leave room: deepen the activation ring (short K loops need several tiles of TMA lookahead)
      int a_cap = kMaxAStages;
      if (const char* e = std::getenv("SNB_A_STAGES")) a_cap = std::max(2, std::min(kMaxAStages, std::atoi(e)));
      while (a_stages < a_cap && fixed + (a_stages + 1) * kc.halo_stage_bytes + w_bytes <= kSmemBudget) ++a_stages;
      c->smem = fixed + a_stages * kc.halo_stage_bytes + (int)w_bytes;
    } else {
      a_stages = 2;
      // (SNB_OUT_BUFS = 1: one output staging buffer, the room goes to the weight ring -- measured 1 % SLOWER on the
      // BN = 256 layers: their issuer waits on weight slots ~15 % of the time but runs ahead of the pipe, the pipe does not starve)
      int nb = 2;
      if (const char* e = std::getenv("SNB_OUT_BUFS")) nb = std::atoi(e) == 1 ? 1 : 2;
      if (kc.bn <= 64 || pre) nb = 2;   // two epilogue groups own one buffer each
      const int fixed_s = fixed - (2 - nb) * (kc.out_bytes + pool_bytes);
      p.out_bufs = nb;
      b_stages = (kSmemBudget - fixed_s - a_stages * kc.halo_stage_bytes) / kc.b_bytes;
      if (b_stages > kMaxBStages) b_stages = kMaxBStages;
      if (const char* e = std::getenv("SNB_B_STAGES")) {   // pipeline-depth experiments
        const int v = std::atoi(e);
        if (v >= 2 && v < b_stages) b_stages = v;
      }
      if (b_stages < 2) {
        delete c;
        return fail(SNB_E_UNSUPPORTED, "halo pipeline does not fit in shared memory");
      }
      // a third activation stage if there is room left
      if (fixed_s + 3 * kc.halo_stage_bytes + b_stages * kc.b_bytes <= kSmemBudget) a_stages = 3;
      c->smem = fixed_s + a_stages * kc.halo_stage_bytes + b_stages * kc.b_bytes;
    }
    p.a_stages = a_stages;
    p.b_stages = b_stages;
    p.bres = bres ? 1 : 0;
    {
      // second MMA-issuing warp: narrow tiles (an N <= 64 MMA retires faster than one thread issues it); SNB_DUAL_MMA = 0 / 1
      // forces it off / on for every halo layer (A/B runs)
      // By tile (dual_mma = 1) only with resident weights and an activation ring of at least two tiles: a streamed weight
      // ring holds a fraction of a tile, the two issuers would just take turns.
      int dual = (bn <= 64 || bres) ? 1 : 0;
      if (std::getenv("SNB_DUAL_NARROW_ONLY")) dual = bn <= 64 ? 1 : 0;
      if (const char* e = std::getenv("SNB_DUAL_MMA")) dual = std::atoi(e) != 0;
      p.dual_mma = (dual && bres && a_stages >= 2 * p.k_chunks) ? 1 : 0;
      if (p.dual_mma) {
        a_stages -= a_stages % (2 * p.k_chunks);   // a ring slot then belongs to one issuer for good (see the skip in the kernel)
        p.a_stages = a_stages;
      }
      // Fused ConvTranspose phases: split by phase pair instead (no constraint on the activation ring: both issuers consume
      // every stage, its "free" barrier counts two arrivals).  Resident weights only: with a streamed weight ring the
      // non-owner has to OBSERVE the owner's slots, and an observer that falls two fills behind misreads the barrier parity
      // and waits for ever -- compute-sanitizer's timing produced exactly that hang (and the split gained nothing there).
      if (dual && fuse_phases && bres) p.dual_mma = 2;
    }
    c->fn = p.taps == 9 ? kc.fn_halo9 : (fuse_phases ? kc.fn_halo4f : kc.fn_halo4);
    if (pre) {
      c->fn = kc.fn_pre9;
      c->threads = 512;
    } else if (kc.bn <= 64) {
      c->threads = 384;   // second epilogue group (warps 8-11); SNB_EPI_GROUPS = 1 leaves it idle (A/B runs)
      p.epi_groups = 2;
      if (const char* e = std::getenv("SNB_EPI_GROUPS")) p.epi_groups = std::atoi(e) == 1 ? 1 : 2;
    }
    // CTA pairs: streamed weights, an even number of spatial tiles per phase, and a pair kernel for this shape
    const int64_t m_tiles = (int64_t)d->n * ((grid_h + tile_h - 1) / tile_h) * ((grid_w + tile_w - 1) / tile_w);
    const void* fn_pair = p.taps == 9 ? kc.fn_pair9 : kc.fn_pair4;
    if (mode >= 4 && !bres && !fuse_phases && !pre && fn_pair && m_tiles % 2 == 0 && sms >= 2) {
      c->fn = fn_pair;
      c->cluster = 2;
      c->threads = 256;
      p.epi_groups = 0;
      p.dual_mma = 0;   // the multicast weight ring is consumed in lock step by the pair
      p.m_pairs = static_cast<int32_t>(m_tiles / 2);
    }
  }

  p.tiles_x = static_cast<int32_t>((grid_w + tile_w - 1) / tile_w);
  p.tiles_y = static_cast<int32_t>((grid_h + tile_h - 1) / tile_h);
  p.n_img = static_cast<int32_t>(d->n);
  const int64_t total = static_cast<int64_t>(fuse_phases ? 1 : p.n_phases) * p.n_img * p.tiles_y * p.tiles_x * p.n_tiles;
  if (total > INT32_MAX) {
    delete c;
    return fail(SNB_E_UNSUPPORTED, "too many tiles");
  }
  p.total_tiles = static_cast<int32_t>(total);
  p.relu = d->relu;
  p.bias = d->d_bias;
  p.head_w = d->d_head_w;
  p.head_b = d->head_b;
  p.head_sigmoid = d->head_sigmoid;
  p.head_out = d->d_head_out;
  p.pre_scale = d->d_pre_scale;
  p.pre_shift = d->d_pre_shift;
  p.out_w = static_cast<int32_t>(out_w);
  p.out_h = static_cast<int32_t>(out_h);
  p.in_w = static_cast<int32_t>(d->w);
  p.in_h = static_cast<int32_t>(d->h);
  p.load_dx = load_dx;
  p.act_slope = d->act_slope;
  p.residual = d->d_residual;
  p.res_cstride = static_cast<int32_t>(d->res_cstride);
  p.res_after_act = d->res_after_act;
  p.ext = (has_res || d->act_slope != 0.f) ? 1 : 0;

  int rc;
  const uint64_t E = (uint64_t)eb;
  {
    uint64_t dims[4] = {(uint64_t)d->cin, (uint64_t)d->w, (uint64_t)d->h, (uint64_t)d->n};
    uint64_t str[3] = {(uint64_t)d->in_cstride * E, (uint64_t)d->w * d->in_cstride * E,
                       (uint64_t)d->h * d->w * d->in_cstride * E};
    uint32_t box[4] = {(uint32_t)bk, (uint32_t)(halo ? kHaloW : tile_w), (uint32_t)(halo ? kHaloH : tile_h), 1};
    rc = encode_map(&p.map_a, const_cast<void*>(d->d_in), 4, dims, str, box, bk * eb, eb);
    if (rc) { delete c; return rc; }
  }
  {
    uint64_t dims[3] = {(uint64_t)d->cin, (uint64_t)d->cout, (uint64_t)(p.n_phases * p.taps)};
    uint64_t str[2] = {(uint64_t)d->cin * E, (uint64_t)d->cout * d->cin * E};
    uint32_t box[3] = {(uint32_t)bk, (uint32_t)bn, 1};
    rc = encode_map(&p.map_b, const_cast<void*>(d->d_weight), 3, dims, str, box, bk * eb, eb);
    if (rc) { delete c; return rc; }
    if (c->cluster == 2) {
      uint32_t half[3] = {(uint32_t)bk, (uint32_t)(bn / 2), 1};
      rc = encode_map(&p.map_bh, const_cast<void*>(d->d_weight), 3, dims, str, half, bk * eb, eb);
      if (rc) { delete c; return rc; }
    }
  }
  const int cw = bn < 128 / eb ? bn : 128 / eb;   // channels per store chunk (ConvCfg::CW)
  if (!head) {
    const int s = (is_convt || up2x) ? 2 : 1;
    // full output extent and, per phase, the sub-grid a phase writes (strided view into the full tensor)
    const int64_t ow = up2x ? 2 * out_w : out_w, oh = up2x ? 2 * out_h : out_h;
    p.up2x = up2x ? 1 : 0;
    for (int ph = 0; ph < ((up2x || is_convt) ? 4 : 1); ++ph) {
      const int py = ph / 2, px = ph % 2;
      char* base = static_cast<char*>(d->d_out) + ((int64_t)py * ow + px) * d->out_cstride * eb;
      const int64_t pw = s == 1 ? ow : (ow - px + 1) / 2, phh = s == 1 ? oh : (oh - py + 1) / 2;
      uint64_t dims[4] = {(uint64_t)d->cout, (uint64_t)pw, (uint64_t)phh, (uint64_t)d->n};
      uint64_t str[3] = {(uint64_t)s * d->out_cstride * E, (uint64_t)s * ow * d->out_cstride * E,
                         (uint64_t)oh * ow * d->out_cstride * E};
      uint32_t box[4] = {(uint32_t)cw, (uint32_t)tile_w, (uint32_t)tile_h, 1};
      rc = encode_map(&p.map_d[ph], base, 4, dims, str, box, cw * eb, eb);
      if (rc) { delete c; return rc; }
    }
  }

  p.pool = pool ? 1 : 0;
  if (pool) {
    uint64_t dims[4] = {(uint64_t)d->cout, (uint64_t)(d->w / 2), (uint64_t)(d->h / 2), (uint64_t)d->n};
    uint64_t str[3] = {(uint64_t)d->pool_cstride * E, (uint64_t)(d->w / 2) * d->pool_cstride * E,
                       (uint64_t)(d->h / 2) * (d->w / 2) * d->pool_cstride * E};
    uint32_t box[4] = {(uint32_t)cw, (uint32_t)(tile_w / 2), (uint32_t)(tile_h / 2), 1};
    rc = encode_map(&p.map_p, d->d_pool_out, 4, dims, str, box, cw * eb, eb);
    if (rc) { delete c; return rc; }
  }

  cudaError_t e = cudaFuncSetAttribute(c->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
  if (e != cudaSuccess) {
    delete c;
    return fail(SNB_E_CUDA, "cudaFuncSetAttribute(smem=%d) failed: %s", kSmemBudget, cudaGetErrorString(e));
  }
  c->grid = std::min<int>(p.total_tiles, sms);
  if (c->cluster == 2) c->grid &= ~1;   // whole CTA pairs; total_tiles is even
  // 2*MACs with the true tap counts (ConvT: every input pixel meets all 16 taps once over the 4 phases)
  // (k3 s2 p0 cropped: 9 real taps per input pixel, the other 7 slots hold zero weights)
  c->flops = is_convt ? 2.0 * (double)d->n * d->h * d->w * (double)d->cin * d->cout * tgeo.flop_taps
                      : 2.0 * (double)d->n * out_h * out_w * (double)d->cin * d->cout * tgeo.flop_taps;
  *out = c;
  return SNB_OK;
}

extern "C" int snb_conv_launch(const snb_conv* c, void* stream) {
  if (!c) return fail(SNB_E_INVALID, "snb_conv_launch: null handle");
  void* args[1] = {const_cast<ConvParams*>(&c->params)};
  cudaError_t e;
  if (c->cluster == 1) {
    e = cudaLaunchKernel(c->fn, dim3(c->grid), dim3(c->threads), args, c->smem, as_stream(stream));
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(c->grid);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = c->smem;
    cfg.stream = as_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = c->cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelExC(&cfg, c->fn, args);
  }
  if (e != cudaSuccess) return fail(SNB_E_CUDA, "conv launch failed: %s", cudaGetErrorString(e));
  return SNB_OK;
}

extern "C" void snb_conv_destroy(snb_conv* c) { delete c; }

extern "C" double snb_conv_flops(const snb_conv* c) { return c ? c->flops : 0.0; }

extern "C" int snb_conv_set_head_out(snb_conv* c, float* d_head_out) {
  if (!c || !d_head_out) return fail(SNB_E_INVALID, "snb_conv_set_head_out: null argument");
  if (c->params.head_w == nullptr) return fail(SNB_E_INVALID, "snb_conv_set_head_out: this convolution has no fused head");
  c->params.head_out = d_head_out;
  return SNB_OK;
}

#ifdef SNB_CONV_PROFILE
// profiling builds only (tools/build_rev.py --profile): read (and clear) the wait-cycle counters of conv_halo_kernel
extern "C" __attribute__((visibility("default"))) int snb_debug_conv_profile(unsigned long long* out8, int reset) {
  if (out8 && cudaMemcpyFromSymbol(out8, snb::g_conv_prof, sizeof(unsigned long long) * 8) != cudaSuccess) return 1;
  if (reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpyToSymbol(snb::g_conv_prof, z, sizeof(z)) != cudaSuccess) return 1;
  }
  return 0;
}
#endif
