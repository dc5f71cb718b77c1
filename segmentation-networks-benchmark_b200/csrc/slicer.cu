// ImageSlicer on the device: crop plan, reflect-101 split, fused normalising split, weighted overlap-add merge.
// Follows lib/tiles.py:35-161 (plan, split, merge), lib/augmentations.py:452-511 (NormalizeImage via a host
// built LUT, D4 TTA index maps) and inria_submit.py:305 (threshold).  All of it is HBM-bound byte/integer
// work: one thread per output element (x fastest) so stores are fully coalesced and loads are row segments.
#include <cuda_bf16.h>

#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "sm100_ptx.cuh"
#include "snb_internal.h"

namespace snb {

// cv2.BORDER_REFLECT_101 source index: mirror without repeating the edge sample; multi-reflection safe
// (SURVEY 8c': checked against cv2.copyMakeBorder with margins larger than the image).
__host__ __device__ __forceinline__ int64_t reflect101(int64_t p, int64_t n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
  return p;
}

// D4 views in tta_d4_aug order (lib/augmentations.py:476-491): value of view v at (i, j) is tile[si][sj].
__device__ __forceinline__ void d4_src(int v, int i, int j, int T, int& si, int& sj) {
  switch (v) {
    default: si = i; sj = j; break;                     // identity
    case 1: si = j; sj = T - 1 - i; break;              // rot90(.,1)
    case 2: si = T - 1 - i; sj = T - 1 - j; break;      // rot90(.,2)
    case 3: si = T - 1 - j; sj = i; break;              // rot90(.,3)
    case 4: si = i; sj = T - 1 - j; break;              // fliplr
    case 5: si = T - 1 - j; sj = T - 1 - i; break;      // fliplr(rot90(.,1))
    case 6: si = T - 1 - i; sj = j; break;              // fliplr(rot90(.,2))
    case 7: si = j; sj = i; break;                      // fliplr(rot90(.,3))
  }
}

// inverse: where in view v does original tile position (i, j) live (tta_d4_deaug, lib/augmentations.py:494-511)
__device__ __forceinline__ void d4_dst(int v, int i, int j, int T, int& a, int& b) {
  switch (v) {
    default: a = i; b = j; break;
    case 1: a = T - 1 - j; b = i; break;
    case 2: a = T - 1 - i; b = T - 1 - j; break;
    case 3: a = j; b = T - 1 - i; break;
    case 4: a = i; b = T - 1 - j; break;
    case 5: a = T - 1 - j; b = T - 1 - i; break;
    case 6: a = T - 1 - i; b = j; break;
    case 7: a = j; b = i; break;
  }
}

struct BorderPixel {
  uint8_t bytes[64];
};

// ---------------------------------------------------------------------------------------------- split_hwc
// One warp per tile row.  Almost every output byte of a row belongs to ONE contiguous run of source bytes (the pixels whose
// source column lies inside the image), so the run is copied as aligned 16-byte stores fed by two aligned 16-byte loads and
// a funnel shift (source and destination are misaligned by an arbitrary byte count: 3-byte pixels, odd margins); only the
// reflected / constant border pixels at the two ends of a row and the sub-16-byte fringes of the run go through the
// per-byte reflect-101 map.  Bit-exact for any element size.  (The border pixel is a __grid_constant__ parameter: indexed
// dynamically it would otherwise be copied byte by byte into every thread's local memory, which ncu showed as 36 % of
// all stall samples.)
constexpr int kSplitRows = 8;   // tile rows (= warps) per block

__global__ void __launch_bounds__(256) split_hwc_kernel(SlicerGeom g, const uint8_t* __restrict__ src, int pixel_bytes,
                                                        int border_mode, const __grid_constant__ BorderPixel border,
                                                        uint8_t* __restrict__ dst,
                                                        int64_t tile_begin, int64_t total_rows) {
  const int T = (int)g.tile;
  const int row_bytes = T * pixel_bytes;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kSplitRows + (threadIdx.x >> 5);
  if (row >= total_rows) return;
  // 32-bit unsigned divisions (the host checks the ranges): the emulated 64-bit ones were ~250 of the instructions a
  // thread spends on its three 16-byte chunks
  const uint32_t t = (uint32_t)row / (uint32_t)T;
  const int ty = (int)((uint32_t)row - t * (uint32_t)T);
  const uint32_t tile = (uint32_t)tile_begin + t;
  const uint32_t tyi = tile / (uint32_t)g.tiles_x;
  const int cy = (int)(tyi * (uint32_t)g.step), cx = (int)((tile - tyi * (uint32_t)g.tiles_x) * (uint32_t)g.step);
  const int py = cy + ty - (int)g.margin_top;
  const int H = (int)g.image_h, Wd = (int)g.image_w, ml = (int)g.margin_left;
  const bool row_inside = py >= 0 && py < H;
  const int sy = (int)reflect101(py, H);
  const uint8_t* src_row = src + (int64_t)sy * Wd * pixel_bytes;
  const uint8_t* src_end = src + (int64_t)H * Wd * pixel_bytes;
  uint8_t* dst_row = dst + row * row_bytes;
  // pixels tx in [x_lo, x_hi) read source column cx + tx - ml inside the image: one contiguous run of bytes
  int x_lo = max(0, ml - cx), x_hi = min(T, Wd + ml - cx);
  if (border_mode == 1 && !row_inside) x_hi = x_lo;          // the whole row is constant border
  if (x_hi < x_lo) x_hi = x_lo;
  const int b_lo = x_lo * pixel_bytes, b_hi = x_hi * pixel_bytes;
  const uint8_t* s0 = src_row + (int64_t)(cx - ml) * pixel_bytes;   // source byte of destination byte b (inside the run) = s0[b]
  const uintptr_t d0 = reinterpret_cast<uintptr_t>(dst_row);
  int v_lo = (int)(((d0 + b_lo + 15) & ~uintptr_t(15)) - d0), v_hi = (int)(((d0 + b_hi) & ~uintptr_t(15)) - d0);
  if (v_hi <= v_lo) v_lo = v_hi = 0;
  // four 16-byte chunks per lane and trip: all loads are issued before the first store (bytes in flight, not
  // instructions, are what a copy kernel needs)
  for (int b0 = v_lo + lane * 16; b0 < v_hi; b0 += 4 * 32 * 16) {
    uint4 o[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int b = b0 + u * 32 * 16;
      if (b >= v_hi) break;
      const uint8_t* sp = s0 + b;
      const uintptr_t sa = reinterpret_cast<uintptr_t>(sp);
      const uint4* vp = reinterpret_cast<const uint4*>(sa & ~uintptr_t(15));
      const uint32_t off = (uint32_t)(sa & 15);
      if (off == 0) {
        o[u] = __ldg(vp);
      } else if (reinterpret_cast<const uint8_t*>(vp + 2) <= src_end) {
        // two aligned 16-byte loads cover the 16 wanted bytes; pick the 5 words around them and funnel-shift
        const uint4 v0 = __ldg(vp), v1 = __ldg(vp + 1);
        const uint32_t sh = (off & 3) * 8;
        uint32_t w0, w1, w2, w3, w4;
        switch (off >> 2) {
          case 0: w0 = v0.x; w1 = v0.y; w2 = v0.z; w3 = v0.w; w4 = v1.x; break;
          case 1: w0 = v0.y; w1 = v0.z; w2 = v0.w; w3 = v1.x; w4 = v1.y; break;
          case 2: w0 = v0.z; w1 = v0.w; w2 = v1.x; w3 = v1.y; w4 = v1.z; break;
          default: w0 = v0.w; w1 = v1.x; w2 = v1.y; w3 = v1.z; w4 = v1.w; break;
        }
        o[u] = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
                          __funnelshift_r(w3, w4, sh));
      } else {                                                  // the second vector would cross the end of the image buffer
        uint8_t by[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) by[k] = __ldg(sp + k);
        o[u] = make_uint4(by[0] | (by[1] << 8) | (by[2] << 16) | ((uint32_t)by[3] << 24), by[4] | (by[5] << 8) | (by[6] << 16) | ((uint32_t)by[7] << 24),
                          by[8] | (by[9] << 8) | (by[10] << 16) | ((uint32_t)by[11] << 24), by[12] | (by[13] << 8) | (by[14] << 16) | ((uint32_t)by[15] << 24));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int b = b0 + u * 32 * 16;
      if (b < v_hi) *reinterpret_cast<uint4*>(dst_row + b) = o[u];
    }
  }
  // everything else byte by byte through the border map: [0, v_lo) and [v_hi, row_bytes)
  const int n_head = v_lo, n_rest = n_head + (row_bytes - v_hi);
  for (int i = lane; i < n_rest; i += 32) {
    const int xb = i < n_head ? i : v_hi + (i - n_head);
    const int tx = xb / pixel_bytes;
    const int pb = xb - tx * pixel_bytes;
    const int px = cx + tx - ml;
    uint8_t v;
    if (border_mode == 1 && !(row_inside && px >= 0 && px < Wd)) v = border.bytes[pb];
    else v = __ldg(src_row + (int64_t)reflect101(px, Wd) * pixel_bytes + pb);
    dst_row[xb] = v;
  }
}

// ---------------------------------------------------------------------------------------- split_norm_u8
__device__ __forceinline__ float norm_fetch(const SlicerGeom& g, const uint8_t* __restrict__ src,
                                            const float* __restrict__ lut, int channels, int tta, int64_t cy,
                                            int64_t cx, int vi, int vj, int c) {
  int si, sj;
  d4_src(tta, vi, vj, (int)g.tile, si, sj);
  const int64_t sy = reflect101(cy + si - g.margin_top, g.image_h);
  const int64_t sx = reflect101(cx + sj - g.margin_left, g.image_w);
  const uint8_t v = __ldg(src + (sy * g.image_w + sx) * channels + c);
  return __ldg(lut + c * 256 + v);
}

__global__ void split_norm_nchw_kernel(SlicerGeom g, const uint8_t* __restrict__ src, const float* __restrict__ lut,
                                       int channels, int tta, float* __restrict__ dst, int64_t tile_begin,
                                       int64_t total) {
  const int64_t T = g.tile;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int x = (int)(r % T); r /= T;
    const int y = (int)(r % T); r /= T;
    const int c = (int)(r % channels);
    const int64_t t = r / channels;
    const int64_t tile = tile_begin + t;
    const int64_t cy = (tile / g.tiles_x) * g.step, cx = (tile % g.tiles_x) * g.step;
    dst[i] = norm_fetch(g, src, lut, channels, tta, cy, cx, y, x, c);
  }
}

// PATCH32 rows of RY tile rows per block.  Stage 1: the (RY+2) x (T+2) neighbourhood of the strip is gathered once
// (D4 view map, reflect-101, LUT, bf16) into shared memory as 4 x bf16 per pixel, zeros outside the tile (the conv
// zero padding).  Stage 2: every thread assembles 8 consecutive k of one pixel from shared memory and writes one
// 16-byte vector, so global stores are fully coalesced and every source pixel is fetched ~1.3x instead of 27x.
constexpr int kPatchRY = 8;

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// F32 = false: bf16 rows (64 bytes per pixel);  F32 = true: float rows rounded to TF32 (128 bytes per pixel), the
// first-layer operand of the TF32 ("fp32 mode") path.
template <int channels, bool F32>
__global__ void __launch_bounds__(256) split_norm_patch32_kernel(SlicerGeom g, const uint8_t* __restrict__ src,
                                                                 const float* __restrict__ lut, int tta,
                                                                 uint4* __restrict__ dst, int64_t tile_begin) {
  extern __shared__ uint4 s_raw[];                // [(RY+2)][T+2] pixels: 4 x bf16 (8 bytes) or 4 x float (16 bytes) each
  uint2* s_px = reinterpret_cast<uint2*>(s_raw);
  float4* s_pf = reinterpret_cast<float4*>(s_raw);
  __shared__ float s_lut[4 * 256];
  const int T = (int)g.tile;
  const int strips = (T + kPatchRY - 1) / kPatchRY;
  const int64_t t = blockIdx.x / strips;
  const int y0 = (int)(blockIdx.x % strips) * kPatchRY;
  const int64_t tile = tile_begin + t;
  const int64_t cy = (tile / g.tiles_x) * g.step, cx = (tile % g.tiles_x) * g.step;
  for (int i = threadIdx.x; i < channels * 256; i += blockDim.x) s_lut[i] = lut[i];
  // Source byte offset of neighbourhood pixel (hy, hx) = s_row[hy] + s_col[hx]: in every D4 view the source row depends on
  // only one of the two view coordinates and the source column on the other, so the view map, the reflect-101 border
  // and the 64-bit index arithmetic are evaluated once per row / column of the strip instead of once per pixel.
  // -1 = outside the tile (the convolution's zero padding).
  const int PW = T + 2;
  __shared__ int s_row[kPatchRY + 2];
  int* s_col = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(s_raw) + (size_t)(kPatchRY + 2) * PW * (F32 ? 16 : 8));   // [PW]
  const bool transposed = (tta & 1) != 0;                          // views 1, 3, 5, 7 swap the roles of the coordinates
  for (int i = threadIdx.x; i < kPatchRY + 2 + PW; i += blockDim.x) {
    const bool is_row = i < kPatchRY + 2;
    const int v = is_row ? y0 + i - 1 : i - (kPatchRY + 2) - 1;    // view coordinate vy (rows) or vx (columns)
    int off = -1;
    if (v >= 0 && v < T) {
      int si, sj;
      d4_src(tta, is_row ? v : 0, is_row ? 0 : v, T, si, sj);      // only the coordinate that depends on v is used below
      // rows of a plain view / columns of a transposed view select the source ROW, the others the source COLUMN
      if (is_row != transposed) off = (int)(reflect101(cy + si - g.margin_top, g.image_h) * g.image_w * channels);
      else off = (int)(reflect101(cx + sj - g.margin_left, g.image_w) * channels);
    }
    if (is_row) s_row[i] = off; else s_col[i - (kPatchRY + 2)] = off;
  }
  __syncthreads();
  for (int hy = 0; hy < kPatchRY + 2; ++hy) {
    const int rp = s_row[hy];
    for (int hx = threadIdx.x; hx < PW; hx += blockDim.x) {
      const int cp = s_col[hx];
      float f[4] = {0.f, 0.f, 0.f, 0.f};
      if ((rp | cp) >= 0) {
        const uint8_t* px = src + rp + cp;
#pragma unroll
        for (int c = 0; c < channels; ++c) f[c] = s_lut[c * 256 + __ldg(px + c)];
      }
      const int i = hy * PW + hx;
      if (F32) {
        s_pf[i] = make_float4(to_tf32(f[0]), to_tf32(f[1]), to_tf32(f[2]), to_tf32(f[3]));
      } else {
        __nv_bfloat162 lo = __floats2bfloat162_rn(f[0], f[1]), hi = __floats2bfloat162_rn(f[2], f[3]);
        s_px[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      }
    }
  }
  __syncthreads();
  const int rows = min(kPatchRY, T - y0);
  if (F32) {
    // one thread = one pixel: 32 floats = eight 16-byte stores (k = tap * channels + c, zero beyond 9 * channels)
    const float* s_el = reinterpret_cast<const float*>(s_raw);
    for (int pix = threadIdx.x; pix < rows * T; pix += blockDim.x) {
      const int y = pix / T, x = pix - y * T;
      float el[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const int tap = k / channels, c = k - tap * channels;
        el[k] = tap < 9 ? s_el[((y + tap / 3) * PW + x + tap % 3) * 4 + c] : 0.f;
      }
      uint4* o = dst + ((t * T + y0 + y) * (int64_t)T + x) * 8;
#pragma unroll
      for (int qd = 0; qd < 8; ++qd)
        o[qd] = make_uint4(__float_as_uint(el[qd * 4]), __float_as_uint(el[qd * 4 + 1]), __float_as_uint(el[qd * 4 + 2]),
                           __float_as_uint(el[qd * 4 + 3]));
    }
    return;
  }
  if (channels == 3) {
    // one thread = one pixel: 9 neighbour loads (8 bytes each), register shuffles, four 16-byte stores (64 contiguous bytes)
    for (int pix = threadIdx.x; pix < rows * T; pix += blockDim.x) {
      const int y = pix / T, x = pix - y * T;
      unsigned short el[32];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const uint2 v = s_px[(y + tap / 3) * PW + x + tap % 3];
        el[tap * 3 + 0] = (unsigned short)(v.x & 0xffffu);
        el[tap * 3 + 1] = (unsigned short)(v.x >> 16);
        el[tap * 3 + 2] = (unsigned short)(v.y & 0xffffu);
      }
#pragma unroll
      for (int k = 27; k < 32; ++k) el[k] = 0;
      uint4* o = dst + ((t * T + y0 + y) * (int64_t)T + x) * 4;
#pragma unroll
      for (int qd = 0; qd < 4; ++qd)
        o[qd] = make_uint4(el[qd * 8 + 0] | ((uint32_t)el[qd * 8 + 1] << 16), el[qd * 8 + 2] | ((uint32_t)el[qd * 8 + 3] << 16),
                           el[qd * 8 + 4] | ((uint32_t)el[qd * 8 + 5] << 16), el[qd * 8 + 6] | ((uint32_t)el[qd * 8 + 7] << 16));
    }
    return;
  }
  const unsigned short* s_el = reinterpret_cast<const unsigned short*>(s_px);
  for (int i = threadIdx.x; i < rows * T * 4; i += blockDim.x) {
    const int qd = i & 3;
    const int pix = i >> 2;
    const int y = pix / T, x = pix - y * T;
    unsigned short e[8];
#pragma unroll
    for (int k8 = 0; k8 < 8; ++k8) {
      const int k = qd * 8 + k8;
      const int tap = k / channels, c = k - tap * channels;
      e[k8] = tap < 9 ? s_el[((y + tap / 3) * PW + x + tap % 3) * 4 + c] : (unsigned short)0;
    }
    dst[((t * T + y0 + y) * (int64_t)T + x) * 4 + qd] =
        make_uint4(e[0] | ((uint32_t)e[1] << 16), e[2] | ((uint32_t)e[3] << 16), e[4] | ((uint32_t)e[5] << 16),
                   e[6] | ((uint32_t)e[7] << 16));
  }
}

// SNB_LAYOUT_NHWC3_BF16: the normalised tile as packed NHWC bf16 with 3 channels (6 bytes per pixel: exactly the algorithmic
// output size of SURVEY 8d).  A block writes kNhwcRows rows of one tile; the D4 view map, the reflect-101 border and the
// 64-bit index arithmetic are evaluated once per row / column (tables in shared memory), a thread converts 8 consecutive
// pixels (24 bytes in, LUT already rounded to bf16) and writes 48 contiguous bytes as three 16-byte stores.
constexpr int kNhwcRows = 32;

__global__ void __launch_bounds__(256) split_norm_nhwc3_kernel(SlicerGeom g, const uint8_t* __restrict__ src,
                                                               const float* __restrict__ lut, int tta,
                                                               uint4* __restrict__ dst, int64_t tile_begin) {
  extern __shared__ int s_tab[];                   // [kNhwcRows] row offsets, [T] column offsets
  __shared__ unsigned short s_lut[3 * 256];
  const int T = (int)g.tile;
  const int strips = (T + kNhwcRows - 1) / kNhwcRows;
  const int64_t t = blockIdx.x / strips;
  const int y0 = (int)(blockIdx.x % strips) * kNhwcRows;
  const int64_t tile = tile_begin + t;
  const int64_t cy = (tile / g.tiles_x) * g.step, cx = (tile % g.tiles_x) * g.step;
  // NormalizeImage is affine per channel, so bf16(fma(v, a_c, b_c)) normally reproduces the bf16-rounded LUT entry for all
  // 256 levels; the block checks that exhaustively (768 comparisons) and only then takes the arithmetic path -- the
  // shared-memory gather of 3 LUT entries per pixel (random banks) is what bounds the table path
  float la[3], lb[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    lb[c] = __ldg(lut + c * 256);
    la[c] = (__ldg(lut + c * 256 + 255) - lb[c]) * (1.f / 255.f);
  }
  int ok = 1;
  for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) {
    const __nv_bfloat16 b = __float2bfloat16(lut[i]);
    const unsigned short bits = *reinterpret_cast<const unsigned short*>(&b);
    s_lut[i] = bits;
    const int c = i >> 8;
    const __nv_bfloat16 a = __float2bfloat16(fmaf((float)(i & 255), c == 0 ? la[0] : (c == 1 ? la[1] : la[2]),
                                                  c == 0 ? lb[0] : (c == 1 ? lb[1] : lb[2])));
    ok &= (*reinterpret_cast<const unsigned short*>(&a) == bits) ? 1 : 0;
  }
  const bool arith = __syncthreads_and(ok) != 0;
  const uint8_t* src_end = src + g.image_h * g.image_w * 3;
  int* s_row = s_tab;
  int* s_col = s_tab + kNhwcRows;
  const bool transposed = (tta & 1) != 0;
  for (int i = threadIdx.x; i < kNhwcRows + T; i += blockDim.x) {
    const bool is_row = i < kNhwcRows;
    const int v = is_row ? y0 + i : i - kNhwcRows;
    int off = 0;
    if (v < T) {
      int si, sj;
      d4_src(tta, is_row ? v : 0, is_row ? 0 : v, T, si, sj);
      if (is_row != transposed) off = (int)(reflect101(cy + si - g.margin_top, g.image_h) * g.image_w * 3);
      else off = (int)(reflect101(cx + sj - g.margin_left, g.image_w) * 3);
    }
    if (is_row) s_row[i] = off; else s_col[i - kNhwcRows] = off;
  }
  __syncthreads();
  const int groups = T / 8;                        // 8-pixel groups per row
  const int rows = min(kNhwcRows, T - y0);
  for (int i = threadIdx.x; i < rows * groups; i += blockDim.x) {
    const int y = i / groups, x0 = (i - y * groups) * 8;
    const int rp = s_row[y];
    unsigned short e[24];
    uint8_t lv[24];
    const int c0 = s_col[x0];
    const uint8_t* sp = src + rp + c0;
    const uintptr_t sa = reinterpret_cast<uintptr_t>(sp);
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(sa & ~uintptr_t(3));
    if (s_col[x0 + 7] - c0 == 21 && reinterpret_cast<const uint8_t*>(wp + 7) <= src_end) {
      // 8 consecutive source pixels (no reflection inside the group): 24 contiguous bytes through 7 aligned words
      const uint32_t sh = (uint32_t)(sa & 3) * 8;
      uint32_t w[7];
#pragma unroll
      for (int k = 0; k < 7; ++k) w[k] = __ldg(wp + k);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const uint32_t b4 = __funnelshift_r(w[k], w[k + 1], sh);
        lv[4 * k + 0] = (uint8_t)(b4 & 255u); lv[4 * k + 1] = (uint8_t)((b4 >> 8) & 255u);
        lv[4 * k + 2] = (uint8_t)((b4 >> 16) & 255u); lv[4 * k + 3] = (uint8_t)(b4 >> 24);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint8_t* px = src + rp + s_col[x0 + k];
        lv[3 * k + 0] = __ldg(px); lv[3 * k + 1] = __ldg(px + 1); lv[3 * k + 2] = __ldg(px + 2);
      }
    }
    if (arith) {
#pragma unroll
      for (int k = 0; k < 24; ++k) {
        const int c = k % 3;
        // (float) level on the FMA pipe: 0x4B000000 | level is the float 2^23 + level (24 I2F per thread and trip were
        // queueing on the 16-per-clock conversion unit)
        const float fl = __uint_as_float(0x4B000000u | (uint32_t)lv[k]) - 8388608.f;
        const __nv_bfloat16 v = __float2bfloat16(fmaf(fl, la[c], lb[c]));
        e[k] = *reinterpret_cast<const unsigned short*>(&v);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 24; ++k) e[k] = s_lut[(k % 3) * 256 + lv[k]];
    }
    uint4* o = dst + (((t * T + y0 + y) * (int64_t)T + x0) * 3) / 8;     // 8 pixels x 6 bytes = 48 bytes = 3 vectors
#pragma unroll
    for (int q = 0; q < 3; ++q)
      o[q] = make_uint4(e[8 * q + 0] | ((uint32_t)e[8 * q + 1] << 16), e[8 * q + 2] | ((uint32_t)e[8 * q + 3] << 16),
                        e[8 * q + 4] | ((uint32_t)e[8 * q + 5] << 16), e[8 * q + 6] | ((uint32_t)e[8 * q + 7] << 16));
  }
}

// float [n][3][h][w] -> packed NHWC bf16 [n][h][w][3] (the nn.Module.forward entry of the same first-layer kernel)
__global__ void nchw_to_nhwc3_kernel(const float* __restrict__ src, int H, int W, uint4* __restrict__ dst, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t groups = W / 8;
    const int x0 = (int)(i % groups) * 8;
    int64_t r = i / groups;
    const int y = (int)(r % H);
    const int64_t n = r / H;
    const float* p0 = src + ((n * 3) * (int64_t)H + y) * W + x0;
    const int64_t plane = (int64_t)H * W;
    unsigned short e[24];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p0 + c * plane)), b = __ldg(reinterpret_cast<const float4*>(p0 + c * plane) + 1);
      const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const __nv_bfloat16 v = __float2bfloat16(f[k]);
        e[3 * k + c] = *reinterpret_cast<const unsigned short*>(&v);
      }
    }
#pragma unroll
    for (int q = 0; q < 3; ++q)
      dst[i * 3 + q] = make_uint4(e[8 * q + 0] | ((uint32_t)e[8 * q + 1] << 16), e[8 * q + 2] | ((uint32_t)e[8 * q + 3] << 16),
                                  e[8 * q + 4] | ((uint32_t)e[8 * q + 5] << 16), e[8 * q + 6] | ((uint32_t)e[8 * q + 7] << 16));
  }
}

// one thread = one 16-byte vector of a pixel's patch row: 8 bf16 (4 vectors per pixel) or 4 TF32-rounded floats (8)
template <bool F32>
__global__ void nchw_to_patch32_kernel(const float* __restrict__ src, int channels, int H, int W,
                                       uint4* __restrict__ dst, int64_t total) {
  constexpr int VPP = F32 ? 8 : 4;     // vectors per pixel
  constexpr int EPV = F32 ? 4 : 8;     // elements per vector
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int qd = (int)(i % VPP);
    int64_t r = i / VPP;
    const int x = (int)(r % W); r /= W;
    const int y = (int)(r % H);
    const int64_t n = r / H;
    const float* img = src + n * channels * (int64_t)H * W;
    float f[8];
#pragma unroll
    for (int e = 0; e < EPV; ++e) {
      const int k = qd * EPV + e;
      const int tap = k / channels, c = k - tap * channels;
      const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
      f[e] = (tap < 9 && yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(img + ((int64_t)c * H + yy) * W + xx) : 0.f;
    }
    if (F32) {
      dst[i] = make_uint4(__float_as_uint(to_tf32(f[0])), __float_as_uint(to_tf32(f[1])), __float_as_uint(to_tf32(f[2])),
                          __float_as_uint(to_tf32(f[3])));
    } else {
      __nv_bfloat162 p0 = __floats2bfloat162_rn(f[0], f[1]), p1 = __floats2bfloat162_rn(f[2], f[3]);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(f[4], f[5]), p3 = __floats2bfloat162_rn(f[6], f[7]);
      dst[i] = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                          *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
    }
  }
}

// --------------------------------------------------------------------------------------------------- merge
template <typename T>
__device__ __forceinline__ double load_as_double(const void* p, int64_t i) {
  return (double)static_cast<const T*>(p)[i];
}

// value of tile `tile` at original-frame position (i, j), channel c
template <int TILE_DT, int TTA>
__device__ __forceinline__ double tile_value(const void* __restrict__ tiles, int64_t tile, int i, int j, int c, int T,
                                             int C) {
  if (TTA == 1) {
    const int64_t idx = ((tile * T + i) * T + j) * C + c;
    if (TILE_DT == SNB_DT_U8) return load_as_double<uint8_t>(tiles, idx);
    if (TILE_DT == SNB_DT_F64) return load_as_double<double>(tiles, idx);
    return load_as_double<float>(tiles, idx);
  } else {
    // tta_d4_deaug: float32 sum of the 8 inverse-transformed views in listed order, then * 0.125f
    const float* tp = static_cast<const float*>(tiles);
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      int a, b;
      d4_dst(v, i, j, T, a, b);
      const float x = tp[(((tile * 8 + v) * T + a) * T + b) * C + c];
      s = v == 0 ? x : __fadd_rn(s, x);
    }
    return (double)__fmul_rn(s, 0.125f);
  }
}

// blockIdx.y = image row (no 64-bit division anywhere); a thread owns VEC consecutive elements of the row so the
// crop-range arithmetic is shared and stores are 16-byte (float4) / 4-byte (uchar4) vectors.  The accumulation per
// element is untouched: float64, crop order, no FMA contraction -> bit-exact against numpy.
template <int TILE_DT, int TTA, int VEC>
__global__ void __launch_bounds__(256) merge_kernel(SlicerGeom g, const void* __restrict__ tiles, int C,
                                                    const double* __restrict__ weight, void* __restrict__ out,
                                                    int out_dtype, uint8_t* __restrict__ mask, float thr, int row0) {
  const int T = (int)g.tile, S = (int)g.step;
  const int WC = (int)g.image_w * C;
  const int tiles_x = (int)g.tiles_x, tiles_y = (int)g.tiles_y;
  const int y_img = blockIdx.y + row0;                       // image row (row0 = first row of the band being merged)
  const int Y = y_img + (int)g.margin_top;                   // padded-canvas row
  // crops covering row Y: iy*S <= Y < iy*S + T
  const int iy0 = Y - T + 1 <= 0 ? 0 : (Y - T + S) / S;
  const int iy1 = min(Y / S, tiles_y - 1);
  for (int xc0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC; xc0 < WC; xc0 += gridDim.x * blockDim.x * VEC) {
    float qf[VEC];
    double qd[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const int xc = xc0 + e;
      const int x = C == 1 ? xc : xc / C;
      const int c = C == 1 ? 0 : xc - x * C;
      const int X = x + (int)g.margin_left;
      const int ix0 = X - T + 1 <= 0 ? 0 : (X - T + S) / S;
      const int ix1 = min(X / S, tiles_x - 1);
      double acc = 0.0, norm = 0.0;
      if (xc < WC) {
        for (int iy = iy0; iy <= iy1; ++iy) {          // crop order: y outer, x inner (lib/tiles.py:94-96,150)
          const int ty = Y - iy * S;
          for (int ix = ix0; ix <= ix1; ++ix) {
            const int tx = X - ix * S;
            const double w = __ldg(weight + ty * T + tx);
            const double v = tile_value<TILE_DT, TTA>(tiles, (int64_t)iy * tiles_x + ix, ty, tx, c, T, C);
            acc = __dadd_rn(acc, __dmul_rn(v, w));     // no FMA contraction: numpy rounds the product first
            norm = __dadd_rn(norm, w);
          }
        }
      }
      norm = norm < DBL_EPSILON ? DBL_EPSILON : norm;  // np.clip(norm, eps, None)
      qd[e] = __ddiv_rn(acc, norm);
      qf[e] = __double2float_rn(qd[e]);
    }
    const int64_t i = (int64_t)y_img * WC + xc0;
    if (VEC == 4 && xc0 + 3 < WC) {   // WC % 4 == 0 and 16-byte aligned rows are guaranteed by the launcher for VEC == 4
      if (out) {
        if (out_dtype == SNB_DT_F32) *reinterpret_cast<float4*>(static_cast<float*>(out) + i) = make_float4(qf[0], qf[1], qf[2], qf[3]);
        else if (out_dtype == SNB_DT_F64) {
          double* o = static_cast<double*>(out) + i;
          *reinterpret_cast<double2*>(o) = make_double2(qd[0], qd[1]);
          *reinterpret_cast<double2*>(o + 2) = make_double2(qd[2], qd[3]);
        } else {
          *reinterpret_cast<uchar4*>(static_cast<uint8_t*>(out) + i) =
              make_uchar4((uint8_t)(int)qd[0], (uint8_t)(int)qd[1], (uint8_t)(int)qd[2], (uint8_t)(int)qd[3]);
        }
      }
      if (mask)
        *reinterpret_cast<uchar4*>(mask + i) = make_uchar4(qf[0] > thr ? 255 : 0, qf[1] > thr ? 255 : 0,
                                                           qf[2] > thr ? 255 : 0, qf[3] > thr ? 255 : 0);
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        if (xc0 + e >= WC) break;
        if (out) {
          if (out_dtype == SNB_DT_F32) static_cast<float*>(out)[i + e] = qf[e];
          else if (out_dtype == SNB_DT_F64) static_cast<double*>(out)[i + e] = qd[e];
          else static_cast<uint8_t*>(out)[i + e] = (uint8_t)(int)qd[e];  // astype(uint8) truncates
        }
        if (mask) mask[i + e] = qf[e] > thr ? 255 : 0;
      }
    }
  }
}

// Fast path of the merge for the inference layout (float32 probabilities, one channel, no TTA) when the geometry is
// 4-aligned (tile, step, left margin and width multiples of 4): a thread owns 4 consecutive pixels of one row, which
// then share their covering crops, so a crop contributes one float4 tile load and two double2 weight loads.  The
// per-pixel arithmetic (float64, crop order, rounded product, IEEE divide) is identical to the generic kernel.
__global__ void __launch_bounds__(256) merge_f32c1_vec4_kernel(SlicerGeom g, const float* __restrict__ tiles,
                                                               const double* __restrict__ weight,
                                                               float* __restrict__ out, uint8_t* __restrict__ mask,
                                                               float thr, int row0) {
  const int T = (int)g.tile, S = (int)g.step, W = (int)g.image_w;
  const int tiles_x = (int)g.tiles_x, tiles_y = (int)g.tiles_y;
  const int y_img = blockIdx.y + row0;
  const int Y = y_img + (int)g.margin_top;
  const int iy0 = Y - T + 1 <= 0 ? 0 : (Y - T + S) / S;
  const int iy1 = min(Y / S, tiles_y - 1);
  for (int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4; x0 < W; x0 += gridDim.x * blockDim.x * 4) {
    const int X = x0 + (int)g.margin_left;                  // multiple of 4; crop edges are multiples of 4 too
    const int ix0 = X - T + 1 <= 0 ? 0 : (X - T + S) / S;   // same for X .. X+3
    const int ix1 = min(X / S, tiles_x - 1);
    double acc[4] = {0.0, 0.0, 0.0, 0.0}, norm[4] = {0.0, 0.0, 0.0, 0.0};
    // walk the covering crops with pointer increments: next crop in x is one tile further and S pixels to the left,
    // next crop row is tiles_x tiles further and S rows up
    const int ty0 = Y - iy0 * S, tx0 = X - ix0 * S;
    const float* trow = tiles + (((int64_t)iy0 * tiles_x + ix0) * T + ty0) * T + tx0;
    const double* wrow = weight + ty0 * T + tx0;
    const int64_t t_dx = (int64_t)T * T - S, t_dy = (int64_t)tiles_x * T * T - (int64_t)S * T;
    const int w_dy = -S * T;
    const int nx = ix1 - ix0, ny = iy1 - iy0;
#pragma unroll 1
    for (int jy = 0; jy <= ny; ++jy, trow += t_dy, wrow += w_dy) {
      const float* tp = trow;
      const double* wp = wrow;
#pragma unroll 1
      for (int jx = 0; jx <= nx; ++jx, tp += t_dx, wp -= S) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(tp));
        const double2 w01 = __ldg(reinterpret_cast<const double2*>(wp));
        const double2 w23 = __ldg(reinterpret_cast<const double2*>(wp + 2));
        acc[0] = __dadd_rn(acc[0], __dmul_rn((double)v.x, w01.x)); norm[0] = __dadd_rn(norm[0], w01.x);
        acc[1] = __dadd_rn(acc[1], __dmul_rn((double)v.y, w01.y)); norm[1] = __dadd_rn(norm[1], w01.y);
        acc[2] = __dadd_rn(acc[2], __dmul_rn((double)v.z, w23.x)); norm[2] = __dadd_rn(norm[2], w23.x);
        acc[3] = __dadd_rn(acc[3], __dmul_rn((double)v.w, w23.y)); norm[3] = __dadd_rn(norm[3], w23.y);
      }
    }
    float q[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const double nrm = norm[e] < DBL_EPSILON ? DBL_EPSILON : norm[e];
      q[e] = __double2float_rn(__ddiv_rn(acc[e], nrm));
    }
    const int64_t i = (int64_t)y_img * W + x0;
    if (out) *reinterpret_cast<float4*>(out + i) = make_float4(q[0], q[1], q[2], q[3]);
    if (mask)
      *reinterpret_cast<uchar4*>(mask + i) =
          make_uchar4(q[0] > thr ? 255 : 0, q[1] > thr ? 255 : 0, q[2] > thr ? 255 : 0, q[3] > thr ? 255 : 0);
  }
}

// ---- periodic merge (float32 probabilities, one channel, tile <= 2 * step) ----------------------------------------
// Canvas pixel X = kx*step + r is covered by crop kx at tile column r and, when r < tile - step, by crop kx-1 at column
// r + step: the weights a pixel needs depend only on its residues (X mod step, Y mod step).  A thread owns PX
// consecutive residues of one image row, reads its <= 4*PX float64 weights ONCE, precomputes the loop-invariant norm and
// its correctly rounded reciprocal, and walks the periods of the row: per pixel only the tile values move.
//   * single-cover pixels: q = RN32(RN64(RN64(v*w) / w)) = v exactly (the float64 round trip perturbs v by < 2^-52
//     relative, far inside the float32 rounding interval), so they are copied when eps <= w < 2^60;
//   * a / b with b invariant: y = RN(1/b) (__drcp_rn), q0 = RN(a*y), then two FMA corrections.  q1 is within half an ulp
//     + 2^-104 of a/b; by Markstein's theorem (y correctly rounded, q1 faithful, r1 = a - b*q1 exact) q2 = RN(q1 + r1*y)
//     IS the correctly rounded quotient, i.e. bit-equal to numpy's float64 division (tests/test_host_logic.py checks
//     the sequence against exact rational arithmetic).  Valid while nothing under/overflows: b is kept in
//     [eps, 2^60) and a float32 result in [2^-100, FLT_MAX] proves a was in range; anything else (except a == 0, for
//     which the sequence returns +0 exactly) takes the IEEE division;
//   * periods at the canvas edge (a crop of the pattern does not exist) and exotic weights take a slow, obviously
//     correct routine.
// Accumulation order (crop order: y outer, x inner), rounded products and the quotient are those of
// lib/tiles.py:146-161, hence bit-exact.
//
// Operands are staged in shared memory by 1-D bulk copies (cp.async.bulk -> mbarrier): a row needs the tile row `ty` of
// every crop in its 1-2 covering crop rows (tiles_x * T floats each, contiguous per crop) and the matching weight rows
// (T doubles).  Bytes in flight per SM are then set by the staged rows (30-60 KB each), not by registers, and every
// tile element is read from HBM exactly once, every output byte written once (algorithmic traffic).
__device__ __forceinline__ float div_by_invariant_fast(double a, double b, double y) {
  const double q0 = __dmul_rn(a, y);
  const double r0 = __fma_rn(-b, q0, a);
  const double q1 = __fma_rn(r0, y, q0);
  const double r1 = __fma_rn(-b, q1, a);
  return __double2float_rn(__fma_rn(r1, y, q1));
}
// the fast quotient is proven when its float32 value is in [2^-100, FLT_MAX] (a was in range) or a == 0 (it returned +0)
__device__ __forceinline__ bool div_by_invariant_ok(float f, double a) {
  return (fabsf(f) >= 0x1p-100f && fabsf(f) <= FLT_MAX) || a == 0.0;
}

template <int PX>
__device__ __forceinline__ void lds_px(const float* p, float (&v)[PX]) {
  if constexpr (PX == 4) { const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  else if constexpr (PX == 2) { const float2 t = *reinterpret_cast<const float2*>(p); v[0] = t.x; v[1] = t.y; }
  else v[0] = *p;
}

template <int PX>
__device__ __forceinline__ void lds_w(const double* p, double (&w)[PX]) {
  if constexpr (PX >= 2) {
#pragma unroll
    for (int e = 0; e < PX; e += 2) {
      const double2 t = *reinterpret_cast<const double2*>(p + e);
      w[e] = t.x; w[e + 1] = t.y;
    }
  } else w[0] = *p;
}

template <int PX>
__device__ __forceinline__ void ldg_w(const double* __restrict__ p, double (&w)[PX]) {
  if constexpr (PX >= 2) {
#pragma unroll
    for (int e = 0; e < PX; e += 2) {
      const double2 t = __ldg(reinterpret_cast<const double2*>(p + e));
      w[e] = t.x; w[e + 1] = t.y;
    }
  } else w[0] = __ldg(p);
}

// acc += v * w in the reference's arithmetic (rounded product, then rounded sum)
template <int PX>
__device__ __forceinline__ void accp(const float (&v)[PX], const double (&w)[PX], double (&acc)[PX]) {
#pragma unroll
  for (int e = 0; e < PX; ++e) acc[e] = __dadd_rn(acc[e], __dmul_rn((double)v[e], w[e]));
}

template <int PX, bool HAS_OUT, bool HAS_MASK>
__device__ __forceinline__ void store_px(float* o, uint8_t* m, const float (&q)[PX], float thr) {
  if constexpr (HAS_OUT) {
    if constexpr (PX == 4) *reinterpret_cast<float4*>(o) = make_float4(q[0], q[1], q[2], q[3]);
    else if constexpr (PX == 2) *reinterpret_cast<float2*>(o) = make_float2(q[0], q[1]);
    else *o = q[0];
  }
  if constexpr (HAS_MASK) {
    if constexpr (PX == 4)
      *reinterpret_cast<uchar4*>(m) = make_uchar4(q[0] > thr ? 255 : 0, q[1] > thr ? 255 : 0, q[2] > thr ? 255 : 0, q[3] > thr ? 255 : 0);
    else if constexpr (PX == 2) *reinterpret_cast<uchar2*>(m) = make_uchar2(q[0] > thr ? 255 : 0, q[1] > thr ? 255 : 0);
    else *m = q[0] > thr ? 255 : 0;
  }
}

// One period, any crop pattern, any weights: norm on the fly, IEEE division.  Self-contained (re-reads its weights from
// global memory) and not inlined: it runs for the 1-2 edge periods of a row and for exotic weights only.
template <int PX, bool HAS_OUT, bool HAS_MASK>
__device__ __noinline__ void merge_slow_period(const double* sw0, const double* sw1, const float* st0r, int row2_off, int T,
                                               int S, int tiles_x, int r, bool two_rows, int kx, float* o, uint8_t* m,
                                               float thr) {
  const bool useA = r < T - S && kx >= 1, useB = kx <= tiles_x - 1;
  float q[PX];
#pragma unroll
  for (int e = 0; e < PX; ++e) {
    double acc = 0.0, n = 0.0;
    for (int row = 0; row < (two_rows ? 2 : 1); ++row) {
      const double* sw = row ? sw1 : sw0;
      const float* st = st0r + row * row2_off + kx * T + e;
      if (useA) {
        const double w = __ldg(sw + r + S + e);
        acc = __dadd_rn(acc, __dmul_rn((double)st[S - T], w));
        n = __dadd_rn(n, w);
      }
      if (useB) {
        const double w = __ldg(sw + r + e);
        acc = __dadd_rn(acc, __dmul_rn((double)st[0], w));
        n = __dadd_rn(n, w);
      }
    }
    q[e] = __double2float_rn(__ddiv_rn(acc, n < DBL_EPSILON ? DBL_EPSILON : n));   // np.clip(norm, eps, None)
  }
  store_px<PX, HAS_OUT, HAS_MASK>(o, m, q, thr);
}

// U interior periods of one thread, `pstep` floats / `ostep` pixels apart: every load is issued first, the U * PX
// accumulate-and-divide chains are independent (they interleave in the FP64 pipe), and ONE deferred branch covers the
// never-taken IEEE-division fallback, so nothing serialises the chains.
template <int PX, bool A, bool TWO, int U, bool HAS_OUT, bool HAS_MASK>
__device__ __forceinline__ void merge_periods(const float* p, int row2_off, int dA, int pstep, int ostep,
                                              const double (&wA0)[PX], const double (&wB0)[PX], const double (&wA1)[PX],
                                              const double (&wB1)[PX], const double (&nrm)[PX], const double (&rcp)[PX],
                                              float* o, uint8_t* m, float thr) {
  float vA0[U][PX], vB0[U][PX], vA1[U][PX], vB1[U][PX];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const float* pu = p + u * pstep;
    if constexpr (A) lds_px<PX>(pu + dA, vA0[u]);
    lds_px<PX>(pu, vB0[u]);
    if constexpr (TWO) {
      if constexpr (A) lds_px<PX>(pu + row2_off + dA, vA1[u]);
      lds_px<PX>(pu + row2_off, vB1[u]);
    }
  }
  double acc[U][PX];
  float q[U][PX];
  bool ok = true;
#pragma unroll
  for (int u = 0; u < U; ++u) {
#pragma unroll
    for (int e = 0; e < PX; ++e) acc[u][e] = 0.0;
    if constexpr (A) accp<PX>(vA0[u], wA0, acc[u]);
    accp<PX>(vB0[u], wB0, acc[u]);
    if constexpr (TWO) {
      if constexpr (A) accp<PX>(vA1[u], wA1, acc[u]);
      accp<PX>(vB1[u], wB1, acc[u]);
    }
#pragma unroll
    for (int e = 0; e < PX; ++e) {
      q[u][e] = div_by_invariant_fast(acc[u][e], nrm[e], rcp[e]);
      ok = ok && div_by_invariant_ok(q[u][e], acc[u][e]);
    }
  }
  if (!ok) {
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int e = 0; e < PX; ++e) q[u][e] = __double2float_rn(__ddiv_rn(acc[u][e], nrm[e]));
  }
#pragma unroll
  for (int u = 0; u < U; ++u) store_px<PX, HAS_OUT, HAS_MASK>(o + u * ostep, m + u * ostep, q[u], thr);
}

// interior periods of one thread: `n` periods, `kstep` periods apart; A = the crop to the left also covers the pixel,
// TWO = two crop rows cover the image row
template <int PX, bool A, bool TWO, bool HAS_OUT, bool HAS_MASK>
__device__ __forceinline__ void merge_interior(const float* p, int row2_off, int dA, int pstep, int ostep, int n,
                                               const double (&wA0)[PX], const double (&wB0)[PX], const double (&wA1)[PX],
                                               const double (&wB1)[PX], const double (&nrm)[PX], const double (&rcp)[PX],
                                               float* o, uint8_t* m, float thr) {
  constexpr int U = PX >= 4 ? 2 : 4;                  // 8 independent chains in flight (4 for PX == 1)
  int it = 0;
#pragma unroll 1
  for (; it + U <= n; it += U, p += U * pstep, o += U * ostep, m += U * ostep)
    merge_periods<PX, A, TWO, U, HAS_OUT, HAS_MASK>(p, row2_off, dA, pstep, ostep, wA0, wB0, wA1, wB1, nrm, rcp, o, m, thr);
#pragma unroll 1
  for (; it < n; ++it, p += pstep, o += ostep, m += ostep)
    merge_periods<PX, A, TWO, 1, HAS_OUT, HAS_MASK>(p, row2_off, dA, pstep, ostep, wA0, wB0, wA1, wB1, nrm, rcp, o, m, thr);
}

// The <= 4 * PX float64 weights of one thread for one image row, fetched from global memory (the 2 MB table is L2
// resident) BEFORE the thread waits for its staged tile rows, so the fetch latency hides behind the bulk copies and the
// weights cost no shared memory (4 instead of 3 resident CTAs per SM at 512 / 384).
template <int PX>
struct MergeWeights {
  double wA0[PX], wB0[PX], wA1[PX], wB1[PX];
  double nrm[PX], rcp[PX];          // interior norm (clipped at eps) and its correctly rounded reciprocal
  bool w_ok, wa_ok, copy, fast;     // weight of the own / left crop allows the copy shortcut; single-cover thread; fast division usable
};

template <int PX>
__device__ __forceinline__ void merge_fetch_weights(const double* __restrict__ gw0, const double* __restrict__ gw1, int r,
                                                    int S, bool hasA, bool two_rows, MergeWeights<PX>& w) {
#pragma unroll
  for (int e = 0; e < PX; ++e) w.wA0[e] = w.wA1[e] = w.wB1[e] = 0.0;
  ldg_w<PX>(gw0 + r, w.wB0);
  if (hasA) ldg_w<PX>(gw0 + r + S, w.wA0);
  if (two_rows) {
    ldg_w<PX>(gw1 + r, w.wB1);
    if (hasA) ldg_w<PX>(gw1 + r + S, w.wA1);
  }
  // single crop row: every pixel covered by ONE crop (no left neighbour, or an edge period) is a copy when the
  // weight allows it (w_ok / wa_ok: own crop / crop to the left)
  bool w_ok = true, wa_ok = true;
#pragma unroll
  for (int e = 0; e < PX; ++e) {
    w_ok = w_ok && w.wB0[e] >= DBL_EPSILON && w.wB0[e] < 0x1p60;
    wa_ok = wa_ok && w.wA0[e] >= DBL_EPSILON && w.wA0[e] < 0x1p60;
  }
  const bool copy = !hasA && !two_rows && w_ok;
  bool fast = true;
#pragma unroll
  for (int e = 0; e < PX; ++e) w.nrm[e] = w.rcp[e] = 1.0;
  if (!copy) {
#pragma unroll
    for (int e = 0; e < PX; ++e) {
      double n = 0.0;
      if (hasA) n = __dadd_rn(n, w.wA0[e]);
      n = __dadd_rn(n, w.wB0[e]);
      if (two_rows) {
        if (hasA) n = __dadd_rn(n, w.wA1[e]);
        n = __dadd_rn(n, w.wB1[e]);
      }
      n = n < DBL_EPSILON ? DBL_EPSILON : n;          // np.clip(norm, eps, None)
      fast = fast && n < 0x1p60;                      // NaN / huge weights take the slow routine
      w.nrm[e] = n;
      w.rcp[e] = __drcp_rn(n);
    }
  }
  w.w_ok = w_ok; w.wa_ok = wa_ok; w.copy = copy; w.fast = fast;
}

// All periods kx = kfirst, kfirst + kstep, ... <= kx_hi of one thread (residues r .. r+PX-1) for one staged image row.
// sw0/sw1: global weight rows of the first/second covering crop row; st0r = staged tile rows + r (crop kx at kx*T, second
// crop row row2_off floats further); out_row/mask_row = output row pointers (pixel X lives at out_row[X - ml]).
template <int PX, bool HAS_OUT, bool HAS_MASK>
__device__ __forceinline__ void merge_row(int T, int S, int tiles_x, int ml, int r, bool two_rows, const double* sw0,
                                          const double* sw1, const MergeWeights<PX>& mw, const float* st0r, int row2_off,
                                          int kfirst, int kx_hi, int kstep, float* out_row, uint8_t* mask_row, float thr) {
  if (kfirst > kx_hi) return;
  const bool hasA = r < T - S;
  const double (&wA0)[PX] = mw.wA0;
  const double (&wB0)[PX] = mw.wB0;
  const double (&wA1)[PX] = mw.wA1;
  const double (&wB1)[PX] = mw.wB1;
  const double (&nrm)[PX] = mw.nrm;
  const double (&rcp)[PX] = mw.rcp;
  const bool w_ok = mw.w_ok, wa_ok = mw.wa_ok, copy = mw.copy, fast = mw.fast;
  int kx = kfirst;
  int n = (kx_hi - kx) / kstep + 1;
  const int off0 = r - ml;                            // pixel of period kx sits at out_row[kx*S + off0]
  auto slow = [&](int k) {
    merge_slow_period<PX, HAS_OUT, HAS_MASK>(sw0, sw1, st0r, row2_off, T, S, tiles_x, r, two_rows, k,
                                             HAS_OUT ? out_row + (k * S + off0) : nullptr,
                                             HAS_MASK ? mask_row + (k * S + off0) : nullptr, thr);
  };
  // an edge period of a single-crop-row image row is covered by one crop: copy from column `col` of crop slot `k`
  auto edge = [&](int k, bool from_left, bool ok_w) {
    if (!two_rows && ok_w) {
      float q[PX];
      lds_px<PX>(st0r + k * T + (from_left ? S - T : 0), q);
      store_px<PX, HAS_OUT, HAS_MASK>(HAS_OUT ? out_row + (k * S + off0) : nullptr,
                                      HAS_MASK ? mask_row + (k * S + off0) : nullptr, q, thr);
    } else {
      slow(k);
    }
  };
  if (!fast) {
    for (; n > 0; --n, kx += kstep) slow(kx);
    return;
  }
  if (hasA && kx == 0) { edge(0, false, w_ok); kx += kstep; --n; }    // no crop to the left of the first one
  if (n > 0 && kx + (n - 1) * kstep >= tiles_x) { edge(kx + (n - 1) * kstep, true, wa_ok); --n; }   // right part of the last crop
  if (n <= 0) return;
  const float* p = st0r + kx * T;
  float* o = HAS_OUT ? out_row + (kx * S + off0) : nullptr;
  uint8_t* m = HAS_MASK ? mask_row + (kx * S + off0) : nullptr;
  const int pstep = kstep * T, ostep = kstep * S, dA = S - T;
  if (copy) {
#pragma unroll 4
    for (int it = 0; it < n; ++it, p += pstep, o += ostep, m += ostep) {
      float q[PX];
      lds_px<PX>(p, q);
      store_px<PX, HAS_OUT, HAS_MASK>(o, m, q, thr);
    }
  } else if (two_rows) {
    if (hasA) merge_interior<PX, true, true, HAS_OUT, HAS_MASK>(p, row2_off, dA, pstep, ostep, n, wA0, wB0, wA1, wB1, nrm, rcp, o, m, thr);
    else merge_interior<PX, false, true, HAS_OUT, HAS_MASK>(p, row2_off, dA, pstep, ostep, n, wA0, wB0, wA1, wB1, nrm, rcp, o, m, thr);
  } else {
    if (hasA) merge_interior<PX, true, false, HAS_OUT, HAS_MASK>(p, row2_off, dA, pstep, ostep, n, wA0, wB0, wA1, wB1, nrm, rcp, o, m, thr);
    else merge_interior<PX, false, false, HAS_OUT, HAS_MASK>(p, row2_off, dA, pstep, ostep, n, wA0, wB0, wA1, wB1, nrm, rcp, o, m, thr);
  }
}

// geometry of image row y: covering crop rows (crop order) and the tile row of the first one
struct MergeRow {
  int n_rows, iy0, ty0;
};
__device__ __forceinline__ MergeRow merge_row_geom(int y, int mt, int T, int S, int tiles_y) {
  const int Y = y + mt;
  const int ky = Y / S, ry = Y - ky * S;
  const bool up = ky >= 1 && ry < T - S;              // crop row ky-1 covers Y (first in crop order)
  MergeRow g;
  g.n_rows = (up && ky <= tiles_y - 1) ? 2 : 1;
  g.iy0 = up ? ky - 1 : ky;
  g.ty0 = Y - g.iy0 * S;
  return g;
}

// bulk copies of one row's tile rows, issued by a full warp: crops [ix0, ix0 + nx) of the 1-2 covering crop rows;
// stage = [st0[nx_max*T] | st1[nx_max*T]]
__device__ __forceinline__ void merge_stage_row(const MergeRow& mr, int T, int S, int tiles_x, int ix0, int nx, int nx_max,
                                                const float* tiles, float* st0, uint64_t* bar, int lane) {
  if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)(mr.n_rows * nx * T * 4));
  __syncwarp();
  for (int j = lane; j < mr.n_rows * nx; j += 32) {
    const int rr = j / nx, ix = j - rr * nx;
    bulk_load_1d(st0 + (rr * nx_max + ix) * T,
                 tiles + (((int64_t)(mr.iy0 + rr) * tiles_x + ix0 + ix) * T + (mr.ty0 - rr * S)) * T, (uint32_t)T * 4, bar);
  }
}

// ---- variant 1: one CTA per (image row, segment of `kp` periods), S / PX threads, several CTAs resident per SM
// threads per CTA = step / PX (<= kStagedThreads<PX>); PX = 2 is compiled for 3 resident CTAs of <= 256 threads (a 64-register cap for 4 CTAs spills and loses 20 %)
template <int PX> constexpr int kStagedThreads = PX == 1 ? 512 : 256;
template <int PX> constexpr int kStagedMinBlocks = PX == 2 ? 3 : 1;

template <int PX, bool HAS_OUT, bool HAS_MASK>
__global__ void __launch_bounds__(kStagedThreads<PX>, kStagedMinBlocks<PX>) merge_f32c1_staged_kernel(
    SlicerGeom g, const float* __restrict__ tiles, const double* __restrict__ weight, float* __restrict__ out,
    uint8_t* __restrict__ mask, float thr, int xs, int kp, int row0) {
  extern __shared__ __align__(16) uint8_t merge_smem[];
  const int T = (int)g.tile, S = (int)g.step, W = (int)g.image_w;
  const int tiles_x = (int)g.tiles_x, ml = (int)g.margin_left;
  uint64_t* bar = reinterpret_cast<uint64_t*>(merge_smem);
  float* st0 = reinterpret_cast<float*>(merge_smem + 16);
  const int yb = blockIdx.x / xs, seg = blockIdx.x - yb * xs;
  const int y = yb + row0;                                     // image row (row0 = first row of the band being merged)
  // periods [k0, k1] of this segment need crops [k0 - 1, k1] clipped to the crop grid
  const int k0 = seg * kp, k1 = min(k0 + kp - 1, tiles_x);
  const int ix0 = max(k0 - 1, 0), nx = min(k1, tiles_x - 1) - ix0 + 1, nx_max = min(kp + 1, tiles_x);
  const MergeRow mr = merge_row_geom(y, (int)g.margin_top, T, S, (int)g.tiles_y);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x < 32) merge_stage_row(mr, T, S, tiles_x, ix0, nx, nx_max, tiles, st0, bar, threadIdx.x);
  const int r = threadIdx.x * PX;
  if (r >= S) return;
  // periods with an output pixel: ml <= kx*S + r <= ml + W - PX; the last one may be the right part of the last crop
  const int kx_lo = max(ml > r ? (ml - r + S - 1) / S : 0, k0);
  int kx_hi = ml + W - PX - r >= 0 ? (ml + W - PX - r) / S : -1;
  kx_hi = min(min(kx_hi, tiles_x - 1 + (r < T - S ? 1 : 0)), k1);
  const double* gw0 = weight + mr.ty0 * T;
  const double* gw1 = weight + (mr.ty0 - S) * T;       // only dereferenced when two crop rows cover the image row
  MergeWeights<PX> mw;
  if (kx_lo <= kx_hi) merge_fetch_weights<PX>(gw0, gw1, r, S, r < T - S, mr.n_rows == 2, mw);
  mbar_wait(bar, 0);
  merge_row<PX, HAS_OUT, HAS_MASK>(T, S, tiles_x, ml, r, mr.n_rows == 2, gw0, gw1, mw, st0 + r - ix0 * T, nx_max * T, kx_lo,
                                   kx_hi, 1, out + (int64_t)y * W, mask + (int64_t)y * W, thr);
}

// ---- variant 2: persistent CTAs (one per SM) with a producer warp keeping a ring of staged rows full; 24 consumer
// warps share a staged row: consumer (kgroup, residue group) handles periods kx = kx_lo + kgroup, + KG, ...
constexpr int kMergeConsumers = 768;

template <int PX, bool HAS_OUT, bool HAS_MASK>
__global__ void __launch_bounds__(kMergeConsumers + 32, 1) merge_f32c1_ring_kernel(
    SlicerGeom g, const float* __restrict__ tiles, const double* __restrict__ weight, float* __restrict__ out,
    uint8_t* __restrict__ mask, float thr, int n_stages, int stage_bytes) {
  extern __shared__ __align__(16) uint8_t merge_smem[];
  const int T = (int)g.tile, S = (int)g.step, W = (int)g.image_w, H = (int)g.image_h;
  const int tiles_x = (int)g.tiles_x, tiles_y = (int)g.tiles_y, ml = (int)g.margin_left, mt = (int)g.margin_top;
  uint64_t* full = reinterpret_cast<uint64_t*>(merge_smem);      // [8]
  uint64_t* empty = full + 8;                                    // [8]
  uint8_t* stage0 = merge_smem + 128;
  if (threadIdx.x == 0) {
    for (int i = 0; i < n_stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], kMergeConsumers / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {
    int s = 0;
    uint32_t par = 1;                                  // waiting on parity 1 of a fresh barrier returns immediately
    for (int y = blockIdx.x; y < H; y += gridDim.x) {
      const MergeRow mr = merge_row_geom(y, mt, T, S, tiles_y);
      mbar_wait(&empty[s], par);
      merge_stage_row(mr, T, S, tiles_x, 0, tiles_x, tiles_x, tiles,
                      reinterpret_cast<float*>(stage0 + (size_t)s * stage_bytes), &full[s], lane);
      if (++s == n_stages) { s = 0; par ^= 1; }
    }
    return;
  }
  const int ctid = threadIdx.x - 32;
  const int RG = S / PX;                               // residue groups per period
  const int KG = kMergeConsumers / RG;                 // period groups
  const int kgroup = ctid / RG;
  const int r = (ctid - kgroup * RG) * PX;
  const int kx_lo = ml > r ? (ml - r + S - 1) / S : 0;
  int kx_hi = ml + W - PX - r >= 0 ? (ml + W - PX - r) / S : -1;
  kx_hi = min(kx_hi, tiles_x - 1 + (r < T - S ? 1 : 0));
  if (kgroup >= KG) kx_hi = -1;                        // spare threads only take part in the barriers
  int s = 0;
  uint32_t par = 0;
  for (int y = blockIdx.x; y < H; y += gridDim.x) {
    const MergeRow mr = merge_row_geom(y, mt, T, S, tiles_y);
    const float* st0 = reinterpret_cast<const float*>(stage0 + (size_t)s * stage_bytes);
    const double* gw0 = weight + mr.ty0 * T;
    const double* gw1 = weight + (mr.ty0 - S) * T;
    MergeWeights<PX> mw;
    if (kx_lo + kgroup <= kx_hi) merge_fetch_weights<PX>(gw0, gw1, r, S, r < T - S, mr.n_rows == 2, mw);
    mbar_wait(&full[s], par);
    merge_row<PX, HAS_OUT, HAS_MASK>(T, S, tiles_x, ml, r, mr.n_rows == 2, gw0, gw1, mw, st0 + r, tiles_x * T,
                                     kx_lo + kgroup, kx_hi, KG, out + (int64_t)y * W, mask + (int64_t)y * W, thr);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);              // this warp is done reading the stage
    if (++s == n_stages) { s = 0; par ^= 1; }
  }
}

template <int PX, bool HAS_OUT, bool HAS_MASK>
static bool launch_merge_ring(const SlicerGeom& g, const float* tiles, const double* weight, float* out, uint8_t* mask,
                              float thr, cudaStream_t st) {
  const int64_t stage = 8 * g.tile * g.tiles_x;
  if (g.step / PX > kMergeConsumers || g.step % PX) return false;
  int n_stages = (int)std::min<int64_t>(8, (227 * 1024 - 128) / stage);
  if (const char* e = std::getenv("SNB_MERGE_STAGES")) n_stages = std::min(n_stages, std::max(1, std::atoi(e)));
  if (n_stages < 2) return false;
  static unsigned long long configured = 0;     // per device ordinal (the attribute is per device)
  if (!configured_on_this_device(&configured)) {
    if (cudaFuncSetAttribute(merge_f32c1_ring_kernel<PX, HAS_OUT, HAS_MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             227 * 1024) != cudaSuccess)
      return false;
    mark_configured_on_this_device(&configured);
  }
  const int grid = (int)std::min<int64_t>(sm_count(), g.image_h);
  merge_f32c1_ring_kernel<PX, HAS_OUT, HAS_MASK><<<grid, kMergeConsumers + 32, 128 + (size_t)n_stages * stage, st>>>(
      g, tiles, weight, out, mask, thr, n_stages, (int)stage);
  return true;
}

template <int PX, bool HAS_OUT, bool HAS_MASK>
static bool launch_merge_staged(const SlicerGeom& g, const float* tiles, const double* weight, float* out, uint8_t* mask,
                                float thr, cudaStream_t st, int row0, int rows) {
  const int threads = (int)((g.step / PX + 31) / 32 * 32);
  if (g.step % PX || threads > kStagedThreads<PX>) return false;
  // A CTA walks `kp` periods of one row (xs segments per row).  Measured on B200 (tools/merge_bench.py): ~10-14 periods
  // per thread amortise the per-row setup (weights, norm, reciprocal) best; fewer periods lose to that setup and to the
  // weight rows every CTA stages, more periods (one CTA per 45-period row at 224/112) leave too few CTAs per SM.
  const int64_t periods = g.tiles_x + 1;
  int64_t xs = periods <= 16 ? 1 : (periods + 9) / 10;
  if (const char* e = std::getenv("SNB_MERGE_XS")) xs = std::max(1, std::min<int>((int)periods, std::atoi(e)));
  int64_t kp = (periods + xs - 1) / xs;
  while (kp > 1 && 16 + 2 * std::min(kp + 1, g.tiles_x) * g.tile * 4 > 100 * 1024) --kp;   // >= 2 CTAs per SM
  xs = (periods + kp - 1) / kp;
  const size_t smem = 16 + 2 * (size_t)(std::min(kp + 1, g.tiles_x) * g.tile * 4);
  if (smem > 227 * 1024 || rows * xs > INT32_MAX) return false;
  static unsigned long long configured = 0;     // per device ordinal
  if (!configured_on_this_device(&configured)) {
    if (cudaFuncSetAttribute(merge_f32c1_staged_kernel<PX, HAS_OUT, HAS_MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             227 * 1024) != cudaSuccess)
      return false;
    mark_configured_on_this_device(&configured);
  }
  merge_f32c1_staged_kernel<PX, HAS_OUT, HAS_MASK><<<(unsigned)(rows * xs), threads, smem, st>>>(
      g, tiles, weight, out, mask, thr, (int)xs, (int)kp, row0);
  return true;
}

// mode: 's' = CTA per row, 'r' = persistent ring; px = pixels per thread
template <bool HAS_OUT, bool HAS_MASK>
static bool launch_merge_periodic(char mode, int px, const SlicerGeom& g, const float* tiles, const double* weight,
                                  float* out, uint8_t* mask, float thr, cudaStream_t st, int row0, int rows) {
  if (mode == 'r' && row0 == 0 && rows == (int)g.image_h) {     // the persistent ring walks whole images only
    if (px == 1) return launch_merge_ring<1, HAS_OUT, HAS_MASK>(g, tiles, weight, out, mask, thr, st);
    if (px == 2) return launch_merge_ring<2, HAS_OUT, HAS_MASK>(g, tiles, weight, out, mask, thr, st);
    return launch_merge_ring<4, HAS_OUT, HAS_MASK>(g, tiles, weight, out, mask, thr, st);
  }
  if (px == 1) return launch_merge_staged<1, HAS_OUT, HAS_MASK>(g, tiles, weight, out, mask, thr, st, row0, rows);
  if (px == 2) return launch_merge_staged<2, HAS_OUT, HAS_MASK>(g, tiles, weight, out, mask, thr, st, row0, rows);
  return launch_merge_staged<4, HAS_OUT, HAS_MASK>(g, tiles, weight, out, mask, thr, st, row0, rows);
}

static int grid_for(int64_t total, int block) {
  const int64_t need = (total + block - 1) / block;
  const int64_t cap = (int64_t)sm_count() * 16;  // grid-stride: a few waves of resident CTAs
  return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}

static int64_t ceil_div_i64(int64_t a, int64_t b) { return a >= 0 ? (a + b - 1) / b : -((-a) / b); }

}  // namespace snb

using namespace snb;

extern "C" int snb_slicer_create(int64_t image_h, int64_t image_w, int64_t tile_size, int64_t tile_step,
                                 int64_t image_margin, snb_slicer** out) {
  if (!out) return fail(SNB_E_INVALID, "snb_slicer_create: null output");
  *out = nullptr;
  if (image_h <= 0 || image_w <= 0 || tile_size <= 0) return fail(SNB_E_INVALID, "bad image or tile size");
  if (tile_step < 1 || tile_step > tile_size)  // lib/tiles.py:56-57
    return fail(SNB_E_INVALID, "tile_step=%lld must be in [1, tile_size=%lld]", (long long)tile_step, (long long)tile_size);
  if (image_margin < 0) return fail(SNB_E_INVALID, "negative image_margin");
  SlicerGeom g{};
  g.image_h = image_h; g.image_w = image_w; g.tile = tile_size; g.step = tile_step;
  const int64_t overlap = tile_size - tile_step;
  if (image_margin == 0) {  // lib/tiles.py:66-78
    int64_t nw = ceil_div_i64(image_w - overlap, tile_step); if (nw < 1) nw = 1;
    int64_t nh = ceil_div_i64(image_h - overlap, tile_step); if (nh < 1) nh = 1;
    const int64_t extra_w = tile_step * nw - (image_w - overlap);
    const int64_t extra_h = tile_step * nh - (image_h - overlap);
    g.margin_left = extra_w / 2; g.margin_right = extra_w - g.margin_left;
    g.margin_top = extra_h / 2; g.margin_bottom = extra_h - g.margin_top;
  } else {  // lib/tiles.py:80-90
    if ((image_w - overlap + 2 * image_margin) % tile_step != 0 || (image_h - overlap + 2 * image_margin) % tile_step != 0)
      return fail(SNB_E_INVALID, "image_margin=%lld does not tile the image", (long long)image_margin);
    g.margin_left = g.margin_right = g.margin_top = g.margin_bottom = image_margin;
  }
  const int64_t span_y = image_h + g.margin_top + g.margin_bottom - tile_size + 1;  // range(0, span, step)
  const int64_t span_x = image_w + g.margin_left + g.margin_right - tile_size + 1;
  g.tiles_y = span_y <= 0 ? 0 : (span_y + tile_step - 1) / tile_step;
  g.tiles_x = span_x <= 0 ? 0 : (span_x + tile_step - 1) / tile_step;
  snb_slicer* s = new (std::nothrow) snb_slicer();
  if (!s) return fail(SNB_E_INVALID, "out of host memory");
  s->g = g;
  *out = s;
  return SNB_OK;
}

extern "C" void snb_slicer_destroy(snb_slicer* s) { delete s; }

extern "C" int snb_slicer_info(const snb_slicer* s, int64_t info[8]) {
  if (!s || !info) return fail(SNB_E_INVALID, "snb_slicer_info: null argument");
  info[0] = s->g.margin_left; info[1] = s->g.margin_right; info[2] = s->g.margin_top; info[3] = s->g.margin_bottom;
  info[4] = s->g.tiles_x * s->g.tiles_y; info[5] = s->g.tiles_x; info[6] = s->g.tiles_y; info[7] = s->g.tile;
  return SNB_OK;
}

extern "C" int snb_slicer_crops(const snb_slicer* s, int64_t* xy) {
  if (!s || !xy) return fail(SNB_E_INVALID, "snb_slicer_crops: null argument");
  int64_t i = 0;
  for (int64_t iy = 0; iy < s->g.tiles_y; ++iy)
    for (int64_t ix = 0; ix < s->g.tiles_x; ++ix, ++i) {
      xy[2 * i] = ix * s->g.step;
      xy[2 * i + 1] = iy * s->g.step;
    }
  return SNB_OK;
}

static int check_tile_range(const snb_slicer* s, int64_t tile_begin, int64_t tile_count) {
  const int64_t n = s->g.tiles_x * s->g.tiles_y;
  if (tile_begin < 0 || tile_count < 0 || tile_begin + tile_count > n)
    return fail(SNB_E_INVALID, "tile range [%lld, +%lld) outside the %lld crops", (long long)tile_begin,
                (long long)tile_count, (long long)n);
  return SNB_OK;
}

extern "C" int snb_split_hwc(const snb_slicer* s, const void* d_src, int64_t channels, int64_t elem_bytes,
                             int border_mode, const void* border_value, void* d_dst, int64_t tile_begin,
                             int64_t tile_count, void* stream) {
  if (!s || !d_src || !d_dst) return fail(SNB_E_INVALID, "snb_split_hwc: null argument");
  if (channels <= 0 || elem_bytes <= 0 || channels * elem_bytes > 64)
    return fail(SNB_E_INVALID, "pixel size %lld x %lld bytes unsupported (max 64 bytes)", (long long)channels, (long long)elem_bytes);
  if (border_mode != 0 && border_mode != 1) return fail(SNB_E_UNSUPPORTED, "border_mode %d (only REFLECT101=0, CONSTANT=1)", border_mode);
  if (int rc = check_tile_range(s, tile_begin, tile_count)) return rc;
  if (tile_count == 0) return SNB_OK;
  const int64_t pixel_bytes = channels * elem_bytes;
  BorderPixel bp{};
  if (border_mode == 1 && border_value) std::memcpy(bp.bytes, border_value, (size_t)pixel_bytes);
  const int64_t row_bytes = s->g.tile * pixel_bytes;
  if (tile_count * s->g.tile > 65535LL * 32768 || row_bytes > INT32_MAX / 2 || s->g.image_w * pixel_bytes > INT32_MAX / 2)
    return fail(SNB_E_UNSUPPORTED, "split too large");
  if (reinterpret_cast<uintptr_t>(d_src) & 3) return fail(SNB_E_INVALID, "the source image must be 4-byte aligned");
  const int64_t rows = tile_count * s->g.tile;
  const int64_t blocks = (rows + kSplitRows - 1) / kSplitRows;
  if (blocks > INT32_MAX) return fail(SNB_E_UNSUPPORTED, "split too large");
  split_hwc_kernel<<<(unsigned)blocks, 32 * kSplitRows, 0, as_stream(stream)>>>(
      s->g, static_cast<const uint8_t*>(d_src), (int)pixel_bytes, border_mode, bp, static_cast<uint8_t*>(d_dst), tile_begin, rows);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_split_norm_u8(const snb_slicer* s, const uint8_t* d_src, int64_t channels, const float* d_lut,
                                 int tta, int layout, void* d_dst, int64_t tile_begin, int64_t tile_count,
                                 void* stream) {
  if (!s || !d_src || !d_lut || !d_dst) return fail(SNB_E_INVALID, "snb_split_norm_u8: null argument");
  if (tta < 0 || tta > 7) return fail(SNB_E_INVALID, "tta view %d not in 0..7", tta);
  if (channels < 1 || channels > 4) return fail(SNB_E_INVALID, "channels=%lld unsupported", (long long)channels);
  if (int rc = check_tile_range(s, tile_begin, tile_count)) return rc;
  if (tile_count == 0) return SNB_OK;
  const int64_t T = s->g.tile;
  if (layout == SNB_LAYOUT_NCHW_F32) {
    const int64_t total = tile_count * channels * T * T;
    split_norm_nchw_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(
        s->g, d_src, d_lut, (int)channels, tta, static_cast<float*>(d_dst), tile_begin, total);
  } else if (layout == SNB_LAYOUT_PATCH32 || layout == SNB_LAYOUT_PATCH32_F32) {
    const bool f32 = layout == SNB_LAYOUT_PATCH32_F32;
    if (channels > 3) return fail(SNB_E_INVALID, "PATCH32 holds 9 taps x <=3 channels");
    if (reinterpret_cast<uintptr_t>(d_dst) & 15) return fail(SNB_E_INVALID, "PATCH32 destination must be 16-byte aligned");
    const int strips = (int)((T + kPatchRY - 1) / kPatchRY);
    const size_t smem = (size_t)(kPatchRY + 2) * (T + 2) * (f32 ? sizeof(float4) : sizeof(uint2)) + (size_t)(T + 2) * sizeof(int);
    if (s->g.image_h * s->g.image_w * channels > INT32_MAX) return fail(SNB_E_UNSUPPORTED, "image too large for 32-bit source offsets");
    if (smem > 200 * 1024) return fail(SNB_E_UNSUPPORTED, "tile size %lld too large for the PATCH32 split", (long long)T);
    if (tile_count * strips > INT32_MAX) return fail(SNB_E_UNSUPPORTED, "too many strips");
#define SNB_SPLIT_P32(C, F)                                                                                       \
  do {                                                                                                            \
    if (smem > 48 * 1024)                                                                                         \
      SNB_CUDA_CHECK(cudaFuncSetAttribute(split_norm_patch32_kernel<C, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)smem));                                                            \
    split_norm_patch32_kernel<C, F><<<(unsigned)(tile_count * strips), 256, smem, as_stream(stream)>>>(           \
        s->g, d_src, d_lut, tta, static_cast<uint4*>(d_dst), tile_begin);                                         \
  } while (0)
    if (f32) {
      if (channels == 3) SNB_SPLIT_P32(3, true);
      else if (channels == 2) SNB_SPLIT_P32(2, true);
      else SNB_SPLIT_P32(1, true);
    } else {
      if (channels == 3) SNB_SPLIT_P32(3, false);
      else if (channels == 2) SNB_SPLIT_P32(2, false);
      else SNB_SPLIT_P32(1, false);
    }
#undef SNB_SPLIT_P32
  } else if (layout == SNB_LAYOUT_NHWC3_BF16) {
    if (channels != 3) return fail(SNB_E_INVALID, "the packed 3-channel layout needs a 3-channel image");
    if (T % 8) return fail(SNB_E_INVALID, "the packed 3-channel layout needs a tile size that is a multiple of 8");
    if (reinterpret_cast<uintptr_t>(d_dst) & 15) return fail(SNB_E_INVALID, "destination must be 16-byte aligned");
    if (s->g.image_h * s->g.image_w * channels > INT32_MAX) return fail(SNB_E_UNSUPPORTED, "image too large for 32-bit source offsets");
    const int strips = (int)((T + kNhwcRows - 1) / kNhwcRows);
    if (tile_count * strips > INT32_MAX) return fail(SNB_E_UNSUPPORTED, "too many strips");
    const size_t smem = (size_t)(kNhwcRows + T) * sizeof(int);
    if (smem > 48 * 1024) return fail(SNB_E_UNSUPPORTED, "tile size %lld too large", (long long)T);
    split_norm_nhwc3_kernel<<<(unsigned)(tile_count * strips), 256, smem, as_stream(stream)>>>(
        s->g, d_src, d_lut, tta, static_cast<uint4*>(d_dst), tile_begin);
  } else {
    return fail(SNB_E_INVALID, "unknown layout %d", layout);
  }
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_nchw_f32_to_nhwc3(const float* d_src, int64_t n, int64_t h, int64_t w, void* d_dst, void* stream) {
  if (!d_src || !d_dst) return fail(SNB_E_INVALID, "snb_nchw_f32_to_nhwc3: null argument");
  if (n <= 0 || h <= 0 || w <= 0 || w % 8) return fail(SNB_E_INVALID, "bad shape (w must be a multiple of 8)");
  if ((reinterpret_cast<uintptr_t>(d_dst) & 15) || (reinterpret_cast<uintptr_t>(d_src) & 15))
    return fail(SNB_E_INVALID, "pointers must be 16-byte aligned");
  const int64_t total = n * h * (w / 8);
  nchw_to_nhwc3_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(d_src, (int)h, (int)w, static_cast<uint4*>(d_dst), total);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_nchw_f32_to_patch32(const float* d_src, int64_t n, int64_t channels, int64_t h, int64_t w,
                                       void* d_dst, int out_f32, void* stream) {
  if (!d_src || !d_dst) return fail(SNB_E_INVALID, "snb_nchw_f32_to_patch32: null argument");
  if (channels < 1 || channels > 3) return fail(SNB_E_INVALID, "PATCH32 holds 9 taps x <=3 channels");
  if (n <= 0 || h <= 0 || w <= 0) return fail(SNB_E_INVALID, "bad shape");
  if (reinterpret_cast<uintptr_t>(d_dst) & 15) return fail(SNB_E_INVALID, "destination must be 16-byte aligned");
  const int64_t total = n * h * w * (out_f32 ? 8 : 4);
  if (out_f32)
    nchw_to_patch32_kernel<true><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(d_src, (int)channels, (int)h, (int)w,
                                                                                       static_cast<uint4*>(d_dst), total);
  else
    nchw_to_patch32_kernel<false><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(d_src, (int)channels, (int)h, (int)w,
                                                                                        static_cast<uint4*>(d_dst), total);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

template <int TILE_DT, int TTA>
static void launch_merge(const snb_slicer* s, const void* d_tiles, int C, const double* d_weight, void* d_out,
                         int out_dtype, uint8_t* d_mask, float thr, cudaStream_t st, int row0, int rows) {
  const int64_t wc = s->g.image_w * C;
  // vector path: every row starts 16-byte aligned in all outputs
  const bool vec = wc % 4 == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_mask) & 3) == 0;
  if (vec) {
    const dim3 grid((unsigned)std::min<int64_t>((wc / 4 + 255) / 256, 1024), (unsigned)rows);
    merge_kernel<TILE_DT, TTA, 4><<<grid, 256, 0, st>>>(s->g, d_tiles, C, d_weight, d_out, out_dtype, d_mask, thr, row0);
  } else {
    const dim3 grid((unsigned)std::min<int64_t>((wc + 255) / 256, 1024), (unsigned)rows);
    merge_kernel<TILE_DT, TTA, 1><<<grid, 256, 0, st>>>(s->g, d_tiles, C, d_weight, d_out, out_dtype, d_mask, thr, row0);
  }
}

extern "C" int snb_merge_rows(const snb_slicer* s, const void* d_tiles, int tile_dtype, int64_t channels, int tta,
                              const double* d_weight, void* d_out, int out_dtype, uint8_t* d_mask, float thr,
                              int64_t row_begin, int64_t row_count, void* stream);

extern "C" int snb_merge(const snb_slicer* s, const void* d_tiles, int tile_dtype, int64_t channels, int tta,
                         const double* d_weight, void* d_out, int out_dtype, uint8_t* d_mask, float thr,
                         void* stream) {
  if (!s) return fail(SNB_E_INVALID, "snb_merge: null argument");
  return snb_merge_rows(s, d_tiles, tile_dtype, channels, tta, d_weight, d_out, out_dtype, d_mask, thr, 0, s->g.image_h, stream);
}

extern "C" int snb_merge_rows(const snb_slicer* s, const void* d_tiles, int tile_dtype, int64_t channels, int tta,
                              const double* d_weight, void* d_out, int out_dtype, uint8_t* d_mask, float thr,
                              int64_t row_begin, int64_t row_count, void* stream) {
  if (!s || !d_tiles || !d_weight) return fail(SNB_E_INVALID, "snb_merge: null argument");
  if (row_begin < 0 || row_count < 0 || row_begin + row_count > s->g.image_h)
    return fail(SNB_E_INVALID, "rows [%lld, %lld) outside the image", (long long)row_begin, (long long)(row_begin + row_count));
  if (row_count == 0) return SNB_OK;
  const int row0 = (int)row_begin, rows = (int)row_count;
  if (!d_out && !d_mask) return fail(SNB_E_INVALID, "snb_merge: no output requested");
  if (channels < 1 || channels > 64) return fail(SNB_E_INVALID, "channels=%lld unsupported", (long long)channels);
  if (tta != 1 && tta != 8) return fail(SNB_E_INVALID, "tta must be 1 or 8");
  if (tta == 8 && tile_dtype != SNB_DT_F32) return fail(SNB_E_INVALID, "TTA merge takes float32 views");
  if (out_dtype != SNB_DT_F32 && out_dtype != SNB_DT_F64 && out_dtype != SNB_DT_U8)
    return fail(SNB_E_INVALID, "out_dtype %d unsupported", out_dtype);
  if (s->g.tiles_x * s->g.tiles_y == 0) return fail(SNB_E_INVALID, "slicer has no crops");
  if (s->g.image_h > 65535 || s->g.image_w * channels > INT32_MAX / 2 || s->g.tile * s->g.tile > INT32_MAX / 2)
    return fail(SNB_E_UNSUPPORTED, "image too large for the merge kernel");
  cudaStream_t st = as_stream(stream);
  const int C = (int)channels;
  const SlicerGeom& g = s->g;
  if (tta == 1 && tile_dtype == SNB_DT_F32 && C == 1 && out_dtype == SNB_DT_F32 && g.tile % 4 == 0 && g.step % 4 == 0 &&
      g.margin_left % 4 == 0 && g.image_w % 4 == 0 && (reinterpret_cast<uintptr_t>(d_tiles) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(d_weight) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(d_mask) & 3) == 0) {
    // SNB_MERGE_MODE: "staged1|2|4" = CTA per row, "ring1|2|4" = persistent ring, "gather" = the per-pixel gather kernel
    const char* mode_env = std::getenv("SNB_MERGE_MODE");
    const std::string mm = mode_env ? mode_env : "staged2";
    bool done = false;
    if (g.tile <= 2 * g.step && mm != "gather" && mm.size() >= 5) {
      const float* tp = static_cast<const float*>(d_tiles);
      float* op = static_cast<float*>(d_out);
      const char mode = mm[0];
      const int px = mm.back() - '0';
      if (px == 1 || px == 2 || px == 4) {
        if (op && d_mask) done = launch_merge_periodic<true, true>(mode, px, g, tp, d_weight, op, d_mask, thr, st, row0, rows);
        else if (op) done = launch_merge_periodic<true, false>(mode, px, g, tp, d_weight, op, d_mask, thr, st, row0, rows);
        else done = launch_merge_periodic<false, true>(mode, px, g, tp, d_weight, op, d_mask, thr, st, row0, rows);
      }
    }
    if (!done) {
      const dim3 grid((unsigned)std::min<int64_t>((g.image_w / 4 + 255) / 256, 1024), (unsigned)rows);
      merge_f32c1_vec4_kernel<<<grid, 256, 0, st>>>(g, static_cast<const float*>(d_tiles), d_weight,
                                                    static_cast<float*>(d_out), d_mask, thr, row0);
    }
    SNB_LAUNCH_CHECK();
    return SNB_OK;
  }
  if (tta == 8) launch_merge<SNB_DT_F32, 8>(s, d_tiles, C, d_weight, d_out, out_dtype, d_mask, thr, st, row0, rows);
  else if (tile_dtype == SNB_DT_F32) launch_merge<SNB_DT_F32, 1>(s, d_tiles, C, d_weight, d_out, out_dtype, d_mask, thr, st, row0, rows);
  else if (tile_dtype == SNB_DT_U8) launch_merge<SNB_DT_U8, 1>(s, d_tiles, C, d_weight, d_out, out_dtype, d_mask, thr, st, row0, rows);
  else if (tile_dtype == SNB_DT_F64) launch_merge<SNB_DT_F64, 1>(s, d_tiles, C, d_weight, d_out, out_dtype, d_mask, thr, st, row0, rows);
  else return fail(SNB_E_INVALID, "tile_dtype %d unsupported", tile_dtype);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
