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
// blockIdx.y = (tile, tile row); threads walk the row's bytes, UNIT output bytes each (4 -> one 32-bit store).
// Every output byte is gathered through the reflect-101 map, so the copy is bit-exact for any element size.
constexpr int kSplitRows = 8;   // tile rows per block

template <int UNIT>
__global__ void __launch_bounds__(256) split_hwc_kernel(SlicerGeom g, const uint8_t* __restrict__ src, int pixel_bytes,
                                                        int border_mode, BorderPixel border, uint8_t* __restrict__ dst,
                                                        int64_t tile_begin, int64_t total_rows) {
  const int T = (int)g.tile;
  const int row_bytes = T * pixel_bytes;
  for (int row = blockIdx.y * kSplitRows; row < (int)min((int64_t)(blockIdx.y + 1) * kSplitRows, total_rows); ++row) {
  const int t = row / T, ty = row - t * T;
  const int64_t tile = tile_begin + t;
  const int cy = (int)((tile / g.tiles_x) * g.step), cx = (int)((tile % g.tiles_x) * g.step);
  const int py = cy + ty - (int)g.margin_top;
  const int H = (int)g.image_h, Wd = (int)g.image_w;
  const bool row_inside = py >= 0 && py < H;
  const int sy = (int)reflect101(py, H);
  const uint8_t* src_row = src + (int64_t)sy * Wd * pixel_bytes;
  uint8_t* dst_row = dst + (int64_t)row * row_bytes;
  for (int b0 = (blockIdx.x * blockDim.x + threadIdx.x) * UNIT; b0 < row_bytes; b0 += gridDim.x * blockDim.x * UNIT) {
    uint8_t out[UNIT];
#pragma unroll
    for (int b = 0; b < UNIT; ++b) {
      const int xb = b0 + b;
      const int tx = xb / pixel_bytes;
      const int pb = xb - tx * pixel_bytes;
      const int px = cx + tx - (int)g.margin_left;
      if (border_mode == 1 && !(row_inside && px >= 0 && px < Wd)) {
        out[b] = border.bytes[pb];
      } else {
        out[b] = __ldg(src_row + (int)reflect101(px, Wd) * pixel_bytes + pb);
      }
    }
    if (UNIT == 4) {
      *reinterpret_cast<uint32_t*>(dst_row + b0) =
          (uint32_t)out[0] | ((uint32_t)out[1 % UNIT] << 8) | ((uint32_t)out[2 % UNIT] << 16) | ((uint32_t)out[3 % UNIT] << 24);
    } else {
      dst_row[b0] = out[0];
    }
  }
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
  __syncthreads();
  const int PW = T + 2;
  for (int i = threadIdx.x; i < (kPatchRY + 2) * PW; i += blockDim.x) {
    const int hy = i / PW, hx = i - hy * PW;
    const int vy = y0 + hy - 1, vx = hx - 1;     // position in the (D4-transformed) tile
    float f[4] = {0.f, 0.f, 0.f, 0.f};
    if (vy >= 0 && vy < T && vx >= 0 && vx < T) {
      int si, sj;
      d4_src(tta, vy, vx, T, si, sj);
      const int64_t sy = reflect101(cy + si - g.margin_top, g.image_h);
      const int64_t sx = reflect101(cx + sj - g.margin_left, g.image_w);
      const uint8_t* px = src + (sy * g.image_w + sx) * channels;
#pragma unroll
      for (int c = 0; c < channels; ++c) f[c] = s_lut[c * 256 + __ldg(px + c)];
    }
    if (F32) {
      s_pf[i] = make_float4(to_tf32(f[0]), to_tf32(f[1]), to_tf32(f[2]), to_tf32(f[3]));
    } else {
      __nv_bfloat162 lo = __floats2bfloat162_rn(f[0], f[1]), hi = __floats2bfloat162_rn(f[2], f[3]);
      s_px[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
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
                                                    int out_dtype, uint8_t* __restrict__ mask, float thr) {
  const int T = (int)g.tile, S = (int)g.step;
  const int WC = (int)g.image_w * C;
  const int tiles_x = (int)g.tiles_x, tiles_y = (int)g.tiles_y;
  const int Y = blockIdx.y + (int)g.margin_top;              // padded-canvas row
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
    const int64_t i = (int64_t)blockIdx.y * WC + xc0;
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
                                                               float thr) {
  const int T = (int)g.tile, S = (int)g.step, W = (int)g.image_w;
  const int tiles_x = (int)g.tiles_x, tiles_y = (int)g.tiles_y;
  const int Y = blockIdx.y + (int)g.margin_top;
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
    const int64_t i = (int64_t)blockIdx.y * W + x0;
    if (out) *reinterpret_cast<float4*>(out + i) = make_float4(q[0], q[1], q[2], q[3]);
    if (mask)
      *reinterpret_cast<uchar4*>(mask + i) =
          make_uchar4(q[0] > thr ? 255 : 0, q[1] > thr ? 255 : 0, q[2] > thr ? 255 : 0, q[3] > thr ? 255 : 0);
  }
}

// a / b with b loop-invariant: y = RN(1/b) once per thread (__drcp_rn), then two FMA correction steps per quotient.
// q1 is already within half an ulp plus 2^-104 of a/b; by Markstein's theorem (y correctly rounded, q1 faithful, r1 exact)
// q2 = RN(q1 + r1*y) IS the correctly rounded quotient, i.e. bit-equal to numpy's float64 division (tests/
// test_host_logic.py checks the sequence against exact rational arithmetic).  Valid while nothing under/overflows: the
// caller keeps b in [2^-60, 2^60] and routes a outside [2^-823, 2^777) to __ddiv_rn.
__device__ __forceinline__ double div_by_invariant(double a, double b, double y) {
  const uint32_t e = (static_cast<uint32_t>(__double2hiint(a)) >> 20) & 0x7ffu;
  if (e - 200u < 1600u) {
    const double q0 = __dmul_rn(a, y);
    const double r0 = __fma_rn(-b, q0, a);
    const double q1 = __fma_rn(r0, y, q0);
    const double r1 = __fma_rn(-b, q1, a);
    return __fma_rn(r1, y, q1);
  }
  return __ddiv_rn(a, b);
}

__device__ __forceinline__ void ld_w4(const double* __restrict__ p, double (&w)[4]) {
  const double2 a = __ldg(reinterpret_cast<const double2*>(p)), b = __ldg(reinterpret_cast<const double2*>(p + 2));
  w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y;
}

// acc += v * w in the reference's arithmetic (rounded product, then rounded sum), norm += w
__device__ __forceinline__ void acc4(const float4 v, const double (&w)[4], double (&acc)[4]) {
  acc[0] = __dadd_rn(acc[0], __dmul_rn((double)v.x, w[0]));
  acc[1] = __dadd_rn(acc[1], __dmul_rn((double)v.y, w[1]));
  acc[2] = __dadd_rn(acc[2], __dmul_rn((double)v.z, w[2]));
  acc[3] = __dadd_rn(acc[3], __dmul_rn((double)v.w, w[3]));
}

// Periodic formulation of the float32 / one-channel merge for tile <= 2 * step (at most two covering crops per axis).
// Canvas pixel X = kx*step + r is covered by crop kx at tile column r and, when r < tile - step, by crop kx-1 at column
// r + step: the weights a pixel needs depend only on its residues (X mod step, Y mod step).  A thread therefore owns
// four consecutive residues of one image row, loads its <= 16 float64 weights ONCE, precomputes the (loop-invariant)
// norm and its reciprocal, and walks the ~13 periods of the row: per pixel only the tile values move (177 MB read,
// 125 MB written per 5000x5000 image; the gather kernel above re-reads 16 B of weights per covering crop and pixel
// through L2, ~700 MB).  Pixels covered by a single crop are q = RN32(RN64(RN64(v*w) / w)) = v exactly (the float64 round
// trip perturbs v by < 2^-52 relative, far inside the float32 rounding interval), so they are copied when w >= eps.
// Accumulation order (crop order: y outer, x inner), rounded products and the IEEE quotient are those of
// lib/tiles.py:146-161, hence bit-exact.
__global__ void __launch_bounds__(256) merge_f32c1_period_kernel(SlicerGeom g, const float* __restrict__ tiles,
                                                                 const double* __restrict__ weight,
                                                                 float* __restrict__ out, uint8_t* __restrict__ mask,
                                                                 float thr) {
  const int T = (int)g.tile, S = (int)g.step, W = (int)g.image_w, H = (int)g.image_h;
  const int tiles_x = (int)g.tiles_x, tiles_y = (int)g.tiles_y, ov = T - S, ml = (int)g.margin_left;
  const int r = threadIdx.x * 4;
  const int y = blockIdx.x * blockDim.y + threadIdx.y;
  if (r >= S || y >= H) return;
  const int Y = y + (int)g.margin_top;
  const int ky = Y / S, ry = Y - ky * S;
  const bool up = ky >= 1 && ry < ov;                 // crop row ky-1 covers Y (first in crop order)
  const bool two_rows = up && ky <= tiles_y - 1;
  const int iy0 = up ? ky - 1 : ky;
  const int ty0 = Y - iy0 * S;
  const bool hasA = r < ov;                           // the crop to the left (kx-1) also covers the pixel
  double wA0[4] = {0, 0, 0, 0}, wB0[4], wA1[4] = {0, 0, 0, 0}, wB1[4] = {0, 0, 0, 0};
  ld_w4(weight + ty0 * T + r, wB0);
  if (hasA) ld_w4(weight + ty0 * T + r + S, wA0);
  if (two_rows) {
    ld_w4(weight + (ty0 - S) * T + r, wB1);
    if (hasA) ld_w4(weight + (ty0 - S) * T + r + S, wA1);
  }
  // interior periods: every crop of the pattern exists
  double nrm[4], rcp[4];
  bool fast = true, copy = !hasA && !two_rows;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    double n = 0.0;
    if (hasA) n = __dadd_rn(n, wA0[e]);
    n = __dadd_rn(n, wB0[e]);
    if (two_rows) {
      if (hasA) n = __dadd_rn(n, wA1[e]);
      n = __dadd_rn(n, wB1[e]);
    }
    copy = copy && wB0[e] >= DBL_EPSILON && wB0[e] < 0x1p60;
    n = n < DBL_EPSILON ? DBL_EPSILON : n;            // np.clip(norm, eps, None)
    fast = fast && n < 0x1p60;                        // NaN / huge weights take the IEEE division
    nrm[e] = n;
    rcp[e] = __drcp_rn(n);
  }
  // periods with an output pixel: ml <= kx*S + r <= ml + W - 4
  const int kx_lo = ml > r ? (ml - r + S - 1) / S : 0;
  int kx_hi = ml + W - 4 - r >= 0 ? (ml + W - 4 - r) / S : -1;
  kx_hi = min(kx_hi, tiles_x - 1 + (hasA ? 1 : 0));
  const int64_t TT = (int64_t)T * T;
  const float* pB0 = tiles + ((int64_t)iy0 * tiles_x * T + ty0) * T + r;      // crop (iy0, kx = 0), advanced by kx*TT
  const int64_t dA = S - TT, d1 = (int64_t)tiles_x * TT - (int64_t)S * T;       // crop to the left / crop row below
  float* orow = out ? out + (int64_t)y * W - ml + r : nullptr;
  uint8_t* mrow = mask ? mask + (int64_t)y * W - ml + r : nullptr;

  auto emit = [&](int kx, const float (&q)[4]) {
    const int64_t o = (int64_t)kx * S;
    if (orow) *reinterpret_cast<float4*>(orow + o) = make_float4(q[0], q[1], q[2], q[3]);
    if (mrow)
      *reinterpret_cast<uchar4*>(mrow + o) =
          make_uchar4(q[0] > thr ? 255 : 0, q[1] > thr ? 255 : 0, q[2] > thr ? 255 : 0, q[3] > thr ? 255 : 0);
  };
  // a period at the canvas edge: some crops of the pattern do not exist; norm on the fly, IEEE division
  auto edge = [&](int kx) {
    const bool useA = hasA && kx >= 1, useB = kx <= tiles_x - 1;
    const float* p = pB0 + kx * TT;
    double acc[4] = {0, 0, 0, 0}, n[4] = {0, 0, 0, 0};
    if (useA) {
      acc4(__ldg(reinterpret_cast<const float4*>(p + dA)), wA0, acc);
#pragma unroll
      for (int e = 0; e < 4; ++e) n[e] = __dadd_rn(n[e], wA0[e]);
    }
    if (useB) {
      acc4(__ldg(reinterpret_cast<const float4*>(p)), wB0, acc);
#pragma unroll
      for (int e = 0; e < 4; ++e) n[e] = __dadd_rn(n[e], wB0[e]);
    }
    if (two_rows) {
      if (useA) {
        acc4(__ldg(reinterpret_cast<const float4*>(p + d1 + dA)), wA1, acc);
#pragma unroll
        for (int e = 0; e < 4; ++e) n[e] = __dadd_rn(n[e], wA1[e]);
      }
      if (useB) {
        acc4(__ldg(reinterpret_cast<const float4*>(p + d1)), wB1, acc);
#pragma unroll
        for (int e = 0; e < 4; ++e) n[e] = __dadd_rn(n[e], wB1[e]);
      }
    }
    float q[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) q[e] = __double2float_rn(__ddiv_rn(acc[e], n[e] < DBL_EPSILON ? DBL_EPSILON : n[e]));
    emit(kx, q);
  };

  int k0 = kx_lo, k1 = kx_hi;
  if (k0 <= k1 && hasA && k0 == 0) edge(k0++);
  if (k0 <= k1 && k1 >= tiles_x) edge(k1--);
  if (copy) {
#pragma unroll 4
    for (int kx = k0; kx <= k1; ++kx) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(pB0 + kx * TT));
      const float q[4] = {v.x, v.y, v.z, v.w};
      emit(kx, q);
    }
    return;
  }
#pragma unroll 2
  for (int kx = k0; kx <= k1; ++kx) {
    const float* p = pB0 + kx * TT;
    float4 vA0 = make_float4(0.f, 0.f, 0.f, 0.f), vA1 = vA0, vB1 = vA0;
    const float4 vB0 = __ldg(reinterpret_cast<const float4*>(p));
    if (hasA) vA0 = __ldg(reinterpret_cast<const float4*>(p + dA));
    if (two_rows) {
      vB1 = __ldg(reinterpret_cast<const float4*>(p + d1));
      if (hasA) vA1 = __ldg(reinterpret_cast<const float4*>(p + d1 + dA));
    }
    double acc[4] = {0, 0, 0, 0};
    if (hasA) acc4(vA0, wA0, acc);
    acc4(vB0, wB0, acc);
    if (two_rows) {
      if (hasA) acc4(vA1, wA1, acc);
      acc4(vB1, wB1, acc);
    }
    float q[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      q[e] = __double2float_rn(fast ? div_by_invariant(acc[e], nrm[e], rcp[e]) : __ddiv_rn(acc[e], nrm[e]));
    emit(kx, q);
  }
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
  const bool words = row_bytes % 4 == 0 && (reinterpret_cast<uintptr_t>(d_dst) & 3) == 0;
  // gridDim.y is limited to 65535 rows per launch: split the tile range
  const int64_t tiles_per_launch = std::max<int64_t>(1, 65535LL * kSplitRows / s->g.tile);
  for (int64_t t0 = 0; t0 < tile_count; t0 += tiles_per_launch) {
    const int64_t nt = std::min(tiles_per_launch, tile_count - t0);
    uint8_t* dst = static_cast<uint8_t*>(d_dst) + t0 * s->g.tile * row_bytes;
    const int unit = words ? 4 : 1;
    const int64_t rows = nt * s->g.tile;
    const dim3 grid((unsigned)std::min<int64_t>((row_bytes / unit + 255) / 256, 64), (unsigned)((rows + kSplitRows - 1) / kSplitRows));
    if (words)
      split_hwc_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(s->g, static_cast<const uint8_t*>(d_src), (int)pixel_bytes,
                                                              border_mode, bp, dst, tile_begin + t0, rows);
    else
      split_hwc_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(s->g, static_cast<const uint8_t*>(d_src), (int)pixel_bytes,
                                                              border_mode, bp, dst, tile_begin + t0, rows);
  }
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
    const size_t smem = (size_t)(kPatchRY + 2) * (T + 2) * (f32 ? sizeof(float4) : sizeof(uint2));
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
  } else {
    return fail(SNB_E_INVALID, "unknown layout %d", layout);
  }
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
                         int out_dtype, uint8_t* d_mask, float thr, cudaStream_t st) {
  const int64_t wc = s->g.image_w * C;
  // vector path: every row starts 16-byte aligned in all outputs
  const bool vec = wc % 4 == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_mask) & 3) == 0;
  if (vec) {
    const dim3 grid((unsigned)std::min<int64_t>((wc / 4 + 255) / 256, 1024), (unsigned)s->g.image_h);
    merge_kernel<TILE_DT, TTA, 4><<<grid, 256, 0, st>>>(s->g, d_tiles, C, d_weight, d_out, out_dtype, d_mask, thr);
  } else {
    const dim3 grid((unsigned)std::min<int64_t>((wc + 255) / 256, 1024), (unsigned)s->g.image_h);
    merge_kernel<TILE_DT, TTA, 1><<<grid, 256, 0, st>>>(s->g, d_tiles, C, d_weight, d_out, out_dtype, d_mask, thr);
  }
}

extern "C" int snb_merge(const snb_slicer* s, const void* d_tiles, int tile_dtype, int64_t channels, int tta,
                         const double* d_weight, void* d_out, int out_dtype, uint8_t* d_mask, float thr,
                         void* stream) {
  if (!s || !d_tiles || !d_weight) return fail(SNB_E_INVALID, "snb_merge: null argument");
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
    static const bool use_gather = std::getenv("SNB_MERGE_GATHER") != nullptr;   // A/B switch for profiles/
    if (g.tile <= 2 * g.step && g.step / 4 <= 256 && !use_gather) {
      // periodic kernel: blockDim = (step / 4 residue groups, rows that fit in 256 threads)
      const int bx = (int)(g.step / 4), by = std::max(1, 256 / bx);
      merge_f32c1_period_kernel<<<(unsigned)((g.image_h + by - 1) / by), dim3(bx, by), 0, st>>>(
          g, static_cast<const float*>(d_tiles), d_weight, static_cast<float*>(d_out), d_mask, thr);
    } else {
      const dim3 grid((unsigned)std::min<int64_t>((g.image_w / 4 + 255) / 256, 1024), (unsigned)g.image_h);
      merge_f32c1_vec4_kernel<<<grid, 256, 0, st>>>(g, static_cast<const float*>(d_tiles), d_weight,
                                                    static_cast<float*>(d_out), d_mask, thr);
    }
    SNB_LAUNCH_CHECK();
    return SNB_OK;
  }
  if (tta == 8) launch_merge<SNB_DT_F32, 8>(s, d_tiles, C, d_weight, d_out, out_dtype, d_mask, thr, st);
  else if (tile_dtype == SNB_DT_F32) launch_merge<SNB_DT_F32, 1>(s, d_tiles, C, d_weight, d_out, out_dtype, d_mask, thr, st);
  else if (tile_dtype == SNB_DT_U8) launch_merge<SNB_DT_U8, 1>(s, d_tiles, C, d_weight, d_out, out_dtype, d_mask, thr, st);
  else if (tile_dtype == SNB_DT_F64) launch_merge<SNB_DT_F64, 1>(s, d_tiles, C, d_weight, d_out, out_dtype, d_mask, thr, st);
  else return fail(SNB_E_INVALID, "tile_dtype %d unsupported", tile_dtype);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
