// ImageSlicer on the device: crop plan, reflect-101 split, fused normalising split, weighted overlap-add merge.
// Follows lib/tiles.py:35-161 (plan, split, merge), lib/augmentations.py:452-511 (NormalizeImage via a host
// built LUT, D4 TTA index maps) and inria_submit.py:305 (threshold).  All of it is HBM-bound byte/integer
// work: one thread per output element (x fastest) so stores are fully coalesced and loads are row segments.
#include <cuda_bf16.h>

#include <cfloat>
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
// dst viewed as 32-bit words (row bytes are a multiple of 4 is required by the host wrapper) or bytes.
template <int UNIT>  // bytes produced per thread: 4 (byte-gather) or 1
__global__ void split_hwc_kernel(SlicerGeom g, const uint8_t* __restrict__ src, int64_t pixel_bytes, int border_mode,
                                 BorderPixel border, uint8_t* __restrict__ dst, int64_t tile_begin,
                                 int64_t total_units) {
  const int64_t row_bytes = g.tile * pixel_bytes;
  const int64_t tile_bytes = g.tile * row_bytes;
  for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < total_units;
       u += (int64_t)gridDim.x * blockDim.x) {
    const int64_t byte0 = u * UNIT;
    const int64_t t = byte0 / tile_bytes;
    const int64_t r = byte0 - t * tile_bytes;
    const int64_t ty = r / row_bytes;
    const int64_t rb = r - ty * row_bytes;
    const int64_t tile = tile_begin + t;
    const int64_t cy = (tile / g.tiles_x) * g.step, cx = (tile % g.tiles_x) * g.step;
    const int64_t py = cy + ty - g.margin_top;
    uint8_t out[UNIT];
#pragma unroll
    for (int b = 0; b < UNIT; ++b) {
      const int64_t xb = rb + b;
      const int64_t tx = xb / pixel_bytes;
      const int64_t pb = xb - tx * pixel_bytes;
      const int64_t px = cx + tx - g.margin_left;
      const bool inside = py >= 0 && py < g.image_h && px >= 0 && px < g.image_w;
      if (border_mode == 1 && !inside) {
        out[b] = border.bytes[pb];
      } else {
        const int64_t sy = reflect101(py, g.image_h), sx = reflect101(px, g.image_w);
        out[b] = src[(sy * g.image_w + sx) * pixel_bytes + pb];
      }
    }
    if (UNIT == 4) {
      *reinterpret_cast<uint32_t*>(dst + byte0) =
          (uint32_t)out[0] | ((uint32_t)out[1] << 8) | ((uint32_t)out[2 % UNIT] << 16) | ((uint32_t)out[3 % UNIT] << 24);
    } else {
      dst[byte0] = out[0];
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

template <int channels>
__global__ void __launch_bounds__(256) split_norm_patch32_kernel(SlicerGeom g, const uint8_t* __restrict__ src,
                                                                 const float* __restrict__ lut, int tta,
                                                                 uint4* __restrict__ dst, int64_t tile_begin) {
  extern __shared__ uint2 s_px[];                 // [(RY+2)][T+2] pixels, 4 x bf16 each
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
    __nv_bfloat162 lo = __floats2bfloat162_rn(f[0], f[1]), hi = __floats2bfloat162_rn(f[2], f[3]);
    s_px[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
  __syncthreads();
  const unsigned short* s_el = reinterpret_cast<const unsigned short*>(s_px);
  const int rows = min(kPatchRY, T - y0);
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

__global__ void nchw_to_patch32_kernel(const float* __restrict__ src, int channels, int H, int W,
                                       uint4* __restrict__ dst, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int qd = (int)(i & 3);
    int64_t r = i >> 2;
    const int x = (int)(r % W); r /= W;
    const int y = (int)(r % H);
    const int64_t n = r / H;
    const float* img = src + n * channels * (int64_t)H * W;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = qd * 8 + e;
      const int tap = k / channels, c = k - tap * channels;
      const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
      f[e] = (tap < 9 && yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(img + ((int64_t)c * H + yy) * W + xx) : 0.f;
    }
    __nv_bfloat162 p0 = __floats2bfloat162_rn(f[0], f[1]), p1 = __floats2bfloat162_rn(f[2], f[3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(f[4], f[5]), p3 = __floats2bfloat162_rn(f[6], f[7]);
    dst[i] = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                        *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
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

template <int TILE_DT, int TTA>
__global__ void merge_kernel(SlicerGeom g, const void* __restrict__ tiles, int C, const double* __restrict__ weight,
                             void* __restrict__ out, int out_dtype, uint8_t* __restrict__ mask, float thr,
                             int64_t total) {
  const int64_t T = g.tile, S = g.step;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int c = (int)(r % C); r /= C;
    const int64_t x = r % g.image_w;
    const int64_t y = r / g.image_w;
    const int64_t Y = y + g.margin_top, X = x + g.margin_left;  // padded-canvas coordinates
    // crops covering (Y, X): iy*S <= Y < iy*S + T
    int64_t iy0 = Y - T + 1 <= 0 ? 0 : (Y - T + S) / S;
    int64_t iy1 = Y / S; if (iy1 > g.tiles_y - 1) iy1 = g.tiles_y - 1;
    int64_t ix0 = X - T + 1 <= 0 ? 0 : (X - T + S) / S;
    int64_t ix1 = X / S; if (ix1 > g.tiles_x - 1) ix1 = g.tiles_x - 1;
    double acc = 0.0, norm = 0.0;
    for (int64_t iy = iy0; iy <= iy1; ++iy) {        // crop order: y outer, x inner (lib/tiles.py:94-96,150)
      const int ty = (int)(Y - iy * S);
      for (int64_t ix = ix0; ix <= ix1; ++ix) {
        const int tx = (int)(X - ix * S);
        const double w = __ldg(weight + (int64_t)ty * T + tx);
        const double v = tile_value<TILE_DT, TTA>(tiles, iy * g.tiles_x + ix, ty, tx, c, (int)T, C);
        acc = __dadd_rn(acc, __dmul_rn(v, w));       // no FMA contraction: numpy rounds the product first
        norm = __dadd_rn(norm, w);
      }
    }
    norm = norm < DBL_EPSILON ? DBL_EPSILON : norm;  // np.clip(norm, eps, None)
    const double q = __ddiv_rn(acc, norm);
    const float qf = __double2float_rn(q);
    if (out) {
      if (out_dtype == SNB_DT_F32) static_cast<float*>(out)[i] = qf;
      else if (out_dtype == SNB_DT_F64) static_cast<double*>(out)[i] = q;
      else static_cast<uint8_t*>(out)[i] = (uint8_t)(int)q;  // astype(uint8) truncates
    }
    if (mask) mask[i] = qf > thr ? 255 : 0;
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
  const int64_t total_bytes = tile_count * s->g.tile * s->g.tile * pixel_bytes;
  const bool words = (s->g.tile * pixel_bytes) % 4 == 0 && (reinterpret_cast<uintptr_t>(d_dst) & 3) == 0;
  if (words) {
    const int64_t units = total_bytes / 4;
    split_hwc_kernel<4><<<grid_for(units, 256), 256, 0, as_stream(stream)>>>(
        s->g, static_cast<const uint8_t*>(d_src), pixel_bytes, border_mode, bp, static_cast<uint8_t*>(d_dst), tile_begin, units);
  } else {
    split_hwc_kernel<1><<<grid_for(total_bytes, 256), 256, 0, as_stream(stream)>>>(
        s->g, static_cast<const uint8_t*>(d_src), pixel_bytes, border_mode, bp, static_cast<uint8_t*>(d_dst), tile_begin, total_bytes);
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
  } else if (layout == SNB_LAYOUT_PATCH32) {
    if (channels > 3) return fail(SNB_E_INVALID, "PATCH32 holds 9 taps x <=3 channels");
    if (reinterpret_cast<uintptr_t>(d_dst) & 15) return fail(SNB_E_INVALID, "PATCH32 destination must be 16-byte aligned");
    const int strips = (int)((T + kPatchRY - 1) / kPatchRY);
    const size_t smem = (size_t)(kPatchRY + 2) * (T + 2) * sizeof(uint2);
    if (smem > 200 * 1024) return fail(SNB_E_UNSUPPORTED, "tile size %lld too large for the PATCH32 split", (long long)T);
    if (tile_count * strips > INT32_MAX) return fail(SNB_E_UNSUPPORTED, "too many strips");
#define SNB_SPLIT_P32(C)                                                                                          \
  do {                                                                                                            \
    if (smem > 48 * 1024)                                                                                         \
      SNB_CUDA_CHECK(cudaFuncSetAttribute(split_norm_patch32_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)smem));                                                            \
    split_norm_patch32_kernel<C><<<(unsigned)(tile_count * strips), 256, smem, as_stream(stream)>>>(              \
        s->g, d_src, d_lut, tta, static_cast<uint4*>(d_dst), tile_begin);                                         \
  } while (0)
    if (channels == 3) SNB_SPLIT_P32(3);
    else if (channels == 2) SNB_SPLIT_P32(2);
    else SNB_SPLIT_P32(1);
#undef SNB_SPLIT_P32
  } else {
    return fail(SNB_E_INVALID, "unknown layout %d", layout);
  }
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_nchw_f32_to_patch32(const float* d_src, int64_t n, int64_t channels, int64_t h, int64_t w,
                                       void* d_dst, void* stream) {
  if (!d_src || !d_dst) return fail(SNB_E_INVALID, "snb_nchw_f32_to_patch32: null argument");
  if (channels < 1 || channels > 3) return fail(SNB_E_INVALID, "PATCH32 holds 9 taps x <=3 channels");
  if (n <= 0 || h <= 0 || w <= 0) return fail(SNB_E_INVALID, "bad shape");
  if (reinterpret_cast<uintptr_t>(d_dst) & 15) return fail(SNB_E_INVALID, "destination must be 16-byte aligned");
  const int64_t total = n * h * w * 4;
  nchw_to_patch32_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(d_src, (int)channels, (int)h, (int)w,
                                                                               static_cast<uint4*>(d_dst), total);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

template <int TILE_DT, int TTA>
static void launch_merge(const snb_slicer* s, const void* d_tiles, int C, const double* d_weight, void* d_out,
                         int out_dtype, uint8_t* d_mask, float thr, cudaStream_t st) {
  const int64_t total = s->g.image_h * s->g.image_w * C;
  merge_kernel<TILE_DT, TTA><<<grid_for(total, 256), 256, 0, st>>>(s->g, d_tiles, C, d_weight, d_out, out_dtype, d_mask,
                                                                  thr, total);
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
  cudaStream_t st = as_stream(stream);
  const int C = (int)channels;
  if (tta == 8) launch_merge<SNB_DT_F32, 8>(s, d_tiles, C, d_weight, d_out, out_dtype, d_mask, thr, st);
  else if (tile_dtype == SNB_DT_F32) launch_merge<SNB_DT_F32, 1>(s, d_tiles, C, d_weight, d_out, out_dtype, d_mask, thr, st);
  else if (tile_dtype == SNB_DT_U8) launch_merge<SNB_DT_U8, 1>(s, d_tiles, C, d_weight, d_out, out_dtype, d_mask, thr, st);
  else if (tile_dtype == SNB_DT_F64) launch_merge<SNB_DT_F64, 1>(s, d_tiles, C, d_weight, d_out, out_dtype, d_mask, thr, st);
  else return fail(SNB_E_INVALID, "tile_dtype %d unsupported", tile_dtype);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
