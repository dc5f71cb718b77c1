// Internal helpers shared by the translation units behind include/snb_b200.h.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/snb_b200.h"

struct CUtensorMap_st;

namespace snb {

// thread-local last-error buffer; returns `code` so call sites can `return fail(...)`
int fail(int code, const char* fmt, ...);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define SNB_CUDA_CHECK(expr)                                                                        \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return ::snb::fail(SNB_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                                 \
  } while (0)

#define SNB_LAUNCH_CHECK()                                                                          \
  do {                                                                                              \
    cudaError_t _e = cudaGetLastError();                                                            \
    if (_e != cudaSuccess)                                                                          \
      return ::snb::fail(SNB_E_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),    \
                         __FILE__, __LINE__);                                                       \
  } while (0)

// crop plan shared by slicer.cu (lib/tiles.py:35-96)
struct SlicerGeom {
  int64_t image_h, image_w, tile, step;
  int64_t margin_left, margin_right, margin_top, margin_bottom;
  int64_t tiles_x, tiles_y;
};

int sm_count();

// Tap geometry of a convolution kind (shared by the forward kernels, conv_tcgen05.cu, and the weight-gradient kernel,
// conv_wgrad.cu): tile grid (where accumulators are computed), output extent, tap offsets per phase and the offset added
// to the halo-box origin.  `valid`: conv3x3 0 = padding 1, 1 = no padding ((h-2) x (w-2) outputs), 2 = "full" ((h+2) x
// (w+2) outputs: the input gradient of a valid conv3x3); conv2x2 / conv2x2-adjoint: 1 = one output less per axis.
struct ConvTapGeom {
  int n_phases, taps;
  int8_t tap_dy[4][9], tap_dx[4][9];
  int load_off;
  int64_t grid_h, grid_w, out_h, out_w;
  double flop_taps;      // real taps per grid pixel summed over the phases (FLOP accounting)
};
int conv_tap_geometry(int kind, int valid, int64_t h, int64_t w, ConvTapGeom* out);

// cuTensorMapEncodeTiled wrapper (conv_tcgen05.cu): bf16 / fp32 map over up to 4 dims, strides in bytes for dims 1..rank-1
int encode_map(CUtensorMap_st* map, void* base, int rank, const uint64_t* dims, const uint64_t* strides, const uint32_t* box,
               int swizzle_bytes, int elem_bytes);

bool configured_on_this_device(unsigned long long* mask);
void mark_configured_on_this_device(unsigned long long* mask);

}  // namespace snb

struct snb_slicer {
  snb::SlicerGeom g;
};
