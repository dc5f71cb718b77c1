// Internal helpers shared by the translation units behind include/snb_b200.h.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/snb_b200.h"

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
bool configured_on_this_device(unsigned long long* mask);
void mark_configured_on_this_device(unsigned long long* mask);

}  // namespace snb

struct snb_slicer {
  snb::SlicerGeom g;
};
