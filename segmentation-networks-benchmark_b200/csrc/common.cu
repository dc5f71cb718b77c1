// Error plumbing and device queries behind include/snb_b200.h.
#include <cstdarg>
#include <cstdio>

#include "snb_internal.h"

namespace snb {

static thread_local char g_last_error[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}

int sm_count() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  cached = n;
  return n;
}

}  // namespace snb

extern "C" int snb_version(void) { return 100; }

extern "C" const char* snb_last_error(void) { return snb::g_last_error; }

extern "C" int snb_device_sm_count(void) {
  int n = snb::sm_count();
  if (n <= 0) return snb::fail(SNB_E_CUDA, "no CUDA device available");
  return n;
}
