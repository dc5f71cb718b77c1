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

// SM count of the CURRENT device (cached per device ordinal: a process may drive several GPUs)
int sm_count() {
  constexpr int kMaxDev = 64;
  static int cached[kMaxDev] = {0};
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (dev >= 0 && dev < kMaxDev && cached[dev] > 0) return cached[dev];
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  if (dev >= 0 && dev < kMaxDev) cached[dev] = n;
  return n;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: remember which devices a kernel was configured on
bool configured_on_this_device(unsigned long long* mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
  return (*mask >> dev) & 1ull;
}
void mark_configured_on_this_device(unsigned long long* mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) *mask |= 1ull << dev;
}

}  // namespace snb

extern "C" int snb_version(void) { return 100; }

extern "C" const char* snb_last_error(void) { return snb::g_last_error; }

extern "C" int snb_device_sm_count(void) {
  int n = snb::sm_count();
  if (n <= 0) return snb::fail(SNB_E_CUDA, "no CUDA device available");
  return n;
}
