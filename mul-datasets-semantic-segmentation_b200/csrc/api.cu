// api.cu — library-level entry points: version, error text, device info.
#include <stdarg.h>
#include <stdio.h>

#include <mutex>

#include "common.cuh"

namespace mdseg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  // Per-device cache; the only mutable global state of the library.
  static std::mutex mu;
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lk(mu);
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev] = n;
  }
  return cache[dev];
}

}  // namespace mdseg

extern "C" {

int mdseg_version(void) { return MDSEG_VERSION; }
const char* mdseg_last_error(void) { return mdseg::g_err; }
int mdseg_sm_count(void) { return mdseg::sm_count(); }
size_t mdseg_ohem_state_bytes(void) { return sizeof(mdseg_ohem_state); }

}  // extern "C"
