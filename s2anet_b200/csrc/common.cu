// common.cu -- library-wide state: thread-local error message, version, device queries.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace s2a {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached[dev] = v;
  }
  return cached[dev];
}

}  // namespace s2a

extern "C" int s2a_version(void) { return 100; }
extern "C" const char* s2a_last_error(void) { return s2a::g_err; }
