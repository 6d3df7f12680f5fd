// Library-level entry points: version, error string, device queries.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace cusrl_b200 {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace cusrl_b200

extern "C" {

int cusrl_b200_abi_version(void) { return CUSRL_B200_ABI_VERSION; }

const char* cusrl_b200_last_error(void) { return cusrl_b200::g_last_error; }

int cusrl_b200_sm_count(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cusrl_b200::set_last_error("sm_count: %s", cudaGetErrorString(e));
    return -(int)e;
  }
  return cusrl_b200::sm_count();
}

}  // extern "C"
