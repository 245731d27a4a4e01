// api.cu - error reporting and version of libpcp_b200.so
#include <stdarg.h>
#include "common.cuh"

namespace pcp {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return (int)e;
}
// Multiprocessors of the current device.  An immutable device property, cached per device ordinal (the only static state
// of the library; racing first calls write the same value).
int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  if (dev >= 0 && dev < 64) cached[dev] = n;
  return n;
}
}  // namespace pcp

extern "C" int pcp_abi_version(void) { return PCP_ABI_VERSION; }
extern "C" const char* pcp_last_error_string(void) { return pcp::g_err; }
