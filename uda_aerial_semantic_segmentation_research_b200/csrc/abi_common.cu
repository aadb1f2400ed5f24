// Library-level entry points: thread-local error message, ABI version, device check.
#include "common.cuh"
#include <string.h>

namespace uda {
static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace uda

extern "C" const char* uda_last_error(void) { return uda::g_err; }
extern "C" int uda_abi_version(void) { return 1; }
extern "C" int uda_device_supported(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}
