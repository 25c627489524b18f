// C-ABI plumbing: error string, version, device query.
#include <stdarg.h>
#include "common.cuh"

namespace lavb {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace lavb

extern "C" int lavb_abi_version(void) { return LAVB_ABI_VERSION; }
extern "C" const char* lavb_last_error(void) { return lavb::g_err; }
extern "C" int lavb_device_cc(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return -1; }
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) { cudaGetLastError(); return -1; }
  return p.major * 10 + p.minor;
}
