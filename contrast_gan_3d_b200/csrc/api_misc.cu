// Version / error plumbing of the C ABI.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace cg {

char *err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char *where) {
  snprintf(err_buf(), 512, "%s: %s", where, cudaGetErrorString(e));
  return (int)e;
}

int num_sms() {
  static thread_local int cached[16] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace cg

extern "C" {

int cgan3d_version(void) { return CGAN3D_VERSION; }

const char *cgan3d_last_error(void) { return cg::err_buf(); }

uint32_t cgan3d_capabilities(void) { return 0x3u; }

int cgan3d_device_supports_tc(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

}  // extern "C"
