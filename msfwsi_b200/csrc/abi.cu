// Error plumbing and device check behind include/msfwsi_b200.h.
#include <stdarg.h>

#include "common.cuh"

namespace msf {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace msf

extern "C" int msf_abi_version(void) { return MSF_ABI_VERSION; }
extern "C" const char* msf_last_error(void) { return msf::g_err; }

extern "C" int msf_device_check(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  MSF_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  MSF_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  MSF_REQUIRE(prop.major == 10, MSF_ERR_UNSUPPORTED,
              "msfwsi_b200 is built for sm_100a only; current device is sm_%d%d", prop.major, prop.minor);
  return MSF_OK;
}
