// C-ABI plumbing: version, thread-local error string, device info.
#include <stdarg.h>
#include "common.cuh"

namespace cst {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace cst

extern "C" int cst_abi_version(void) { return CST_ABI_VERSION; }
extern "C" const char* cst_last_error(void) { return cst::g_err; }

extern "C" int cst_device_info(char* name, int name_cap, int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  CST_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CST_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (name && name_cap > 0) { strncpy(name, prop.name, name_cap - 1); name[name_cap - 1] = 0; }
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return CST_OK;
}
