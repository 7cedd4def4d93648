// Error plumbing and device queries of the C-ABI.
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include "common.cuh"

namespace avdf {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return AVDF_ERR_CUDA;
  }
  return AVDF_OK;
}
bool pdl_enabled() {
  static const int on = getenv("AVDF_PDL") ? atoi(getenv("AVDF_PDL")) : 1;
  return on != 0;
}

int current_device() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  return dev;
}

int device_sm_count() {
  static int sms[256];
  const int dev = current_device();
  if (dev < 0 || dev >= 256) return 148;
  int v = __atomic_load_n(&sms[dev], __ATOMIC_RELAXED);
  if (v <= 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    __atomic_store_n(&sms[dev], v, __ATOMIC_RELAXED);
  }
  return v;
}
}  // namespace avdf

extern "C" int avdf_abi_version(void) { return AVDF_ABI_VERSION; }
extern "C" const char* avdf_last_error(void) { return avdf::g_err; }

extern "C" int avdf_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
  int dev = 0;
  AVDF_CUDA(cudaGetDevice(&dev));
  int v = 0;
  if (sm_count) { AVDF_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)); *sm_count = v; }
  if (cc_major) { AVDF_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev)); *cc_major = v; }
  if (cc_minor) { AVDF_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev)); *cc_minor = v; }
  return AVDF_OK;
}
