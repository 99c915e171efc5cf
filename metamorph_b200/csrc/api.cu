// metamorph_b200 — C-ABI plumbing shared by all kernels: thread-local error text, device queries.
#include "common.cuh"
#include <stdlib.h>
#include <stdarg.h>
#include <mutex>

static thread_local char g_err[1024] = "";

void mm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

MM_API const char* mm_last_error() { return g_err; }

int mm_pdl_enabled() {
  static const int on = [] {
    const char* e = getenv("MM_PDL");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  return on;
}

int mm_pdl_mode() {
  static const int mode = [] {
    if (!mm_pdl_enabled()) return 0;
    const char* e = getenv("MM_PDL_MODE");
    return e ? atoi(e) : 3;
  }();
  return mode;
}

int mm_num_sms() {
  static int sms = 0;
  static std::once_flag once;
  std::call_once(once, [] {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;  // B200
  });
  return sms;
}

MM_API int mm_abi_version() { return 1; }

// Returns 0 when the current device is sm_100 (B200); a negative code (+ message) otherwise.
MM_API int mm_check_device() {
  int dev = 0, major = 0, minor = 0;
  MM_CHECK_CUDA(cudaGetDevice(&dev));
  MM_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  MM_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    mm_set_error("metamorph_b200 kernels are built for sm_100a only; device is sm_%d%d", major, minor);
    return MM_ERR_ARCH;
  }
  return MM_OK;
}
