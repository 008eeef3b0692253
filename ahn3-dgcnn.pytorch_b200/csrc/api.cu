// api.cu — error plumbing and device queries shared by the C-ABI entry points.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace dg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return DGCNN_E_CUDA;
}

int num_sms() {
  static thread_local int cached_dev = -1, cached = -1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    cached_dev = dev;
    cached = n;
  }
  return cached;
}

}  // namespace dg

extern "C" int dgcnn_version(void) { return 100; }
extern "C" const char* dgcnn_last_error(void) { return dg::g_err; }
extern "C" int dgcnn_num_sms(void) {
  const int n = dg::num_sms();
  if (n < 0) dg::set_error("no CUDA device");
  return n;
}
