// Library-wide state of libicaf_b200: version, thread-local error string, device properties.
#include <cstring>

#include "icaf_internal.cuh"

namespace icaf {

static thread_local char g_err[512] = "";

int set_error(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
int set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
  return ICAF_ERR_CUDA;
}
int check_launch(const char* where) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, where);
  return ICAF_OK;
}
int sm_count_cached() {
  static int sms = 0;
  if (sms <= 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        sms <= 0)
      sms = 148;
  }
  return sms;
}

}  // namespace icaf

extern "C" int icaf_version(void) { return 100; }   // 0.1.0
extern "C" const char* icaf_last_error(void) { return icaf::g_err; }
extern "C" int icaf_sm_count(void) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return sms;
}
