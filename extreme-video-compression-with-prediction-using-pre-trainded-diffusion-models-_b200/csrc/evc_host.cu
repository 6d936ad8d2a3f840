// Error state, driver entry points and device properties for libevcdiff.so.
#include "evc_host.h"

#include <atomic>
#include <mutex>

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int evc_set_error(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg ? msg : "");
  return code;
}

extern "C" const char* evc_last_error(void) { return g_err; }
extern "C" int evc_version(void) { return 1; }
extern "C" int64_t evc_launch_count(void) { return (int64_t)g_launches.load(); }

int evc_check_launch(const char* what) {
  g_launches.fetch_add(1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    char buf[384];
    snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    return evc_set_error(EVC_ERR_CUDA, buf);
  }
  return EVC_OK;
}

PFN_encodeTiled evc_get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(p);
    else (void)cudaGetLastError();
  });
  return fn;
}

int evc_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
    sms = v;
  }
  return sms;
}
