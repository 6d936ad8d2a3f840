// Error state, driver entry points and device properties for libevcdiff.so.
#include <stdlib.h>

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
extern "C" int64_t evc_struct_size(int32_t which) {
  switch (which) {
    case 0: return sizeof(evc_tensor4);
    case 1: return sizeof(evc_gemm_desc);
    case 2: return sizeof(evc_attn_desc);
    case 3: return sizeof(evc_step_coef);
    case 4: return sizeof(evc_pndm_coef);
    default: return -1;
  }
}

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

// Programmatic dependent launch: -1 = not decided (use EVC_PDL or default off), 0 / 1 = set by evc_set_pdl.
// Measured on B200 (profiles/r01_notes.md): +4 % at batch 1 (launch-latency bound), -0.5 % at batch 46.
static std::atomic<int> g_pdl{-1};

int evc_pdl_enabled() {
  static int env = -2;
  if (env == -2) {
    const char* e = getenv("EVC_PDL");
    env = (e == nullptr) ? -1 : (e[0] == '0' ? 0 : 1);
  }
  if (env >= 0) return env;  // the environment variable overrides everything (A/B runs)
  const int v = g_pdl.load();
  return v < 0 ? 0 : v;
}

extern "C" void evc_set_pdl(int enabled) { g_pdl.store(enabled ? 1 : 0); }

int evc_num_sms() {
  static std::atomic<int> sms[64];  // per device: plans are created for the current device
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int v = sms[dev].load();
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
    sms[dev].store(v);
  }
  return v;
}
