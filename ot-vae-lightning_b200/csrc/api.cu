// libotk: process-level plumbing of the C ABI (include/otk.h).
#include "otk_common.cuh"
#include <atomic>
#include <cstdlib>

namespace otk {
char* last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}
static int g_dev_checked[64] = {0};  // 0 unknown, 1 ok, -1 unsupported
static int g_sms[64] = {0};
unsigned long long launches();
int require_device() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { set_last_error("cudaGetDevice", e); return OTK_ERR_CUDA; }
  if (dev < 0 || dev >= 64) return OTK_ERR_UNSUPPORTED_DEVICE;
  if (g_dev_checked[dev] == 0) {
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) { set_last_error("cudaGetDeviceProperties", e); return OTK_ERR_CUDA; }
    g_sms[dev] = p.multiProcessorCount;
    g_dev_checked[dev] = (p.major == 10) ? 1 : -1;
  }
  if (g_dev_checked[dev] < 0) {
    set_last_error_msg("libotk is built for sm_100a (B200) only; no fallback path exists");
    return OTK_ERR_UNSUPPORTED_DEVICE;
  }
  return OTK_OK;
}
static std::atomic<unsigned long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
unsigned long long launches() { return g_launches.load(std::memory_order_relaxed); }
bool pdl_enabled(int site) {
  // sites: 0 main statistics kernels, 1 conditional clear, 2 gated TF32 fallback, 3 merge kernels.  Measured per update call
  // (65536 x 512 chunk): none 97.3 us, sites 0-2 90.7 us, all four 101.2 us - the merge kernels launch plainly by default.
  static const int mask = [] { const char* e = getenv("OTK_PDL"); return e ? atoi(e) : 7; }();
  return (mask >> site) & 1;
}
int sm_count() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < 64 && g_sms[dev] > 0) ? g_sms[dev] : 148;
}
}  // namespace otk

extern "C" unsigned long long otk_launch_count(void) { return otk::launches(); }
extern "C" int otk_abi_version(void) { return OTK_ABI_VERSION; }
extern "C" const char* otk_last_error(void) { return otk::last_error_buffer(); }
extern "C" int otk_device_supported(void) { return otk::require_device() == OTK_OK ? 1 : 0; }
extern "C" const char* otk_status_string(int status) {
  switch (status) {
    case OTK_OK: return "ok";
    case OTK_ERR_INVALID_ARGUMENT: return "invalid argument";
    case OTK_ERR_WORKSPACE: return "workspace missing or too small";
    case OTK_ERR_CUDA: return "CUDA error";
    case OTK_ERR_UNSUPPORTED_DEVICE: return "unsupported device (sm_100a required)";
    case OTK_ERR_NOT_CONVERGED: return "iteration did not converge";
    default: return "unknown status";
  }
}
