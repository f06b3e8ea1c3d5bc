// api.cu - error string, version and device probes of the C ABI (include/lpcyolo.h).
#include <cstdlib>
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

static thread_local char g_err[512] = "";

void lpc_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool lpc_pdl_enabled() {
  static const bool on = [] { const char* e = getenv("LPC_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

static std::atomic<unsigned long long> g_launches{0};
void lpc_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" unsigned long long lpc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" const char* lpc_last_error(void) { return g_err; }
extern "C" int lpc_abi_version(void) { return LPC_ABI_VERSION; }

extern "C" int lpc_device_arch(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) LPC_FAIL(LPC_E_CUDA, "device_arch: no CUDA device");
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  return major * 10 + minor;
}
