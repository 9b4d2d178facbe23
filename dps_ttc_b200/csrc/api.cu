// Library-level entry points: error string, version, launch accounting.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void dps_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void dps_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" {

const char* dps_last_error(void) { return g_err; }
int dps_version(void) { return 100; }
int dps_compiled_sm(void) { return 100; }

int dps_device_sm(int* sm_out) {
  DPS_REQUIRE(sm_out != nullptr, DPS_ERR_INVALID, "dps_device_sm: null output");
  int dev = 0, major = 0, minor = 0;
  DPS_CUDA(cudaGetDevice(&dev));
  DPS_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  DPS_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  *sm_out = 10 * major + minor;
  return DPS_OK;
}

int64_t dps_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
void dps_launch_count_reset(void) { g_launches.store(0, std::memory_order_relaxed); }

}  // extern "C"
