// C-ABI plumbing: error string, launch counter, device properties.
#include <stdarg.h>
#include <string.h>
#include "sahs_common.cuh"

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_sahs_launches{0};

void sahs_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sahs_num_sms() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

extern "C" int sahs_abi_version(void) { return 1; }
extern "C" const char* sahs_last_error(void) { return g_err; }
extern "C" uint64_t sahs_launch_count(void) { return g_sahs_launches.load(std::memory_order_relaxed); }
