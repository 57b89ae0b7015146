// C-ABI plumbing: error string, launch counter, device properties.
#include <stdarg.h>
#include <string.h>
#include "sahs_common.cuh"
#include "field_plan.cuh"

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

// Diagnostic words of the tcgen05 kernels (field forward, dgrad, wgrad, Stage-II conv: 4 ints each) live in mapped pinned host memory, so
// they stay readable after a kernel trapped and the context is gone.  Allocated once per process.
static int* g_status_host = nullptr;
int* sahs_status_words(int which) {
  static int* dev = [] {
    int* h = nullptr;
    int* d = nullptr;
    if (cudaHostAlloc((void**)&h, 16 * sizeof(int), cudaHostAllocMapped) != cudaSuccess) return (int*)nullptr;
    for (int i = 0; i < 16; ++i) h[i] = 0;
    if (cudaHostGetDevicePointer((void**)&d, h, 0) != cudaSuccess) return (int*)nullptr;
    g_status_host = h;
    return d;
  }();
  return dev ? dev + 4 * which : nullptr;
}

extern "C" int sahs_spade_conv_status(int* out4_host) {
  for (int i = 0; i < 4; ++i) out4_host[i] = 0;
  if (!g_status_host) return SAHS_OK;
  const volatile int* w = g_status_host + 12;
  for (int i = 0; i < 4; ++i) out4_host[i] = w[i];
  return SAHS_OK;
}

extern "C" int sahs_field_status(int* out4_host) {
  for (int i = 0; i < 4; ++i) out4_host[i] = 0;
  if (!g_status_host) return SAHS_OK;   // no field kernel launched yet
  for (int k = 0; k < 3; ++k) {
    const volatile int* w = g_status_host + 4 * k;
    if (w[0] != 0) {
      for (int i = 0; i < 4; ++i) out4_host[i] = w[i];
      out4_host[0] += 100 * k;          // 1xx: dgrad kernel, 2xx: wgrad kernel
      break;
    }
  }
  return SAHS_OK;
}

extern "C" int sahs_abi_version(void) { return 2; }   // 2: sahs_field_wgrad takes an upload token
extern "C" int sahs_operand_format(void) { return kRenderTrunkF16 ? 0 : 1; }
extern "C" const char* sahs_last_error(void) { return g_err; }
extern "C" uint64_t sahs_launch_count(void) { return g_sahs_launches.load(std::memory_order_relaxed); }
