// Error reporting, device queries and the launch counter of libnerfb200.so.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

namespace nerfb200 {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
  static int cached = 0;
  // NERFB200_MAX_GRID limits the persistent grids (profiling experiments only)
  static int limit = -1;
  if (limit < 0) {
    const char* e = getenv("NERFB200_MAX_GRID");
    limit = e ? atoi(e) : 0;
  }
  if (limit > 0) return limit;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;  // B200
  }
  return cached;
}

}  // namespace nerfb200

extern "C" const char* nerfb200_last_error(void) { return nerfb200::g_err; }
extern "C" int nerfb200_abi_version(void) { return NERFB200_ABI_VERSION; }
extern "C" long long nerfb200_launch_count(void) {
  return nerfb200::g_launches.load(std::memory_order_relaxed);
}
