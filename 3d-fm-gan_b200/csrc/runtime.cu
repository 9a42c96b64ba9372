// libfm3d runtime glue: error text, launch counter, device queries.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace fm {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

bool pdl_enabled() {
  static const bool on = []() { const char* e = getenv("FM3D_PDL"); return e ? atoi(e) != 0 : true; }();
  return on;
}

}  // namespace fm

extern "C" int fm_version(void) { return 100; }
extern "C" const char* fm_last_error(void) { return fm::g_err; }
extern "C" int64_t fm_launch_count(void) { return fm::g_launches.load(std::memory_order_relaxed); }
// A captured CUDA graph replays kernels without passing through the launch wrappers: the host
// runtime credits the launches of each replay here.
extern "C" void fm_add_launches(int64_t n) { fm::g_launches.fetch_add(n, std::memory_order_relaxed); }
