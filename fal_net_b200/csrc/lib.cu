// lib.cu -- library-level plumbing of libfalnet_sm100.so: error state, launch accounting.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

namespace faln {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int after_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return FALN_ERR_LAUNCH;
  }
  return FALN_OK;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* v = getenv("FALN_PDL");
    return v != nullptr && v[0] != '0';
  }();
  return on;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace faln

extern "C" int faln_version(void) { return 100; }
extern "C" const char* faln_last_error(void) { return faln::g_err; }
extern "C" long long faln_launch_count(void) { return faln::g_launches.load(); }
