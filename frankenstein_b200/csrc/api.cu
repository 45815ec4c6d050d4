// Process-wide bookkeeping of the C-ABI: last error string, launch counter, version.
#include <atomic>
#include <cstdio>
#include <cstring>

#include "common.cuh"

static thread_local char g_last_error[512] = "";
static std::atomic<long long> g_launches{0};

void fk_set_last_error(const char* msg, const char* file, int line) {
  const char* base = strrchr(file, '/');
  snprintf(g_last_error, sizeof(g_last_error), "%s (%s:%d)", msg, base ? base + 1 : file, line);
}
void fk_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int fk_device_ordinal() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= FK_MAX_DEVICES) return 0;
  return dev;
}
int fk_sm_count() {
  static std::atomic<int> n[FK_MAX_DEVICES];
  const int dev = fk_device_ordinal();
  int v = n[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

extern "C" __attribute__((visibility("default"))) const char* fk_last_error(void) { return g_last_error; }
extern "C" __attribute__((visibility("default"))) long long fk_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" __attribute__((visibility("default"))) void fk_reset_launch_count(void) { g_launches.store(0, std::memory_order_relaxed); }
extern "C" __attribute__((visibility("default"))) int fk_abi_version(void) { return 2; }
// Compiled for exactly one architecture; the loader checks this against the device.
extern "C" __attribute__((visibility("default"))) int fk_target_sm(void) { return 100; }
extern "C" __attribute__((visibility("default"))) int fk_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10;
}
