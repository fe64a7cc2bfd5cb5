#include <stdlib.h>
// Error string, version and launch counter of liblinkb200.
#include <stdarg.h>
#include <atomic>

#include "common.cuh"

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void lk_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void lk_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" const char* lk_last_error(void) { return g_err; }
extern "C" int lk_version(void) { return 100; }
extern "C" int64_t lk_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
// "domain:bus:device.function" of a CUDA device (the sysfs name under /sys/bus/pci/devices): the host
// side uses it to place pinned staging buffers on the NUMA node the GPU hangs off (sharding.py).
extern "C" int lk_device_pci_bus_id(int device, char* buf, int len) {
  if (!buf || len < 16) return LK_EINVAL;
  cudaError_t e = cudaDeviceGetPCIBusId(buf, len, device);
  if (e != cudaSuccess) {
    lk_set_error("cudaDeviceGetPCIBusId(%d): %s", device, cudaGetErrorString(e));
    return LK_ECUDA;
  }
  return LK_OK;
}

// Pinned host staging memory for the upload path.  write_combined = 1: cudaHostAllocWriteCombined -- the CPU
// writes it through write-combining buffers and never caches it, so the DMA engine's reads are not snooped
// through the CPU caches; meant for buffers the host only WRITES (features on their way to the device).
extern "C" int lk_host_alloc(int64_t bytes, int write_combined, void** out) {
  if (!out || bytes <= 0) return LK_EINVAL;
  cudaError_t e = cudaHostAlloc(out, (size_t)bytes, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
  if (e != cudaSuccess) {
    lk_set_error("cudaHostAlloc(%lld): %s", (long long)bytes, cudaGetErrorString(e));
    return LK_ECUDA;
  }
  return LK_OK;
}
extern "C" int lk_host_free(void* p) {
  if (!p) return LK_OK;
  cudaError_t e = cudaFreeHost(p);
  if (e != cudaSuccess) {
    lk_set_error("cudaFreeHost: %s", cudaGetErrorString(e));
    return LK_ECUDA;
  }
  return LK_OK;
}

// programmatic dependent launch on/off (LINKB200_PDL=0 restores ordinary launches; common.cuh)
bool lk_pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("LINKB200_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}
bool lk_pdl_enabled_link() {
  static const bool on = [] {
    const char* e = getenv("LINKB200_PDL_LINK");
    return lk_pdl_enabled() && e && e[0] == '1';
  }();
  return on;
}
