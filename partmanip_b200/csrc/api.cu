// Library-level entry points of libpartmanip_b200.so.
#include "common.cuh"

char g_pm_err[512] = {0};

// Sticky protocol-error word of the tcgen05 kernels, one per device, owned by the library and never cleared by a launch
// (tc_common.cuh:ErrSink).  Allocated on the first tcgen05 launch of a device — an eager one: cudaMalloc is illegal inside a
// stream capture, and every caller warms up eagerly before capturing.
static int32_t* g_sticky[64] = {nullptr};
int32_t* pm_tc_sticky_word() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!g_sticky[dev]) {
    int32_t* p = nullptr;
    if (cudaMalloc(&p, 256) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    cudaMemset(p, 0, 256);
    g_sticky[dev] = p;
  }
  return g_sticky[dev];
}

extern "C" {
// first protocol error any tcgen05 kernel reported on the current device since the last clear (0 = none); synchronises the device
int pm_tc_sticky_error(int clear) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || !g_sticky[dev]) return 0;
  int32_t h = 0;
  if (cudaMemcpy(&h, g_sticky[dev], sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  if (clear && h) cudaMemset(g_sticky[dev], 0, sizeof(h));
  return h;
}
const char* pm_last_error(void) { return g_pm_err; }
int pm_version(void) { return 100; }
}
