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

// SMs of the current device (cached per device).  Grids and workspaces are SIZED for PM_NUM_SMS = 148 (B200); kernels whose
// correctness depends on co-residency (the grid barrier of pm_fused_step) additionally cap their grid at the real count.
int pm_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return PM_NUM_SMS;
  if (!cached[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = PM_NUM_SMS;
    cached[dev] = n;
  }
  return cached[dev];
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
