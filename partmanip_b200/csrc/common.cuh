// Shared device/host helpers for libpartmanip_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/partmanip_b200.h"

extern char g_pm_err[512];

#define PM_FAIL(code, ...)                                   \
  do {                                                       \
    snprintf(g_pm_err, sizeof(g_pm_err), __VA_ARGS__);       \
    return (code);                                           \
  } while (0)

#define PM_CHECK_LAUNCH(name)                                                          \
  do {                                                                                 \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess) PM_FAIL(PM_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e__)); \
  } while (0)

#define PM_REQUIRE(cond, code, ...) \
  do {                              \
    if (!(cond)) PM_FAIL(code, __VA_ARGS__); \
  } while (0)

static inline cudaStream_t pm_st(pm_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int pm_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t pm_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline bool pm_aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

constexpr int PM_NUM_SMS = 148;      // B200; sizing constant for grids / workspaces (pm_sm_count() in api.cu is the runtime value)

// ---------------------------------------------------------------- activations (network.py:7-24)
__device__ __forceinline__ float pm_act_fwd(int act, float x) {
  switch (act) {
    case PM_ACT_TANH: return tanhf(x);
    case PM_ACT_RELU: return x > 0.f ? x : 0.f;
    case PM_ACT_ELU: return x > 0.f ? x : expm1f(x);
    case PM_ACT_SELU: {
      const float a = 1.6732632423543772848170429916717f, l = 1.0507009873554804934193349852946f;
      return l * (x > 0.f ? x : a * expm1f(x));
    }
    case PM_ACT_LRELU: return x > 0.f ? x : 0.01f * x;
    case PM_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    default: return x;
  }
}
// tanh(x) = 1 - 2 / (exp(2x) + 1) on the MUFU pipe (ex2.approx + rcp.approx): absolute error <= 4e-7, ~8 instructions instead of
// tanhf's ~40; the other activations as above.  Used where fp32-mode kernels recompute activations (1e-4 parity gate).
__device__ __forceinline__ float pm_act_fwd_fast(int act, float x) {
  if (act == PM_ACT_TANH) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
    return fmaf(-2.f, r, 1.f);
  }
  return pm_act_fwd(act, x);
}
// derivative expressed through the OUTPUT y = act(x) (only outputs are kept / recomputed)
__device__ __forceinline__ float pm_act_bwd(int act, float y) {
  switch (act) {
    case PM_ACT_TANH: return 1.f - y * y;
    case PM_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case PM_ACT_ELU: return y > 0.f ? 1.f : y + 1.f;
    case PM_ACT_SELU: {
      const float a = 1.6732632423543772848170429916717f, l = 1.0507009873554804934193349852946f;
      return y > 0.f ? l : y + l * a;
    }
    case PM_ACT_LRELU: return y > 0.f ? 1.f : 0.01f;
    case PM_ACT_SIGMOID: return y * (1.f - y);
    default: return 1.f;
  }
}

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ float pm_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double pm_warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum, result valid in every thread; blockDim.x multiple of 32, <= 1024
__device__ __forceinline__ float pm_block_sum(float v, float* sm /* >= 32 floats */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = pm_warp_sum(v);
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  float r = (lane < nw) ? sm[lane] : 0.f;
  r = pm_warp_sum(r);
  return r;
}
__device__ __forceinline__ double pm_block_sum_d(double v, double* sm /* >= 32 doubles */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = pm_warp_sum_d(v);
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  double r = (lane < nw) ? sm[lane] : 0.0;
  r = pm_warp_sum_d(r);
  return r;
}
