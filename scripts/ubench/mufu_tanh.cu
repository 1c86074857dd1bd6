// Micro-benchmark: MUFU.TANH throughput per SM for f32, f16x2 and bf16x2 operands (elements per clock per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_tanh scripts/ubench/mufu_tanh.cu && /tmp/mufu_tanh
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(uint32_t* out, long long* cyc, int iters) {
  uint32_t a[8];
  for (int i = 0; i < 8; ++i) a[i] = 0x3c003800u + threadIdx.x * 17 + i * 3;      // fp16 / bf16-ish bit patterns; f32 small numbers
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+r"(a[i]));
      if (MODE == 1) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(a[i]));
      if (MODE == 2) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(a[i]));
    }
  }
  const long long t1 = clock64();
  uint32_t s = 0;
  for (int i = 0; i < 8; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  for (int warps : {4, 8, 16, 32}) {
    for (int mode = 0; mode < 3; ++mode) {
      if (mode == 0) k<0><<<148, warps * 32>>>(out, cyc, iters);
      if (mode == 1) k<1><<<148, warps * 32>>>(out, cyc, iters);
      if (mode == 2) k<2><<<148, warps * 32>>>(out, cyc, iters);
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      const double ops = (double)iters * 8 * warps * 32;             // MUFU lane-ops per CTA
      const double el = ops * (mode == 0 ? 1 : 2);
      printf("warps=%2d mode=%s: %.1f lane-ops/clk/SM, %.1f elements/clk/SM\n", warps, mode == 0 ? "f32   " : (mode == 1 ? "f16x2 " : "bf16x2"),
             ops / h[0], el / h[0]);
    }
  }
  return 0;
}
