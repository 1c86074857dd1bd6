// Dev micro-benchmark: per-SM throughput of MUFU.TANH / MUFU.EX2 / MUFU.RCP / F2FP / FFMA with 1 or 2 warps per SMSP.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, int iters) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 0.001f + i * 0.1f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 3) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(r) : "f"(v[i])); v[i] = __uint_as_float(r); }
      if (OP == 4) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[i]));
      if (OP == 5) { unsigned r = __float_as_uint(v[i]); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(r)); v[i] = __uint_as_float(r); }
      if (OP == 6) { unsigned r = __float_as_uint(v[i]); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(r)); v[i] = __uint_as_float(r); }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}
int main() {
  float* d; cudaMalloc(&d, 1 << 20);
  const char* names[] = {"MUFU.TANH", "MUFU.EX2", "MUFU.RCP", "F2FP.BF16x2", "FFMA", "TANH.BF16x2 (instr)", "TANH.F16x2 (instr)"};
  for (int warps = 4; warps <= 16; warps *= 2)
    for (int op = 0; op < 7; ++op) {
      int iters = 2000;
      float h;
      for (int rep = 0; rep < 2; ++rep) {
        if (op == 0) k<0><<<148, warps * 32>>>(d, iters);
        if (op == 1) k<1><<<148, warps * 32>>>(d, iters);
        if (op == 2) k<2><<<148, warps * 32>>>(d, iters);
        if (op == 3) k<3><<<148, warps * 32>>>(d, iters);
        if (op == 4) k<4><<<148, warps * 32>>>(d, iters);
        if (op == 5) k<5><<<148, warps * 32>>>(d, iters);
        if (op == 6) k<6><<<148, warps * 32>>>(d, iters);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
      double ops = (double)iters * 8 * warps * 32;
      printf("%-12s warps/SM=%2d  cycles=%8.0f  lanes/clk/SM=%6.2f\n", names[op], warps, h, ops / h);
    }
  return 0;
}
