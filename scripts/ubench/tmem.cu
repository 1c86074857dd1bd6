// Dev micro-benchmark: tcgen05.ld (TMEM -> registers) throughput per SM, alone and overlapped with MUFU work.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
// MODE 0: ld;wait back to back | 1: two lds in flight then wait | 2: ld(next) ; 32 MUFU on current ; wait | 3: only 32 MUFU per iter
template <int MODE>
__global__ void k(float* out, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
  uint32_t va[32], vb[32];
  float acc = 0.f;
  for (int i = 0; i < 32; ++i) { va[i] = 0x3f000000u + i; vb[i] = 0x3f100000u + i; }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t col = (it & 3) * 64;
    if (MODE == 0) {
      tmem_ld32(base + col, va); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tmem_ld32(base + col + 32, vb); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    } else if (MODE == 1) {
      tmem_ld32(base + col, va); tmem_ld32(base + col + 32, vb);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    } else if (MODE == 2) {
      tmem_ld32(base + col, vb);
#pragma unroll
      for (int i = 0; i < 32; ++i) { float f = __uint_as_float(va[i]); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(f)); acc += f; }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tmem_ld32(base + col + 32, va);
#pragma unroll
      for (int i = 0; i < 32; ++i) { float f = __uint_as_float(vb[i]); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(f)); acc += f; }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    } else {
#pragma unroll
      for (int i = 0; i < 64; ++i) { float f = __uint_as_float(va[i & 31]); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(f)); va[i & 31] = __float_as_uint(f); }
    }
    acc += __uint_as_float(va[3]) + __uint_as_float(vb[7]);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}
int main() {
  float* d; cudaMalloc(&d, 1 << 20);
  const char* names[] = {"ld;wait x2", "2 ld in flight", "ld overlapped with 32 MUFU", "64 MUFU only"};
  for (int warps = 4; warps <= 8; warps *= 2)
    for (int mode = 0; mode < 4; ++mode) {
      int iters = 2000; float h;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(d, iters);
        if (mode == 1) k<1><<<148, warps * 32>>>(d, iters);
        if (mode == 2) k<2><<<148, warps * 32>>>(d, iters);
        if (mode == 3) k<3><<<148, warps * 32>>>(d, iters);
        cudaDeviceSynchronize();
      }
      cudaError_t e = cudaGetLastError();
      cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
      double bytes = (double)iters * 2 * 4096 * warps;
      printf("%-28s warps/SM=%d cycles/iter(2 x 4KB ld per warp)=%7.1f  TMEM B/clk/SM=%6.1f  %s\n", names[mode], warps, h / iters, mode < 3 ? bytes / h : 0.0, cudaGetErrorString(e));
    }
  return 0;
}
