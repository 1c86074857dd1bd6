// K3c (bf16) — the three true GEMMs of the PointNet head on the 5th-gen tensor cores (tcgen05, cta_group::1).
// reference: algorithms/algo_utils/network.py:152-159 (final_mlp) + autograd.
//
// The fused FFMA head (head.cu) is latency-bound: 10 % of the bf16-mode iteration for 1 % of its FLOPs.  Here the three
// products with a 512-wide dimension run as tcgen05 MMAs, operands converted fp32 -> bf16 while they are staged into
// SWIZZLE_128B shared-memory blocks.  One staging routine serves every operand: it copies a [lines x 64] block whose
// 64-element side is contiguous in global memory; the SAME shared-memory image is then described to the tensor core as
// K-major (lines = M/N index) or MN-major (lines = K index) — no transposes anywhere:
//   fwd    h1 = act(feat . W0^T + b0)          A = feat rows   (K-major)   B = W0 rows   (K-major)    M=128 rows,  N=128, K=F
//          h2 = act(h1 . W1^T + b1)            A = h1 (bf16, written by the epilogue)  B = W1 rows (K-major)       N=32, K=128
//          out = h2 . W2^T + b2                CUDA cores (32 x out MACs per row)
//   dfeat  = dPre1 . W0                        A = dPre1 rows  (K-major)   B = W0 rows   (MN-major)   M=128 rows,  N=256, K=128
//   dW0    = dPre1^T . feat (split over rows)  A = dPre1 rows  (MN-major)  B = feat rows (MN-major)   M=128 (ch),  N=256, K=128 rows
// dPre2 / dPre1 / the small gradients stay in head.cu's row-block kernel (launched without its dfeat loop); the dW0
// partials go through head.cu's fixed-order reduce.  Opt-in (`head_precision: bf16`, 1e-2 gate): at B = 2048 the grids are
// 16-32 CTAs whose operand staging is latency-bound, so the whole iteration measured no faster than with the FFMA head
// (194.9 vs 193.0 ms); the kernels are parity-tested and kept as the base for a split-K / pipelined version.
#include "tc_common.cuh"

namespace {
using namespace pmtc;

constexpr int GT = 256;                 // threads: all stage operands; warps 0-3 run the epilogue (thread = TMEM lane); thread 0 issues
constexpr uint32_t BLK = 16384;         // one SW128 block of 128 lines x 128 B

// copy a [LINES x 64] fp32 block (64-element side contiguous, row stride ld) into one SW128 block as bf16;
// lines >= valid_lines and columns >= valid_cols are zero-filled
template <int LINES>
__device__ __forceinline__ void stage_block(uint8_t* dst, const float* __restrict__ src, int64_t ld, int valid_lines,
                                            int valid_cols, int tid) {
  for (int i = tid; i < LINES * 8; i += GT) {
    const int line = i >> 3, c8 = i & 7;
    float v[8];
    const float* p = src + (int64_t)line * ld + c8 * 8;
    if (line < valid_lines && c8 * 8 + 8 <= valid_cols && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (line < valid_lines && c8 * 8 + j < valid_cols) ? __ldg(p + j) : 0.f;
    }
    *reinterpret_cast<uint4*>(dst + line * 128 + ((c8 ^ (line & 7)) << 4)) =
        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
}

__device__ __forceinline__ uint32_t tmem_alloc_256(uint32_t* slot, int warp) {
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *slot;
}
__device__ __forceinline__ void tmem_free_256(uint32_t base, int warp) {
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(256u) : "memory");
}

// ------------------------------------------------------------------------------------------------ forward
// smem: A stages 2 x 16 KB | B stages 2 x 16 KB | H1 bf16 2 blocks (32 KB) | W1 bf16 2 k-blocks of 32 lines (8 KB) | W2 fp32 | bars
constexpr uint32_t F_A = 0, F_B = 32768, F_H1 = 65536, F_W1 = 98304, F_W2 = 106496, F_BAR = 110592, F_TOTAL = 110592 + 64;

template <int ACT>
__global__ void __launch_bounds__(GT, 1)
head_fwd_tc_kernel(const float* __restrict__ feat, int64_t ldf, int B, int F, pm_head_params P, int out_dim,
                   float* __restrict__ h1, float* __restrict__ h2, float* __restrict__ out, int64_t ldo, ErrSink err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row0 = blockIdx.x * 128;
  const int rows = min(128, B - row0);
  auto bar = [&](int i) { return sbase + F_BAR + 8u * i; };             // 0,1: stage free | 2: L0 done | 3: L1 done
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + F_BAR + 48);
  float* sW2 = reinterpret_cast<float*>(smem + F_W2);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(bar(i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // W1 (32 x 128) as a K-major B operand: k-block kb = 32 lines of 128 B (4 KB)
  for (int i = tid; i < 2 * 32 * 8; i += GT) {
    const int c8 = i & 7, line = (i >> 3) & 31, kb = i >> 8;
    const float* p = P.W1 + line * 128 + kb * 64 + c8 * 8;
    *reinterpret_cast<uint4*>(smem + F_W1 + kb * 4096 + line * 128 + ((c8 ^ (line & 7)) << 4)) =
        make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7]));
  }
  for (int i = tid; i < out_dim * 32; i += GT) sW2[i] = P.W2[i];
  const uint32_t tmem = tmem_alloc_256(slot, warp);
  bool ok = true;
  const uint32_t idesc0 = umma_idesc_ex(128, 128, 0, 0), idesc1 = umma_idesc_ex(128, 32, 0, 0);
  // ---- layer 0: K loop in 64-element chunks, two smem stages, MMAs asynchronous behind the staging of the next chunk
  const int n_chunks = (F + 63) / 64;
  for (int c = 0; c < n_chunks && ok; ++c) {
    const int s = c & 1;
    if (c >= 2) ok = mbar_wait(bar(s), ((c >> 1) - 1) & 1, err, 301);   // MMAs of chunk c-2 are done with this stage
    if (!ok) break;
    const int kc = min(64, F - c * 64);
    stage_block<128>(smem + F_A + s * BLK, feat + (int64_t)row0 * ldf + c * 64, ldf, rows, kc, tid);
    stage_block<128>(smem + F_B + s * BLK, P.W0 + c * 64, F, 128, kc, tid);
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16_1cta(tmem, umma_desc(sbase + F_A + s * BLK + k * 32), umma_desc(sbase + F_B + s * BLK + k * 32), idesc0,
                       (c > 0 || k > 0) ? 1u : 0u);
      umma_commit_1cta(bar(s));
      if (c == n_chunks - 1) umma_commit_1cta(bar(2));
    }
  }
  // ---- epilogue 0 (warps 0-3, thread = row): h1 = act(acc + b0) -> global fp32 + smem bf16 (A operand of layer 1)
  if (ok) ok = mbar_wait(bar(2), 0, err, 302);
  tc_fence_after();
  const int r = tid & 127;
  const uint32_t lane_taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  if (warp < 4 && ok) {
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      uint32_t v[32];
      tmem_ld32(lane_taddr + cc * 32, v);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a0 = pm_act_fwd(ACT, __uint_as_float(v[2 * i]) + __ldg(P.b0 + cc * 32 + 2 * i));
        const float a1 = pm_act_fwd(ACT, __uint_as_float(v[2 * i + 1]) + __ldg(P.b0 + cc * 32 + 2 * i + 1));
        v[2 * i] = __float_as_uint(a0); v[2 * i + 1] = __float_as_uint(a1);
        pk[i] = pack_bf16(a0, a1);
      }
      if (r < rows) {
        float4* dst = reinterpret_cast<float4*>(h1 + (int64_t)(row0 + r) * 128 + cc * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(smem + F_H1 + (cc >> 1) * BLK + sw128(r, (cc & 1) * 4 + q)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
    }
  }
  tc_fence_before();
  fence_proxy_async();
  __syncthreads();
  // ---- layer 1 on the tensor core: acc1[128 x 32] = h1 . W1^T (TMEM columns [128,160))
  if (tid == 0 && ok) {
    tc_fence_after();
#pragma unroll
    for (int k = 0; k < 8; ++k)
      umma_bf16_1cta(tmem + 128, umma_desc(sbase + F_H1 + (k >> 2) * BLK + (k & 3) * 32), umma_desc(sbase + F_W1 + (k >> 2) * 4096 + (k & 3) * 32),
                     idesc1, k > 0);
    umma_commit_1cta(bar(3));
  }
  if (ok) ok = mbar_wait(bar(3), 0, err, 303);
  tc_fence_after();
  if (warp < 4 && ok) {
    uint32_t v[32];
    tmem_ld32(lane_taddr + 128, v);
    tmem_ld_wait();
    float hv[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) hv[i] = pm_act_fwd(ACT, __uint_as_float(v[i]) + __ldg(P.b1 + i));
    if (r < rows) {
      float4* dst = reinterpret_cast<float4*>(h2 + (int64_t)(row0 + r) * 32);
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[i] = make_float4(hv[4 * i], hv[4 * i + 1], hv[4 * i + 2], hv[4 * i + 3]);
      for (int o = 0; o < out_dim; ++o) {                     // layer 2 (no activation)
        float a = __ldg(P.b2 + o);
#pragma unroll
        for (int k = 0; k < 32; ++k) a = fmaf(sW2[o * 32 + k], hv[k], a);
        out[(int64_t)(row0 + r) * ldo + o] = a;
      }
    }
  }
  tmem_free_256(tmem, warp);
}

// ------------------------------------------------------------------------------------------------ dfeat = dPre1 . W0
// grid (row tiles of 128, column tiles of 256).  smem: A = dPre1 tile, 2 k-blocks (32 KB) | B = W0[:, n0:n0+256] as 4 MN blocks
// of 128 K-lines (64 KB) | bars
constexpr uint32_t D_A = 0, D_B = 32768, D_BAR = 98304, D_TOTAL = 98304 + 64;

__global__ void __launch_bounds__(GT, 1)
head_dfeat_tc_kernel(const float* __restrict__ dpre1, int B, const float* __restrict__ W0, int F, float* __restrict__ dfeat,
                     int64_t lddf, int dfeat_cols, ErrSink err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row0 = blockIdx.x * 128, n0 = blockIdx.y * 256;
  const int rows = min(128, B - row0), ncols = min(256, dfeat_cols - n0);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + D_BAR + 16);
  if (tid == 0) {
    mbar_init(sbase + D_BAR, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t tmem = tmem_alloc_256(slot, warp);
  for (int kb = 0; kb < 2; ++kb) stage_block<128>(smem + D_A + kb * BLK, dpre1 + (int64_t)row0 * 128 + kb * 64, 128, rows, 64, tid);
  for (int nb = 0; nb < 4; ++nb)                              // MN block nb: lines = K index (W0 row j), 64 columns n0 + nb*64 ..
    stage_block<128>(smem + D_B + nb * BLK, W0 + n0 + nb * 64, F, 128, min(64, max(0, min(F, dfeat_cols) - n0 - nb * 64)), tid);
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = umma_idesc_ex(128, 256, 0, 1);
#pragma unroll
    for (int k = 0; k < 8; ++k)                                // K = 128 (dPre1 columns / W0 rows)
      umma_bf16_1cta(tmem, umma_desc(sbase + D_A + (k >> 2) * BLK + (k & 3) * 32), umma_desc_mn(sbase + D_B + k * 2048, BLK, 1024), idesc, k > 0);
    umma_commit_1cta(sbase + D_BAR);
  }
  const bool ok = mbar_wait(sbase + D_BAR, 0, err, 311);
  tc_fence_after();
  if (warp < 4 && ok) {
    const int r = tid & 127;
    const uint32_t lane_taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll 1
    for (int cc = 0; cc < 8; ++cc) {
      uint32_t v[32];
      tmem_ld32(lane_taddr + cc * 32, v);
      tmem_ld_wait();
      if (r < rows) {
        float* dst = dfeat + (int64_t)(row0 + r) * lddf + n0 + cc * 32;
        if (cc * 32 + 32 <= ncols && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            reinterpret_cast<float4*>(dst)[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (cc * 32 + i < ncols) dst[i] = __uint_as_float(v[i]);
        }
      }
    }
  }
  tmem_free_256(tmem, warp);
}

// ------------------------------------------------------------------------------------------------ dW0 partials = dPre1^T . feat
// grid (column tiles of 256, row slabs of 128).  smem: A = dPre1 slab as 2 MN blocks (32 KB) | B = feat slab as 4 MN blocks (64 KB)
__global__ void __launch_bounds__(GT, 1)
head_dw0_tc_kernel(const float* __restrict__ dpre1, const float* __restrict__ feat, int64_t ldf, int B, int F,
                   float* __restrict__ part, ErrSink err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n0 = blockIdx.x * 256, slab = blockIdx.y, row0 = slab * 128;
  const int rows = min(128, B - row0), ncols = min(256, F - n0);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + D_BAR + 16);
  if (tid == 0) {
    mbar_init(sbase + D_BAR, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t tmem = tmem_alloc_256(slot, warp);
  for (int mb = 0; mb < 2; ++mb) stage_block<128>(smem + D_A + mb * BLK, dpre1 + (int64_t)row0 * 128 + mb * 64, 128, rows, 64, tid);
  for (int nb = 0; nb < 4; ++nb)
    stage_block<128>(smem + D_B + nb * BLK, feat + (int64_t)row0 * ldf + n0 + nb * 64, ldf, rows, min(64, max(0, F - n0 - nb * 64)), tid);
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = umma_idesc_ex(128, 256, 1, 1);
#pragma unroll
    for (int k = 0; k < 8; ++k)                                // K = 128 batch rows of this slab
      umma_bf16_1cta(tmem, umma_desc_mn(sbase + D_A + k * 2048, BLK, 1024), umma_desc_mn(sbase + D_B + k * 2048, BLK, 1024), idesc, k > 0);
    umma_commit_1cta(sbase + D_BAR);
  }
  const bool ok = mbar_wait(sbase + D_BAR, 0, err, 321);
  tc_fence_after();
  if (warp < 4 && ok) {
    const int n = tid & 127;                                    // TMEM lane == dW0 row (h1 channel)
    const uint32_t lane_taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float* dst = part + ((size_t)slab * 128 + n) * F + n0;
#pragma unroll 1
    for (int cc = 0; cc < 8; ++cc) {
      uint32_t v[32];
      tmem_ld32(lane_taddr + cc * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (cc * 32 + i < ncols) dst[cc * 32 + i] = __uint_as_float(v[i]);
    }
  }
  tmem_free_256(tmem, warp);
}

}  // namespace

extern "C" {

// launched by head.cu's entry points when precision == PM_PREC_BF16
int pm_head_fwd_tc_launch(const float* feat, int64_t ldf, int B, int F, const pm_head_params* p, int out_dim, int act, float* h1,
                          float* h2, float* out, int64_t ldo, int32_t* err_word, cudaStream_t st) {
  const ErrSink err{err_word, pm_tc_sticky_word()};
  const dim3 grid(pm_cdiv(B, 128));
#define PM_HTF(ACTV)                                                                                                     \
  case ACTV: {                                                                                                           \
    static bool attr_set = false;                                                                                        \
    if (!attr_set) {                                                                                                     \
      cudaError_t e = cudaFuncSetAttribute(head_fwd_tc_kernel<ACTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F_TOTAL); \
      if (e != cudaSuccess) PM_FAIL(PM_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));                     \
      attr_set = true;                                                                                                   \
    }                                                                                                                    \
    head_fwd_tc_kernel<ACTV><<<grid, GT, F_TOTAL, st>>>(feat, ldf, B, F, *p, out_dim, h1, h2, out, ldo, err);            \
  } break;
  switch (act) {
    PM_HTF(PM_ACT_NONE) PM_HTF(PM_ACT_TANH) PM_HTF(PM_ACT_RELU) PM_HTF(PM_ACT_ELU) PM_HTF(PM_ACT_SELU) PM_HTF(PM_ACT_LRELU)
    PM_HTF(PM_ACT_SIGMOID)
    default: PM_FAIL(PM_ERR_ARG, "pm_pointnet_head_forward: activation %d", act);
  }
#undef PM_HTF
  return PM_OK;
}

int pm_head_bwd_tc_launch(const float* dpre1, const float* feat, int64_t ldf, int B, int F, const float* W0, float* dfeat,
                          int64_t lddf, int dfeat_cols, float* partB, int32_t* err_word, cudaStream_t st) {
  const ErrSink err{err_word, pm_tc_sticky_word()};
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e1 = cudaFuncSetAttribute(head_dfeat_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D_TOTAL);
    cudaError_t e2 = cudaFuncSetAttribute(head_dw0_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D_TOTAL);
    if (e1 != cudaSuccess || e2 != cudaSuccess) PM_FAIL(PM_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    attr_set = true;
  }
  if (dfeat) head_dfeat_tc_kernel<<<dim3(pm_cdiv(B, 128), pm_cdiv(dfeat_cols, 256)), GT, D_TOTAL, st>>>(dpre1, B, W0, F, dfeat, lddf, dfeat_cols, err);
  head_dw0_tc_kernel<<<dim3(pm_cdiv(F, 256), pm_cdiv(B, 128)), GT, D_TOTAL, st>>>(dpre1, feat, ldf, B, F, partB, err);
  return PM_OK;
}

}  // extern "C"
