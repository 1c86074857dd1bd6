// K3t — dense layers on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM) for fp32 tensors in HBM.
// reference: nn.Linear forward / autograd inside MLP (algorithms/algo_utils/network.py:27-54: the shipped state policy
// 53 -> 512^3 -> 10 of cfg/algos/ppo.yaml:44-47, also DAgger's teacher) and the dense layers of the critical-point encoder
// backward (pointnet.cu).  Same three products as dense.cu, selected through the operands' majors:
//   FWD  Y[m,n]  = act(sum_k X[m,k] W[n,k] + b[n])              A = X  K-major,  B = W  K-major
//   DX   dX[m,k] = (sum_n dY[m,n] W[n,k]) * act'(Xprev[m,k])    A = dY K-major,  B = W  MN-major
//   DW   dW[n,k] = sum_m dY[m,n] X[m,k]                          A = dY MN-major, B = X  MN-major   (split over m)
//
// Operands stay fp32 in global memory; a CTA converts each [128 x 64] tile on the fly into bf16 images in shared memory
// (SWIZZLE_128B: rows of 64 contiguous elements, 16-byte chunk index XOR (row & 7)) — the SAME physical image serves a
// K-major read (rows = M/N index) and an MN-major read (rows = contraction index), so no transposes are ever staged.
// Two arithmetic modes:
//   PARTS = 1  bf16 operands (1e-2 gate)
//   PARTS = 3  each fp32 value split into three bf16 terms a = a1 + a2 + a3 (8 + 8 + 8 mantissa bits, fp32 exponent range —
//              gradients as small as 1e-30 keep full relative precision, which an fp16 split would lose to underflow) and
//              the product formed from the six term pairs above 2^-24: (1,1) (1,2) (2,1) (2,2) (1,3) (3,1), accumulated in
//              fp32 in TMEM.  Error ~ fp32 rounding; gate 1e-4.
// One 128 x 128 output tile per CTA, 64-deep k-blocks, two shared-memory stages: while the tensor core works on stage s the
// CTA's 256 threads convert and store stage s^1 and already hold the global loads of the block after that in registers.
#include "tc_common.cuh"

int pm_sm_count();                               // api.cu

namespace {
using namespace pmtc;

constexpr int GT_THREADS = 256;
constexpr int TBM = 128, TBN = 128, TBK = 64;
constexpr uint32_t IMG_BYTES = 16384;            // one bf16 image of a [128 x 64] tile (K-major) = two [64 x 64] blocks (MN-major)
constexpr int CHUNKS_PER_THREAD = (TBM * TBK / 8) / GT_THREADS;   // 4 chunks of 8 elements per operand per thread

enum { EPI_FWD = 0, EPI_DX = 1, EPI_DW = 2 };

struct GemmTcP {
  const float* A; int64_t lda; int a_mn;         // a_mn = 0: A[M x K] row-major (K contiguous) | 1: A[K x M] row-major
  const float* B; int64_t ldb; int b_mn;         // b_mn = 0: B[N x K] row-major (K contiguous) | 1: B[K x N] row-major
  float* C; int64_t ldc;                         // C[M x N] row-major (DW: + split * M * ldc)
  int M, N, K;
  const float* bias;                             // FWD, [N] or null
  const float* aux; int64_t ldaux;               // DX: previous layer's activation OUTPUT [M x N]
  int act;
  const int32_t* lim_dev;                        // FWD/DX: valid rows (<= M); DW: valid contraction length (<= K); or null
  int k_per_split;                               // DW
  ErrSink err;
};

// ---- one operand tile: global fp32 -> registers (prefetch) -> bf16 parts in shared memory
struct TileRegs {
  float v[CHUNKS_PER_THREAD][8];
};

// chunk i of the tile whose MN origin is mn0 and contraction origin k0; limits are exclusive upper bounds
template <bool MN_MAJOR>
__device__ __forceinline__ void load_tile(TileRegs& r, const float* __restrict__ g, int64_t ld, int mn0, int k0, int mn_lim,
                                          int k_lim, int tid) {
#pragma unroll
  for (int j = 0; j < CHUNKS_PER_THREAD; ++j) {
    const int i = tid + j * GT_THREADS;
    int row, e0, row_lim, e_lim;               // row: index along the strided axis; e0: first element along the contiguous axis
    if (MN_MAJOR) { row = k0 + (i >> 4); e0 = mn0 + (i & 15) * 8; row_lim = k_lim; e_lim = mn_lim; }
    else          { row = mn0 + (i >> 3); e0 = k0 + (i & 7) * 8;  row_lim = mn_lim; e_lim = k_lim; }
    const float* p = g + (int64_t)row * ld + e0;
    if (row < row_lim && e0 + 8 <= e_lim && ((reinterpret_cast<uintptr_t>(p) & 15u) == 0)) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p));
      const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
      r.v[j][0] = a.x; r.v[j][1] = a.y; r.v[j][2] = a.z; r.v[j][3] = a.w;
      r.v[j][4] = b.x; r.v[j][5] = b.y; r.v[j][6] = b.z; r.v[j][7] = b.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) r.v[j][e] = (row < row_lim && e0 + e < e_lim) ? __ldg(p + e) : 0.f;
    }
  }
}

__device__ __forceinline__ float bf16_hi_as_f32(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ float bf16_lo_as_f32(uint32_t w) { return __uint_as_float(w << 16); }

// mn_used: the MMA reads only the first mn_used (multiple of 16) MN indices of the tile: chunks beyond are not stored
template <bool MN_MAJOR, int PARTS>
__device__ __forceinline__ void store_tile(const TileRegs& r, uint8_t* base /* PARTS images, IMG_BYTES apart */, int tid, int mn_used = 128) {
#pragma unroll
  for (int j = 0; j < CHUNKS_PER_THREAD; ++j) {
    const int i = tid + j * GT_THREADS;
    uint32_t off;
    if (MN_MAJOR) { const uint32_t kr = i >> 4, c16 = i & 15; off = (c16 >> 3) * 8192u + kr * 128u + (((c16 & 7u) ^ (kr & 7u)) << 4);
                    if ((int)(c16 * 8) >= mn_used) continue; }
    else          { const uint32_t row = i >> 3, c = i & 7;   off = row * 128u + ((c ^ (row & 7u)) << 4);
                    if ((int)row >= mn_used) continue; }
    float res[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) res[e] = r.v[j][e];
#pragma unroll
    for (int part = 0; part < PARTS; ++part) {
      uint32_t w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        w[q] = pack_bf16(res[2 * q], res[2 * q + 1]);
        if (part + 1 < PARTS) {                 // residual for the next term (exact in fp32)
          res[2 * q] -= bf16_lo_as_f32(w[q]);
          res[2 * q + 1] -= bf16_hi_as_f32(w[q]);
        }
      }
      *reinterpret_cast<uint4*>(base + part * IMG_BYTES + off) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

// PARTS = 1: two 32 KB stages (three CTAs per SM).  PARTS = 3: ONE 96 KB stage, so that two CTAs share an SM and overlap each
// other's load / convert / MMA / epilogue phases — measured on the 250k x 256 x 128 layers of the encoder backward, where every
// CTA has only two k-blocks and an in-CTA pipeline never fills: 633 us with two stages and one CTA per SM.
template <int PARTS> struct SmemPlan {
  static constexpr int NSTG = PARTS == 1 ? 2 : 1;
  static constexpr uint32_t STAGE = 2u * PARTS * IMG_BYTES;       // A parts then B parts
  static constexpr uint32_t BAR = NSTG * STAGE;                   // NSTG + 1 mbarriers + tmem slot
  static constexpr uint32_t TOTAL = BAR + 64;
};

__device__ __forceinline__ float act_fwd_fast(int act, float x) { return pm_act_fwd_fast(act, x); }

template <int PARTS, int EPI, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GT_THREADS, 2)
gemm_tc_kernel(const GemmTcP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using Plan = SmemPlan<PARTS>;
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.y * TBN;
  int M = p.M, K = p.K;
  int k_begin = 0, k_end = K;
  if (EPI == EPI_DW) {
    int kps = p.k_per_split;
    if (p.lim_dev) {                             // spread the VALID contraction rows over all splits
      K = min(K, *p.lim_dev);
      kps = ((K + (int)gridDim.z - 1) / (int)gridDim.z + TBK - 1) / TBK * TBK;
    }
    k_begin = min(K, (int)blockIdx.z * kps);
    k_end = min(K, k_begin + kps);
  } else {
    if (p.lim_dev) M = min(M, *p.lim_dev);
    if ((int)blockIdx.x * TBM >= M) return;      // uniform per CTA: before any barrier / TMEM allocation
  }
  // M tiles are dealt round-robin to the CTAs of a column (FWD / DX launch at most two CTAs per SM and walk the tiles; DW has one)
  const int n_mtiles = (M + TBM - 1) / TBM;
  const int nkb = (k_end - k_begin + TBK - 1) / TBK;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Plan::BAR + 32);
  constexpr int NSTG = Plan::NSTG;
  auto bar = [&](int i) { return sbase + Plan::BAR + 8u * i; };    // 0..NSTG-1: stage free | NSTG: accumulator complete

  if ((sbase & 1023u) != 0 && tid == 0) err_report(p.err, 910);
  if (tid == 0) {
    for (int i = 0; i <= NSTG; ++i) mbar_init(bar(i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // the MMA is as wide as this CTA's live columns, in steps of 16 (a 16-wide output layer costs an eighth of the tensor time of N = 128)
  const int n_used = min(TBN, (p.N - n0 + 15) / 16 * 16);
  const uint32_t idesc = umma_idesc_ex(TBM, n_used, A_MN ? 1 : 0, B_MN ? 1 : 0);
  bool ok = true;

  TileRegs ra, rb;
  if (nkb > 0) {
    load_tile<A_MN>(ra, p.A, p.lda, (int)blockIdx.x * TBM, k_begin, M, k_end, tid);
    load_tile<B_MN>(rb, p.B, p.ldb, n0, k_begin, p.N, k_end, tid);
  }
  uint32_t uses[NSTG];                           // commits issued so far on each stage barrier (uniform across the CTA)
#pragma unroll
  for (int i = 0; i < NSTG; ++i) uses[i] = 0;
  uint32_t tiles_done = 0;
  for (int mt = blockIdx.x; mt < n_mtiles && ok; mt += gridDim.x) {
  const int m0 = mt * TBM;
  const int mt_next = mt + (int)gridDim.x;
  for (int kb = 0; kb < nkb && ok; ++kb) {
    const int s = kb % NSTG;
    if (uses[s] > 0) {                           // the MMAs that last read this stage are done
      ok = mbar_wait(bar(s), (uses[s] - 1) & 1, p.err, 911);
      if (!ok) break;
    }
    uint8_t* st = smem + s * Plan::STAGE;
    store_tile<A_MN, PARTS>(ra, st, tid);
    store_tile<B_MN, PARTS>(rb, st + PARTS * IMG_BYTES, tid, n_used);
    if (kb + 1 < nkb) {                          // next block's global loads fly during the barrier and the MMA issue
      load_tile<A_MN>(ra, p.A, p.lda, m0, k_begin + (kb + 1) * TBK, M, k_end, tid);
      load_tile<B_MN>(rb, p.B, p.ldb, n0, k_begin + (kb + 1) * TBK, p.N, k_end, tid);
    } else if (mt_next < n_mtiles) {             // ... or the NEXT tile's first block: in flight across this tile's MMAs and epilogue
      load_tile<A_MN>(ra, p.A, p.lda, mt_next * TBM, k_begin, M, k_end, tid);
      load_tile<B_MN>(rb, p.B, p.ldb, n0, k_begin, p.N, k_end, tid);
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a0 = sbase + s * Plan::STAGE, b0 = a0 + PARTS * IMG_BYTES;
      bool first = (kb == 0);
#pragma unroll
      for (int ks = 0; ks < TBK / 16; ++ks) {
        // term pairs in ascending magnitude would change nothing measurable; (1,1) first keeps PARTS = 1 a prefix
        constexpr int PA[6] = {0, 0, 1, 1, 0, 2}, PB[6] = {0, 1, 0, 1, 2, 0};
#pragma unroll
        for (int t = 0; t < (PARTS == 1 ? 1 : 6); ++t) {
          const uint32_t aa = a0 + PA[t] * IMG_BYTES, bb = b0 + PB[t] * IMG_BYTES;
          const uint64_t da = A_MN ? umma_desc_mn(aa + ks * 2048, 8192, 1024) : umma_desc(aa + ks * 32);
          const uint64_t db = B_MN ? umma_desc_mn(bb + ks * 2048, 8192, 1024) : umma_desc(bb + ks * 32);
          umma_bf16_1cta(tmem_base, da, db, idesc, first ? 0u : 1u);
          first = false;
        }
      }
      umma_commit_1cta(bar(s));                  // frees the stage when these MMAs have read it
      if (kb == nkb - 1) umma_commit_1cta(bar(NSTG));
    }
    ++uses[s];
    __syncwarp();
  }
  if (ok && nkb > 0) ok = mbar_wait(bar(NSTG), tiles_done & 1, p.err, 912);
  ++tiles_done;
  if (ok) {
    tc_fence_after();
    // ---- epilogue: warp w reads TMEM lanes (w & 3) * 32 .. +32 (rows), columns (w >> 2) * 64 .. +64
    const int q = warp & 3, half = warp >> 2;
    const int row = m0 + q * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + half * 64;
    float* C = p.C;
    if (EPI == EPI_DW) C += (int64_t)blockIdx.z * p.M * p.ldc;
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {             // 16 columns at a time: the next tile's prefetched operands (64 registers) stay live
      if (half * 64 + cc * 16 >= n_used) break;  // warp-uniform: columns the MMA never produced
      uint32_t v[16];
      if (nkb > 0) {
        tmem_ld16(taddr + cc * 16, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0u;
      }
      const int col0 = n0 + half * 64 + cc * 16;
      if (row < M) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float x = __uint_as_float(v[i]);
          const int col = col0 + i;
          if (col < p.N) {
            if (EPI == EPI_FWD) {
              if (p.bias) x += __ldg(p.bias + col);
              x = act_fwd_fast(p.act, x);
            } else if (EPI == EPI_DX) {
              if (p.act != PM_ACT_NONE) x *= pm_act_bwd(p.act, __ldg(p.aux + (int64_t)row * p.ldaux + col));
            }
          }
          v[i] = __float_as_uint(x);
        }
        float* dst = C + (int64_t)row * p.ldc + col0;
        if (col0 + 16 <= p.N && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            reinterpret_cast<uint4*>(dst)[i] = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (col0 + i < p.N) dst[i] = __uint_as_float(v[i]);
        }
      }
    }
    tc_fence_before();                           // the next tile's first MMA overwrites the accumulator: ordered by the next __syncthreads
  }
  }  // tiles
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

__global__ void splitk_reduce_tc_kernel(const float* __restrict__ partial, int splits, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float t = 0.f;
  for (int s = 0; s < splits; ++s) t += partial[(int64_t)s * n + i];
  out[i] = t;
}

// db partials: part[split][n] = sum over the split's rows of dY[row, n]   (rows limited by *lim_dev).  A CTA takes 128 columns
// (32 lanes x float4) of its split's rows, eight rows at a time (one per warp) with two loads in flight per thread, then folds the
// eight row-lanes in shared memory in a fixed order.
__global__ void __launch_bounds__(256)
colsum_split_kernel(const float* __restrict__ dY, int64_t ld, int rows, int N, const int32_t* __restrict__ lim_dev, int rows_per_split,
                    float* __restrict__ part) {
  __shared__ float4 red[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n = blockIdx.x * 128 + lane * 4;
  if (lim_dev) {
    rows = min(rows, *lim_dev);
    rows_per_split = ((rows + (int)gridDim.y - 1) / (int)gridDim.y + TBK - 1) / TBK * TBK;
  }
  const int r0 = min(rows, (int)blockIdx.y * rows_per_split), r1 = min(rows, r0 + rows_per_split);
  const bool vec = (n + 4 <= N) && ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(dY) & 15u) == 0);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
  auto ld4 = [&](int r) -> float4 {
    const float* p = dY + (int64_t)r * ld + n;
    if (vec) return __ldg(reinterpret_cast<const float4*>(p));
    return make_float4(n < N ? __ldg(p) : 0.f, n + 1 < N ? __ldg(p + 1) : 0.f, n + 2 < N ? __ldg(p + 2) : 0.f, n + 3 < N ? __ldg(p + 3) : 0.f);
  };
  int r = r0 + w;
  for (; r + 8 < r1; r += 16) {
    const float4 u = ld4(r), v = ld4(r + 8);
    a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
    b.x += v.x; b.y += v.y; b.z += v.z; b.w += v.w;
  }
  if (r < r1) { const float4 u = ld4(r); a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w; }
  red[w][lane] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  __syncthreads();
  if (w == 0) {
    float4 t = red[0][lane];
#pragma unroll
    for (int k = 1; k < 8; ++k) { const float4 u = red[k][lane]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
    float* o = part + (int64_t)blockIdx.y * N + n;
    if (n < N) o[0] = t.x;
    if (n + 1 < N) o[1] = t.y;
    if (n + 2 < N) o[2] = t.z;
    if (n + 3 < N) o[3] = t.w;
  }
}

inline int dw_splits_tc(int Mrows, int N, int K) {
  const int tiles = pm_cdiv(N, TBM) * pm_cdiv(K, TBN);
  int s = pm_cdiv(PM_NUM_SMS, tiles);
  const int max_s = pm_cdiv(Mrows, 2 * TBK);       // at least two k-blocks per split
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  return s;
}

// FWD / DX: CTAs of an output column walk the M tiles round-robin; enough CTAs to fill the SMs at the kernel's occupancy
// (__launch_bounds__(.., 2)), never more than there are tiles
inline int gemm_tc_grid_x(int M, int n_tiles_n) {
  const int m_tiles = pm_cdiv(M, TBM);
  int gx = pm_cdiv(2 * pm_sm_count(), n_tiles_n);
  if (gx > m_tiles) gx = m_tiles;
  return gx < 1 ? 1 : gx;
}

template <int PARTS, int EPI, bool A_MN, bool B_MN>
int launch_gemm_tc(const GemmTcP& p, dim3 grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<PARTS, EPI, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)SmemPlan<PARTS>::TOTAL);
    if (e != cudaSuccess) PM_FAIL(PM_ERR_CUDA, "cudaFuncSetAttribute(gemm_tc smem): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  gemm_tc_kernel<PARTS, EPI, A_MN, B_MN><<<grid, GT_THREADS, SmemPlan<PARTS>::TOTAL, st>>>(p);
  return PM_OK;
}

}  // namespace

extern "C" {

int pm_linear_forward_tc(const float* x, int64_t ldx, const float* W, const float* b, float* y, int64_t ldy, int M, int N, int K,
                         int act, int precision, const int32_t* m_dev, pm_stream_t s) {
  PM_REQUIRE(x && W && y, PM_ERR_ARG, "pm_linear_forward_tc: null pointer");
  PM_REQUIRE(M > 0 && N > 0 && K > 0 && ldx >= K && ldy >= N, PM_ERR_SHAPE, "pm_linear_forward_tc: M=%d N=%d K=%d", M, N, K);
  PM_REQUIRE(act >= PM_ACT_NONE && act <= PM_ACT_SIGMOID, PM_ERR_ARG, "pm_linear_forward_tc: activation %d", act);
  PM_REQUIRE(precision == PM_PREC_FP32 || precision == PM_PREC_BF16, PM_ERR_ARG, "pm_linear_forward_tc: precision %d", precision);
  GemmTcP p{};
  p.A = x; p.lda = ldx; p.a_mn = 0;
  p.B = W; p.ldb = K; p.b_mn = 0;
  p.C = y; p.ldc = ldy; p.M = M; p.N = N; p.K = K; p.bias = b; p.act = act; p.lim_dev = m_dev;
  p.err = ErrSink{nullptr, pm_tc_sticky_word()};
  const dim3 grid(gemm_tc_grid_x(M, pm_cdiv(N, TBN)), pm_cdiv(N, TBN), 1);
  int rc = precision == PM_PREC_BF16 ? launch_gemm_tc<1, EPI_FWD, false, false>(p, grid, pm_st(s))
                                     : launch_gemm_tc<3, EPI_FWD, false, false>(p, grid, pm_st(s));
  if (rc) return rc;
  PM_CHECK_LAUNCH("pm_linear_forward_tc");
  return PM_OK;
}

constexpr int DB_SPLITS_MAX = 8 * PM_NUM_SMS;
size_t pm_linear_backward_tc_ws_bytes(int M, int N, int K) {
  return ((size_t)dw_splits_tc(M, N, K) * (size_t)N * K + (size_t)DB_SPLITS_MAX * N) * sizeof(float) + 256;
}

int pm_linear_backward_tc(const float* x, int64_t ldx, const float* W, const float* dpre, int64_t lddpre, float* dW, float* db,
                          float* dx, int64_t lddx, int M, int N, int K, int act_prev, int precision, const int32_t* m_dev,
                          void* ws, pm_stream_t s) {
  PM_REQUIRE(x && W && dpre && dW && ws, PM_ERR_ARG, "pm_linear_backward_tc: null pointer");
  PM_REQUIRE(M > 0 && N > 0 && K > 0, PM_ERR_SHAPE, "pm_linear_backward_tc: M=%d N=%d K=%d", M, N, K);
  PM_REQUIRE(precision == PM_PREC_FP32 || precision == PM_PREC_BF16, PM_ERR_ARG, "pm_linear_backward_tc: precision %d", precision);
  cudaStream_t st = pm_st(s);
  const ErrSink sink{nullptr, pm_tc_sticky_word()};
  // ---- dW[N,K] = dpre^T x  (contraction over the M rows, split across CTAs; fixed-order second stage), db = column sums
  const int splits = dw_splits_tc(M, N, K);
  const int kps = pm_cdiv(pm_cdiv(M, splits), TBK) * TBK;
  float* part = reinterpret_cast<float*>(ws);
  float* dbpart = part + (size_t)splits * N * K;
  const size_t ws_db_floats = (size_t)DB_SPLITS_MAX * N;
  {
    GemmTcP p{};
    p.A = dpre; p.lda = lddpre; p.a_mn = 1;         // A[K=rows x M=N'] : dpre as stored
    p.B = x; p.ldb = ldx; p.b_mn = 1;               // B[K=rows x N=K'] : x as stored
    p.C = part; p.ldc = K; p.M = N; p.N = K; p.K = M; p.lim_dev = m_dev; p.k_per_split = kps; p.err = sink;
    const dim3 grid(pm_cdiv(N, TBM), pm_cdiv(K, TBN), splits);
    int rc = precision == PM_PREC_BF16 ? launch_gemm_tc<1, EPI_DW, true, true>(p, grid, st) : launch_gemm_tc<3, EPI_DW, true, true>(p, grid, st);
    if (rc) return rc;
    splitk_reduce_tc_kernel<<<pm_cdiv((int64_t)N * K, 256), 256, 0, st>>>(part, splits, (int64_t)N * K, dW);
    if (db) {
      // the column sums get their own (finer) split: the rows are streamed once and the kernel lives on loads in flight
      int csplits = pm_cdiv(8 * PM_NUM_SMS, pm_cdiv(N, 128));
      const int cmax = (int)((ws_db_floats) / (size_t)N);
      if (csplits > cmax) csplits = cmax;
      if (csplits > pm_cdiv(M, 64)) csplits = pm_cdiv(M, 64);
      if (csplits < 1) csplits = 1;
      const int crows = pm_cdiv(pm_cdiv(M, csplits), TBK) * TBK;
      colsum_split_kernel<<<dim3(pm_cdiv(N, 128), csplits), 256, 0, st>>>(dpre, lddpre, M, N, m_dev, crows, dbpart);
      splitk_reduce_tc_kernel<<<pm_cdiv(N, 256), 256, 0, st>>>(dbpart, csplits, N, db);
    }
  }
  // ---- dx[M,K] = (dpre W) * act'(x)
  if (dx) {
    GemmTcP q{};
    q.A = dpre; q.lda = lddpre; q.a_mn = 0;
    q.B = W; q.ldb = K; q.b_mn = 1;                 // B[K=N' x N=K'] : W as stored
    q.C = dx; q.ldc = lddx; q.M = M; q.N = K; q.K = N; q.act = act_prev; q.aux = x; q.ldaux = ldx; q.lim_dev = m_dev; q.err = sink;
    const dim3 grid(gemm_tc_grid_x(M, pm_cdiv(K, TBN)), pm_cdiv(K, TBN), 1);
    int rc = precision == PM_PREC_BF16 ? launch_gemm_tc<1, EPI_DX, false, true>(q, grid, st) : launch_gemm_tc<3, EPI_DX, false, true>(q, grid, st);
    if (rc) return rc;
  }
  PM_CHECK_LAUNCH("pm_linear_backward_tc");
  return PM_OK;
}

}  // extern "C"
