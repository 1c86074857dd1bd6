// K3 — fp32 dense layers on CUDA cores: y = act(x W^T + b) and its backward.
// reference: nn.Linear inside MLP (algorithms/algo_utils/network.py:27-54) and the PointNet head
// (network.py:152-159).  Also the work-horse of the critical-point encoder backward (pointnet.cu),
// where the row count lives in device memory (m_dev) so nothing syncs with the host.
// 128x128x16 CTA tiles, 8x8 register micro-tiles, 256 threads.  Three operand layouts:
//   FWD  C[m,n] = act(sum_k X[m,k] W[n,k] + b[n])
//   DX   C[m,k] = (sum_n dPre[m,n] W[n,k]) * act'(Xprev[m,k])
//   DW   C[n,k] = sum_m dPre[m,n] X[m,k]        (split over m, fixed-order second-stage reduce)
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;
constexpr int PAD = 4;

enum { EPI_FWD = 0, EPI_DX = 1, EPI_DW = 2 };

struct GemmP {
  const float* A; int64_t a_rs, a_cs;      // A(m,k) = A[m*a_rs + k*a_cs]
  const float* B; int64_t b_rs, b_cs;      // B(k,n) = B[k*b_rs + n*b_cs]
  float* C; int64_t ldc;
  int M, N, K;
  const float* bias;                       // FWD
  const float* aux; int64_t ldaux;         // DX: previous activation output
  int act;
  const int32_t* lim_dev;                  // FWD/DX: valid rows (<= M);  DW: valid reduction length (<= K)
  int k_per_split;                         // DW
  float* db_partial;                       // DW: [splits][M] column sums of dPre (only n-tile 0 writes)
};

template <bool A_KCONTIG, bool B_NCONTIG, int EPI>
__global__ void __launch_bounds__(NT)
gemm_kernel(const GemmP p) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  int M = p.M, K = p.K;
  int k_begin = 0, k_end = K;
  if (EPI == EPI_DW) {
    int kps = p.k_per_split;
    if (p.lim_dev) {   // spread the VALID rows over all splits (the static bound can be 5-10x the live row count)
      K = min(K, *p.lim_dev);
      kps = ((K + (int)gridDim.z - 1) / (int)gridDim.z + BK - 1) / BK * BK;
    }
    k_begin = blockIdx.z * kps;
    k_end = min(K, k_begin + kps);
  } else {
    if (p.lim_dev) M = min(M, *p.lim_dev);
    if (m0 >= M) return;
  }
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float dbacc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) dbacc[i] = 0.f;
  const bool do_db = (EPI == EPI_DW) && p.db_partial && blockIdx.x == 0 && tx == 0;

  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    // ---- stage tiles (generic strides; the contiguous axis is mapped to consecutive threads)
#pragma unroll
    for (int j = 0; j < (BM * BK) / NT; ++j) {
      const int idx = tid + j * NT;
      int m, k;
      if (A_KCONTIG) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < k_end) v = __ldg(p.A + (int64_t)gm * p.a_rs + (int64_t)gk * p.a_cs);
      As[k][m] = v;
    }
#pragma unroll
    for (int j = 0; j < (BN * BK) / NT; ++j) {
      const int idx = tid + j * NT;
      int n, k;
      if (B_NCONTIG) { n = idx % BN; k = idx / BN; } else { k = idx % BK; n = idx / BK; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < p.N && gk < k_end) v = __ldg(p.B + (int64_t)gk * p.b_rs + (int64_t)gn * p.b_cs);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], b[8];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8 + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      if (do_db) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dbacc[i] += a[i];
      }
    }
    __syncthreads();
  }

  // ---- epilogue
  float* C = p.C;
  if (EPI == EPI_DW) C += (int64_t)blockIdx.z * p.M * p.ldc;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + ty * 8 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gn = n0 + tx * 8 + j;
      if (gn >= p.N) continue;
      float v = acc[i][j];
      if (EPI == EPI_FWD) {
        if (p.bias) v += p.bias[gn];
        v = pm_act_fwd(p.act, v);
      } else if (EPI == EPI_DX) {
        if (p.act != PM_ACT_NONE) v *= pm_act_bwd(p.act, p.aux[(int64_t)gm * p.ldaux + gn]);
      }
      C[(int64_t)gm * p.ldc + gn] = v;
    }
  }
  if (do_db) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int gm = m0 + ty * 8 + i;
      if (gm < M) p.db_partial[(int64_t)blockIdx.z * p.M + gm] = dbacc[i];
    }
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float t = 0.f;
  for (int s = 0; s < splits; ++s) t += partial[(int64_t)s * n + i];
  out[i] = t;
}

inline int dw_splits(int Mmax, int N, int K) {
  const int tiles = pm_cdiv(N, BM) * pm_cdiv(K, BN);
  int s = pm_cdiv(2 * PM_NUM_SMS, tiles);
  const int max_s = pm_cdiv(Mmax, 128);
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  return s;
}

}  // namespace

extern "C" {

int pm_linear_forward(const float* x, int64_t ldx, const float* W, const float* b, float* y, int64_t ldy, int M,
                      int N, int K, int act, const int32_t* m_dev, pm_stream_t s) {
  PM_REQUIRE(x && W && y, PM_ERR_ARG, "pm_linear_forward: null pointer");
  PM_REQUIRE(M > 0 && N > 0 && K > 0 && ldx >= K && ldy >= N, PM_ERR_SHAPE, "pm_linear_forward: M=%d N=%d K=%d", M, N, K);
  PM_REQUIRE(act >= PM_ACT_NONE && act <= PM_ACT_SIGMOID, PM_ERR_ARG, "pm_linear_forward: activation %d", act);
  GemmP p{};
  p.A = x; p.a_rs = ldx; p.a_cs = 1;
  p.B = W; p.b_rs = 1; p.b_cs = K;
  p.C = y; p.ldc = ldy; p.M = M; p.N = N; p.K = K; p.bias = b; p.act = act; p.lim_dev = m_dev;
  dim3 grd(pm_cdiv(N, BN), pm_cdiv(M, BM), 1);
  gemm_kernel<true, false, EPI_FWD><<<grd, NT, 0, pm_st(s)>>>(p);
  PM_CHECK_LAUNCH("pm_linear_forward");
  return PM_OK;
}

size_t pm_linear_backward_ws_bytes(int M, int N, int K) {
  return (size_t)dw_splits(M, N, K) * ((size_t)N * K + N) * sizeof(float);
}

int pm_linear_backward(const float* x, int64_t ldx, const float* W, const float* dpre, int64_t lddpre, float* dW,
                       float* db, float* dx, int64_t lddx, int M, int N, int K, int act_prev, const int32_t* m_dev,
                       void* ws, pm_stream_t s) {
  PM_REQUIRE(x && W && dpre && dW && ws, PM_ERR_ARG, "pm_linear_backward: null pointer");
  PM_REQUIRE(M > 0 && N > 0 && K > 0, PM_ERR_SHAPE, "pm_linear_backward: M=%d N=%d K=%d", M, N, K);
  cudaStream_t st = pm_st(s);
  // ---- dW[N,K] = dpre^T x (+ db = colsum dpre), reduction over the M rows split across CTAs
  const int splits = dw_splits(M, N, K);
  int kps = pm_cdiv(M, splits);
  kps = pm_cdiv(kps, BK) * BK;
  float* part = reinterpret_cast<float*>(ws);
  float* dbpart = part + (size_t)splits * N * K;
  GemmP p{};
  p.A = dpre; p.a_rs = 1; p.a_cs = lddpre;          // A(m=n', k=row) = dpre[row*ld + n']
  p.B = x; p.b_rs = ldx; p.b_cs = 1;                // B(k=row, n=k') = x[row*ldx + k']
  p.C = part; p.ldc = K; p.M = N; p.N = K; p.K = M; p.lim_dev = m_dev; p.k_per_split = kps;
  p.db_partial = db ? dbpart : nullptr;
  dim3 grd(pm_cdiv(K, BN), pm_cdiv(N, BM), splits);
  gemm_kernel<false, true, EPI_DW><<<grd, NT, 0, st>>>(p);
  splitk_reduce_kernel<<<pm_cdiv((int64_t)N * K, 256), 256, 0, st>>>(part, splits, (int64_t)N * K, dW);
  if (db) splitk_reduce_kernel<<<pm_cdiv(N, 256), 256, 0, st>>>(dbpart, splits, N, db);
  // ---- dx[M,K] = (dpre W) * act'(x)
  if (dx) {
    GemmP q{};
    q.A = dpre; q.a_rs = lddpre; q.a_cs = 1;
    q.B = W; q.b_rs = K; q.b_cs = 1;                // B(k=n', n=k') = W[n'*K + k']
    q.C = dx; q.ldc = lddx; q.M = M; q.N = K; q.K = N; q.act = act_prev; q.aux = x; q.ldaux = ldx; q.lim_dev = m_dev;
    dim3 g2(pm_cdiv(K, BN), pm_cdiv(M, BM), 1);
    gemm_kernel<true, true, EPI_DX><<<g2, NT, 0, st>>>(q);
  }
  PM_CHECK_LAUNCH("pm_linear_backward");
  return PM_OK;
}

}  // extern "C"
