// K3b — fused PointNet head  Linear(F,128)-act-Linear(128,32)-act-Linear(32,out), fp32, forward and backward.
// reference: algorithms/algo_utils/network.py:152-159 (final_mlp) and 186-198 (cat(max[,mean][,proprio]) -> final_mlp).
//
// The three layers are skinny (N = 128 / 32 / <=32): as separate 128x128-tile GEMMs they fill 16 of 148 SMs and cost
// 12 launches per minibatch (ncu r01b: 20 % of the iteration).  Here a CTA owns 16 batch rows and carries them through
// all three layers in shared memory (forward: 1 launch), and the backward is 3 launches:
//   bwd_a  per 16 rows: dPre2, dPre1 (-> global, for bwd_b), dfeat = dPre1 . W0, and per-CTA partials of the small grads
//   bwd_b  dW0[n,k] = sum_rows dPre1[r,n] feat[r,k] over (64-column, 128-row) blocks -> per-slab partials
//   reduce fixed-order sums of the partials (deterministic; no atomics)
#include "common.cuh"

namespace {

constexpr int HR = 16;            // batch rows per CTA
constexpr int HT = 256;           // threads (backward kernels)
constexpr int HF = 512;           // threads of the forward kernel: 4 warps per scheduler hide the LDS / L2 latencies (2 ran at IPC 0.26)
constexpr int H1D = 128, H2D = 32, OMAX = 32;
constexpr int KC = 32;            // k-chunk of layer 0
constexpr int RB = 64;            // rows per slab in bwd_b (256 CTAs at B = 2048: two per SM)
constexpr int CBK = 64;           // columns per CTA in bwd_b

// per-CTA partial record of bwd_a (floats)
constexpr int QW1 = 0;                          // [32][128]
constexpr int QB1 = 4096;                       // [32]
constexpr int QW2 = 4128;                       // [OMAX][32]
constexpr int QB2 = 5152;                       // [OMAX]
constexpr int QB0 = 5184;                       // [128]
constexpr int QPART = 5312;

struct FwdS {
  float Ws[H1D][KC + 1];
  float Fs[KC][HR];
  float H1s[HR][H1D];
  float W1s[H2D][H1D + 1];
  float H2s[HR][H2D + 1];
  float W2s[OMAX][H2D + 1];
};

template <int ACT>
__global__ void __launch_bounds__(HF)
head_fwd_kernel(const float* __restrict__ feat, int64_t ldf, int B, int F, pm_head_params P, int out_dim,
                float* __restrict__ h1, float* __restrict__ h2, float* __restrict__ out, int64_t ldo) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdS& S = *reinterpret_cast<FwdS*>(smem_raw);
  const int t = threadIdx.x, row0 = blockIdx.x * HR;
  for (int i = t; i < H2D * H1D; i += HF) S.W1s[i >> 7][i & 127] = P.W1[i];
  for (int i = t; i < out_dim * H2D; i += HF) S.W2s[i >> 5][i & 31] = P.W2[i];
  // ---- layer 0: thread = (column n, 4-row group rg)
  const int n = t & 127, rg = t >> 7;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float wreg[8], freg;
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = t + i * HF, nn = idx >> 5, kk = idx & 31;
      wreg[i] = (k0 + kk < F) ? __ldg(P.W0 + (int64_t)nn * F + k0 + kk) : 0.f;
    }
    const int r = t >> 5, kk = t & 31;
    freg = (row0 + r < B && k0 + kk < F) ? __ldg(feat + (int64_t)(row0 + r) * ldf + k0 + kk) : 0.f;
  };
  fetch(0);
  for (int k0 = 0; k0 < F; k0 += KC) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { const int idx = t + i * HF; S.Ws[idx >> 5][idx & 31] = wreg[i]; }
    S.Fs[t & 31][t >> 5] = freg;
    __syncthreads();
    if (k0 + KC < F) fetch(k0 + KC);
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      const float w = S.Ws[n][kk];
      const float4 f0 = *reinterpret_cast<const float4*>(&S.Fs[kk][rg * 4]);
      acc[0] = fmaf(w, f0.x, acc[0]); acc[1] = fmaf(w, f0.y, acc[1]); acc[2] = fmaf(w, f0.z, acc[2]); acc[3] = fmaf(w, f0.w, acc[3]);
    }
    __syncthreads();
  }
  {
    const float bb = P.b0[n];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = rg * 4 + i;
      const float v = pm_act_fwd(ACT, acc[i] + bb);
      S.H1s[r][n] = v;
      if (row0 + r < B) h1[(int64_t)(row0 + r) * H1D + n] = v;
    }
  }
  __syncthreads();
  // ---- layer 1: thread = (column m, row rr)
  {
    const int m = t & 31, rr = t >> 5;
    float a0 = P.b1[m];
#pragma unroll 8
    for (int k = 0; k < H1D; ++k) a0 = fmaf(S.W1s[m][k], S.H1s[rr][k], a0);
    a0 = pm_act_fwd(ACT, a0);
    S.H2s[rr][m] = a0;
    if (row0 + rr < B) h2[(int64_t)(row0 + rr) * H2D + m] = a0;
  }
  __syncthreads();
  // ---- layer 2 (no activation)
  for (int i = t; i < HR * out_dim; i += HF) {
    const int r = i / out_dim, o = i - r * out_dim;
    float a = P.b2[o];
#pragma unroll
    for (int k = 0; k < H2D; ++k) a = fmaf(S.W2s[o][k], S.H2s[r][k], a);
    if (row0 + r < B) out[(int64_t)(row0 + r) * ldo + o] = a;
  }
}

struct BwdS {
  float Ds[HR][OMAX + 1];
  float H2s[HR][H2D + 1];
  float H1s[HR][H1D];
  float DP2s[HR][H2D + 1];
  float DP1t[H1D][HR];          // dPre1 transposed: [n][row]
  float W1s[H2D][H1D];
  float W2s[OMAX][H2D + 1];
};

template <int ACT>
__global__ void __launch_bounds__(HF)
head_bwd_a_kernel(int B, int F, pm_head_params P, int out_dim, const float* __restrict__ h1, const float* __restrict__ h2,
                  const float* __restrict__ dout, int64_t lddo, float* __restrict__ dpre1, float* __restrict__ dfeat,
                  int64_t lddf, int dfeat_cols, float* __restrict__ part_all) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdS& S = *reinterpret_cast<BwdS*>(smem_raw);
  const int t = threadIdx.x, row0 = blockIdx.x * HR;
  float* part = part_all + (size_t)blockIdx.x * QPART;
  for (int i = t; i < H2D * H1D; i += HF) S.W1s[i >> 7][i & 127] = P.W1[i];
  for (int i = t; i < out_dim * H2D; i += HF) S.W2s[i >> 5][i & 31] = P.W2[i];
  for (int i = t; i < HR * out_dim; i += HF) {
    const int r = i / out_dim, o = i - r * out_dim;
    S.Ds[r][o] = (row0 + r < B) ? dout[(int64_t)(row0 + r) * lddo + o] : 0.f;
  }
  {
    const int r = t >> 5, m = t & 31;                               // HR * H2D == HF
    S.H2s[r][m] = (row0 + r < B) ? h2[(int64_t)(row0 + r) * H2D + m] : 0.f;
  }
  for (int i = t; i < HR * H1D; i += HF) {
    const int r = i >> 7, nn = i & 127;
    S.H1s[r][nn] = (row0 + r < B) ? h1[(int64_t)(row0 + r) * H1D + nn] : 0.f;
  }
  __syncthreads();
  // ---- dPre2[r][m] = (dout . W2)[r][m] * act'(h2): thread = (column m, row r)
  {
    const int m = t & 31, r = t >> 5;
    float a = 0.f;
    for (int o = 0; o < out_dim; ++o) a = fmaf(S.Ds[r][o], S.W2s[o][m], a);
    S.DP2s[r][m] = a * pm_act_bwd(ACT, S.H2s[r][m]);
  }
  __syncthreads();
  // ---- dPre1[r][n] = (dPre2 . W1)[r][n] * act'(h1): thread = (column n, 4-row group rg)
  {
    const int n = t & 127, rg = t >> 7;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int m = 0; m < H2D; ++m) {
      const float w = S.W1s[m][n];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(S.DP2s[rg * 4 + i][m], w, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = rg * 4 + i;
      const float v = acc[i] * pm_act_bwd(ACT, S.H1s[r][n]);
      S.DP1t[n][r] = v;
      if (row0 + r < B) dpre1[(int64_t)(row0 + r) * H1D + n] = v;
    }
  }
  __syncthreads();
  // ---- small gradients of this 16-row block
  {
    const int n = t & 127, mg = t >> 7;          // dW1[m][n], m in [mg*8, +8)
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 4
    for (int r = 0; r < HR; ++r) {
      const float hv = S.H1s[r][n];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(S.DP2s[r][mg * 8 + i], hv, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) part[QW1 + (mg * 8 + i) * H1D + n] = acc[i];
    if (t < H1D) {
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < HR; ++r) s += S.DP1t[t][r];
      part[QB0 + t] = s;
    } else if (t < H1D + H2D) {
      const int m = t - H1D;
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < HR; ++r) s += S.DP2s[r][m];
      part[QB1 + m] = s;
    } else if (t < H1D + H2D + OMAX) {
      const int o = t - H1D - H2D;
      float s = 0.f;
      if (o < out_dim)
        for (int r = 0; r < HR; ++r) s += S.Ds[r][o];
      part[QB2 + o] = s;
    }
    for (int i = t; i < out_dim * H2D; i += HF) {
      const int o = i >> 5, m = i & 31;
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < HR; ++r) s = fmaf(S.Ds[r][o], S.H2s[r][m], s);
      part[QW2 + o * H2D + m] = s;
    }
  }
  // ---- dfeat[r][k] = sum_n dPre1[r][n] W0[n][k]   (only the first dfeat_cols columns are consumed: max[,mean] part)
  if (dfeat) {
    for (int k = t; k < dfeat_cols; k += HF) {
      float acc[HR];
#pragma unroll
      for (int i = 0; i < HR; ++i) acc[i] = 0.f;
#pragma unroll 8
      for (int n = 0; n < H1D; ++n) {
        const float w = __ldg(P.W0 + (int64_t)n * F + k);
        const float4 d0 = *reinterpret_cast<const float4*>(&S.DP1t[n][0]);
        const float4 d1 = *reinterpret_cast<const float4*>(&S.DP1t[n][4]);
        const float4 d2 = *reinterpret_cast<const float4*>(&S.DP1t[n][8]);
        const float4 d3 = *reinterpret_cast<const float4*>(&S.DP1t[n][12]);
        acc[0] = fmaf(d0.x, w, acc[0]); acc[1] = fmaf(d0.y, w, acc[1]); acc[2] = fmaf(d0.z, w, acc[2]); acc[3] = fmaf(d0.w, w, acc[3]);
        acc[4] = fmaf(d1.x, w, acc[4]); acc[5] = fmaf(d1.y, w, acc[5]); acc[6] = fmaf(d1.z, w, acc[6]); acc[7] = fmaf(d1.w, w, acc[7]);
        acc[8] = fmaf(d2.x, w, acc[8]); acc[9] = fmaf(d2.y, w, acc[9]); acc[10] = fmaf(d2.z, w, acc[10]); acc[11] = fmaf(d2.w, w, acc[11]);
        acc[12] = fmaf(d3.x, w, acc[12]); acc[13] = fmaf(d3.y, w, acc[13]); acc[14] = fmaf(d3.z, w, acc[14]); acc[15] = fmaf(d3.w, w, acc[15]);
      }
#pragma unroll
      for (int i = 0; i < HR; ++i)
        if (row0 + i < B) dfeat[(int64_t)(row0 + i) * lddf + k] = acc[i];
    }
  }
}

// dW0 partial over a (64-column, 128-row) block: part[slab][n][k] = sum_{r in slab} dPre1[r][n] feat[r][k]
__global__ void __launch_bounds__(HT)
head_bwd_b_kernel(const float* __restrict__ feat, int64_t ldf, int B, int F, const float* __restrict__ dpre1,
                  float* __restrict__ part) {
  __shared__ __align__(16) float Fs[HR][CBK];
  __shared__ __align__(16) float Dsm[HR][H1D];
  const int t = threadIdx.x, k0 = blockIdx.x * CBK, slab = blockIdx.y;
  const int n = t & 127, kg = t >> 7;          // thread: row n of dW0, 32 columns [k0 + kg*32, +32)
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  const int r_begin = slab * RB, r_end = min(B, r_begin + RB);
  for (int rb = r_begin; rb < r_end; rb += HR) {
    __syncthreads();
    for (int i = t; i < HR * CBK; i += HT) {
      const int r = i >> 6, kk = i & 63;
      Fs[r][kk] = (rb + r < r_end && k0 + kk < F) ? __ldg(feat + (int64_t)(rb + r) * ldf + k0 + kk) : 0.f;
    }
    for (int i = t; i < HR * H1D; i += HT) {
      const int r = i >> 7, nn = i & 127;
      Dsm[r][nn] = (rb + r < r_end) ? __ldg(dpre1 + (int64_t)(rb + r) * H1D + nn) : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int r = 0; r < HR; ++r) {
      const float d = Dsm[r][n];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 f = *reinterpret_cast<const float4*>(&Fs[r][kg * 32 + q * 4]);
        acc[4 * q] = fmaf(d, f.x, acc[4 * q]); acc[4 * q + 1] = fmaf(d, f.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(d, f.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(d, f.w, acc[4 * q + 3]);
      }
    }
  }
  float* dst = part + ((size_t)slab * H1D + n) * F + k0 + kg * 32;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (k0 + kg * 32 + i < F) dst[i] = acc[i];
}

__global__ void __launch_bounds__(256)
head_reduce_kernel(const float* __restrict__ partA, int nA, const float* __restrict__ partB, int slabs, int F, int out_dim,
                   pm_head_grads g) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int nW0 = H1D * F;
  if (idx < nW0) {
    float s = 0.f;
    for (int k = 0; k < slabs; ++k) s += partB[(size_t)k * nW0 + idx];
    g.W0[idx] = s;
    return;
  }
  int i = idx - nW0;
  float* dst = nullptr;
  int off = 0;
  if (i < H2D * H1D) { dst = g.W1 + i; off = QW1 + i; }
  else if ((i -= H2D * H1D) < H2D) { dst = g.b1 + i; off = QB1 + i; }
  else if ((i -= H2D) < out_dim * H2D) { dst = g.W2 + i; off = QW2 + i; }
  else if ((i -= out_dim * H2D) < out_dim) { dst = g.b2 + i; off = QB2 + i; }
  else if ((i -= out_dim) < H1D) { dst = g.b0 + i; off = QB0 + i; }
  else return;
  float s = 0.f;
  for (int k = 0; k < nA; ++k) s += partA[(size_t)k * QPART + off];
  *dst = s;
}

struct HeadWs {
  float *dpre1, *partA, *partB;
  int nA, slabs;
  size_t total;
};
inline HeadWs carve_head(void* ws, int B, int F) {
  HeadWs w{};
  w.nA = pm_cdiv(B, HR);
  w.slabs = pm_cdiv(B, RB);
  size_t off = 0;
  char* base = reinterpret_cast<char*>(ws);
  auto take = [&](size_t bytes) { void* p = base ? base + off : nullptr; off += pm_align_up(bytes, 256); return p; };
  w.dpre1 = (float*)take((size_t)B * H1D * 4);
  w.partA = (float*)take((size_t)w.nA * QPART * 4);
  w.partB = (float*)take((size_t)w.slabs * H1D * F * 4);
  w.total = off;
  return w;
}

}  // namespace

extern "C" {

int pm_pointnet_head_forward(const float* feat, int64_t ldf, int B, int F, const pm_head_params* p, int out_dim, int act,
                             int precision, float* h1, float* h2, float* out, int64_t ldo, pm_stream_t s) {
  PM_REQUIRE(feat && p && h1 && h2 && out, PM_ERR_ARG, "pm_pointnet_head_forward: null pointer");
  PM_REQUIRE(B > 0 && F > 0 && ldf >= F && out_dim >= 1 && out_dim <= OMAX && ldo >= out_dim, PM_ERR_SHAPE,
             "pm_pointnet_head_forward: B=%d F=%d out=%d (out <= %d)", B, F, out_dim, OMAX);
  // the head (0.04 % of the network's FLOPs) is fp32 on CUDA cores in every mode: a tcgen05 variant (round 1) was no faster at
  // B = 2048 — 16-32 CTAs, staging-latency bound — and was removed
  PM_REQUIRE(precision == PM_PREC_FP32 || precision == PM_PREC_FP32_FFMA, PM_ERR_UNSUPPORTED, "pm_pointnet_head_forward: precision %d (the head runs in fp32)", precision);
#define PM_HF(ACTV)                                                                                                      \
  case ACTV: {                                                                                                           \
    static bool attr_set = false;                                                                                        \
    if (!attr_set) {                                                                                                     \
      cudaError_t e = cudaFuncSetAttribute(head_fwd_kernel<ACTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FwdS)); \
      if (e != cudaSuccess) PM_FAIL(PM_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));                     \
      attr_set = true;                                                                                                   \
    }                                                                                                                    \
    head_fwd_kernel<ACTV><<<pm_cdiv(B, HR), HF, sizeof(FwdS), pm_st(s)>>>(feat, ldf, B, F, *p, out_dim, h1, h2, out, ldo); \
  } break;
  switch (act) {
    PM_HF(PM_ACT_NONE) PM_HF(PM_ACT_TANH) PM_HF(PM_ACT_RELU) PM_HF(PM_ACT_ELU) PM_HF(PM_ACT_SELU) PM_HF(PM_ACT_LRELU)
    PM_HF(PM_ACT_SIGMOID)
    default: PM_FAIL(PM_ERR_ARG, "pm_pointnet_head_forward: activation %d", act);
  }
#undef PM_HF
  PM_CHECK_LAUNCH("pm_pointnet_head_forward");
  return PM_OK;
}

size_t pm_pointnet_head_backward_ws_bytes(int B, int F) { return carve_head(nullptr, B, F).total; }

int pm_pointnet_head_backward(const float* feat, int64_t ldf, int B, int F, const pm_head_params* p, int out_dim, int act,
                              int precision, const float* h1, const float* h2, const float* dout, int64_t lddo,
                              const pm_head_grads* g, float* dfeat, int64_t lddf, int dfeat_cols, void* ws, size_t ws_bytes,
                              pm_stream_t s) {
  PM_REQUIRE(feat && p && h1 && h2 && dout && g && ws, PM_ERR_ARG, "pm_pointnet_head_backward: null pointer");
  PM_REQUIRE(B > 0 && F > 0 && ldf >= F && out_dim >= 1 && out_dim <= OMAX && lddo >= out_dim, PM_ERR_SHAPE,
             "pm_pointnet_head_backward: B=%d F=%d out=%d (out <= %d)", B, F, out_dim, OMAX);
  PM_REQUIRE(!dfeat || (dfeat_cols > 0 && dfeat_cols <= F && lddf >= dfeat_cols), PM_ERR_SHAPE, "pm_pointnet_head_backward: dfeat_cols=%d", dfeat_cols);
  PM_REQUIRE(precision == PM_PREC_FP32 || precision == PM_PREC_FP32_FFMA, PM_ERR_UNSUPPORTED, "pm_pointnet_head_backward: precision %d (the head runs in fp32)", precision);
  HeadWs w = carve_head(ws, B, F);
  PM_REQUIRE(ws_bytes >= w.total, PM_ERR_ARG, "pm_pointnet_head_backward: workspace %zu < %zu", ws_bytes, w.total);
  cudaStream_t st = pm_st(s);
  float* dfeat_ffma = dfeat;
#define PM_HB(ACTV)                                                                                                      \
  case ACTV: {                                                                                                           \
    static bool attr_set = false;                                                                                        \
    if (!attr_set) {                                                                                                     \
      cudaError_t e = cudaFuncSetAttribute(head_bwd_a_kernel<ACTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdS)); \
      if (e != cudaSuccess) PM_FAIL(PM_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));                     \
      attr_set = true;                                                                                                   \
    }                                                                                                                    \
    head_bwd_a_kernel<ACTV><<<w.nA, HF, sizeof(BwdS), st>>>(B, F, *p, out_dim, h1, h2, dout, lddo, w.dpre1, dfeat_ffma, lddf, \
                                                            dfeat_cols, w.partA);                                        \
  } break;
  switch (act) {
    PM_HB(PM_ACT_NONE) PM_HB(PM_ACT_TANH) PM_HB(PM_ACT_RELU) PM_HB(PM_ACT_ELU) PM_HB(PM_ACT_SELU) PM_HB(PM_ACT_LRELU)
    PM_HB(PM_ACT_SIGMOID)
    default: PM_FAIL(PM_ERR_ARG, "pm_pointnet_head_backward: activation %d", act);
  }
#undef PM_HB
  const int slabs = w.slabs;
  head_bwd_b_kernel<<<dim3(pm_cdiv(F, CBK), w.slabs), HT, 0, st>>>(feat, ldf, B, F, w.dpre1, w.partB);
  const int n_out = H1D * F + H2D * H1D + H2D + out_dim * H2D + out_dim + H1D;
  head_reduce_kernel<<<pm_cdiv(n_out, 256), 256, 0, st>>>(w.partA, w.nA, w.partB, slabs, F, out_dim, *g);
  PM_CHECK_LAUNCH("pm_pointnet_head_backward");
  return PM_OK;
}

}  // extern "C"
