// K1/K2 — PointNet encoder (per-point MLP C->128->256->512 + symmetric max/mean pool), fp32 path.
// reference: algorithms/algo_utils/network.py:141-150 (layers), 165-182 (forward up to the pooled feature).
//
// Forward (encoder_fwd_fp32): one CTA owns a whole cloud and walks it in 64-point tiles.  The
// (points x {128,256,512}) activations the reference materialises in HBM (5 KB/point) live only in
// shared memory / registers; HBM traffic is the 4C-byte point read plus 2-4 KB of pooled output per cloud.
// Backward (pm_pointnet_encode_backward): max-pool routes dfeat[b,c] to ONE point, so only the unique
// "critical" points of each cloud carry gradient.  Their rows are compacted on the device
// (count kept in device memory — no host sync), activations recomputed from the inputs, and layers 3..1
// back-propagated over those rows with the dense kernels of dense.cu.
#include "common.cuh"

extern "C" int pm_pointnet_encode_forward_tc(const float* x, int64_t ldx, int B, int N, int C,
                                             const pm_encoder_params* p, int act, float* feat, int64_t ldf,
                                             int32_t* argmax, void* ws, size_t ws_bytes, pm_stream_t s);
extern "C" size_t pm_pointnet_encode_forward_tc_ws_bytes(int B, int N, int C);
extern "C" int pm_pointnet_encode_backward_tc(const float* x, int64_t ldx, int B, int N, int C, const pm_encoder_params* p,
                                              int act, const float* dfeat, int64_t lddf, const int32_t* argmax,
                                              const pm_encoder_grads* g, void* ws, size_t ws_bytes, pm_stream_t s);
extern "C" size_t pm_pointnet_encode_backward_tc_ws_bytes(int B, int N, int C);

extern "C" int pm_linear_forward_tc(const float* x, int64_t ldx, const float* W, const float* b, float* y, int64_t ldy, int M, int N,
                                    int K, int act, int precision, const int32_t* m_dev, pm_stream_t s);
extern "C" size_t pm_linear_backward_tc_ws_bytes(int M, int N, int K);
extern "C" int pm_linear_backward_tc(const float* x, int64_t ldx, const float* W, const float* dpre, int64_t lddpre, float* dW,
                                     float* db, float* dx, int64_t lddx, int M, int N, int K, int act_prev, int precision,
                                     const int32_t* m_dev, void* ws, pm_stream_t s);

extern "C" int pm_pointnet_encode_forward_tc3(const float* x, int64_t ldx, int B, int N, int C, const pm_encoder_params* p, int act,
                                              float* feat, int64_t ldf, int32_t* argmax, void* ws, size_t ws_bytes, pm_stream_t s);
extern "C" size_t pm_pointnet_encode_forward_tc3_ws_bytes(int B, int N, int C);
extern "C" int pm_pointnet_encode_forward_tc3_supported(int N, int C);

namespace {

constexpr int TP = 64;          // points per tile
constexpr int FT = 256;         // threads
constexpr int KT = 8;           // k-slice of the streamed weight matrix
constexpr int HS = TP;          // row stride of the k-major activation tiles (reads are warp-broadcast; no padding needed)
constexpr int CMAX = 8;

struct FwdSmem {
  float H1t[128][HS];           // layer-1 output, k-major: H1t[ch][pt]       (32 KB) — reused as the max-reduce scratch
  float H2t[256][HS];           // layer-2 output, k-major                    (64 KB)
  float Ws[KT][256];            // streamed weight slice, Ws[k][ch]           ( 8 KB)
  float xs[TP][CMAX];
  float W1s[128][CMAX];
  float b1s[128];
  float b2s[256];
};

// acc[i][j] += sum_k Ht[k][pt0+i] * W[ch(j)][k]   with ch(j) = lane + 32 j  (+ ch_base), over K in slices of KT
template <int K>
__device__ __forceinline__ void tile_gemm(float (&acc)[8][8], const float (*Ht)[HS], const float* __restrict__ Wg,
                                          int ch_base, float (*Ws)[256], int warp, int tid) {
  // each thread stages channel `tid` of the 256-channel chunk: 8 consecutive k (two float4)
  const float* wrow = Wg + (int64_t)(ch_base + tid) * K;
  float4 w0 = __ldg(reinterpret_cast<const float4*>(wrow));
  float4 w1 = __ldg(reinterpret_cast<const float4*>(wrow + 4));
  const int lane = tid & 31;
#pragma unroll 1
  for (int k0 = 0; k0 < K; k0 += KT) {
    Ws[0][tid] = w0.x; Ws[1][tid] = w0.y; Ws[2][tid] = w0.z; Ws[3][tid] = w0.w;
    Ws[4][tid] = w1.x; Ws[5][tid] = w1.y; Ws[6][tid] = w1.z; Ws[7][tid] = w1.w;
    __syncthreads();
    if (k0 + KT < K) {   // prefetch the next slice while this one is consumed
      w0 = __ldg(reinterpret_cast<const float4*>(wrow + k0 + KT));
      w1 = __ldg(reinterpret_cast<const float4*>(wrow + k0 + KT + 4));
    }
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&Ht[k0 + k][warp * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&Ht[k0 + k][warp * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = Ws[k][lane + 32 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(FT, 2)
encoder_fwd_fp32(const float* __restrict__ x, int64_t ldx, int N, int C, pm_encoder_params P, int act,
                 float* __restrict__ feat, float* __restrict__ feat_mean, int64_t ldf,
                 int32_t* __restrict__ argmax, float* __restrict__ h2mean) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdSmem& S = *reinterpret_cast<FwdSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const float* xb = x + (int64_t)b * ldx;

  for (int i = tid; i < 128 * C; i += FT) S.W1s[i / C][i % C] = P.W1[i];
  if (tid < 128) S.b1s[tid] = P.b1[tid];
  S.b2s[tid] = P.b2[tid];

  float best[2] = {-INFINITY, -INFINITY};
  int besti[2] = {0, 0};
  float sum3[2] = {0.f, 0.f};
  float sum2 = 0.f;
  float* redv = &S.H1t[0][0];                                   // [8 warps][256]   (8 KB)
  int* redi = reinterpret_cast<int*>(&S.H1t[0][0]) + 8 * 256;   // [8 warps][256]   (8 KB)
  float* reds = &S.H1t[0][0] + 16 * 256;                        // [8 warps][256]   (8 KB)

  for (int p0 = 0; p0 < N; p0 += TP) {
    const int valid = min(TP, N - p0);
    __syncthreads();   // previous tile's reduce scratch (aliases H1t) fully consumed
    for (int i = tid; i < TP * C; i += FT) {
      const int pt = i / C, c = i % C;
      S.xs[pt][c] = (pt < valid) ? __ldg(xb + (int64_t)(p0 + pt) * C + c) : 0.f;
    }
    __syncthreads();
    // ---- layer 1 on CUDA cores (K = C is too small for anything else)
    {
      const int pt = tid & (TP - 1);
      float xv[CMAX];
#pragma unroll
      for (int c = 0; c < CMAX; ++c) xv[c] = (c < C) ? S.xs[pt][c] : 0.f;
      for (int ch = tid / TP; ch < 128; ch += FT / TP) {
        float a = S.b1s[ch];
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
          if (c < C) a = fmaf(xv[c], S.W1s[ch][c], a);
        S.H1t[ch][pt] = pm_act_fwd(act, a);
      }
    }
    __syncthreads();
    // ---- layer 2: [64 x 128] x [128 x 256]
    {
      float acc[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
      tile_gemm<128>(acc, S.H1t, P.W2, 0, S.Ws, warp, tid);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ch = lane + 32 * j;
        const float bb = S.b2s[ch];
        float4 o0, o1;
        o0.x = pm_act_fwd(act, acc[0][j] + bb); o0.y = pm_act_fwd(act, acc[1][j] + bb);
        o0.z = pm_act_fwd(act, acc[2][j] + bb); o0.w = pm_act_fwd(act, acc[3][j] + bb);
        o1.x = pm_act_fwd(act, acc[4][j] + bb); o1.y = pm_act_fwd(act, acc[5][j] + bb);
        o1.z = pm_act_fwd(act, acc[6][j] + bb); o1.w = pm_act_fwd(act, acc[7][j] + bb);
        *reinterpret_cast<float4*>(&S.H2t[ch][warp * 8]) = o0;
        *reinterpret_cast<float4*>(&S.H2t[ch][warp * 8 + 4]) = o1;
      }
    }
    __syncthreads();
    if (h2mean) {
      float t = 0.f;
      for (int pt = 0; pt < valid; ++pt) t += S.H2t[tid][pt];
      sum2 += t;
    }
    // ---- layer 3 in two 256-channel chunks, pooled on the fly
#pragma unroll 1
    for (int chunk = 0; chunk < 2; ++chunk) {
      float acc[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
      tile_gemm<256>(acc, S.H2t, P.W3, chunk * 256, S.Ws, warp, tid);
      // per-thread max over its 8 points (ascending, strict > keeps the first index), then across the 8 warps
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float mv = -INFINITY, sv = 0.f;
        int mi = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int pt = warp * 8 + i;
          if (pt < valid) {
            const float v = acc[i][j];
            sv += v;
            if (v > mv) { mv = v; mi = p0 + pt; }
          }
        }
        const int ch = lane + 32 * j;
        redv[warp * 256 + ch] = mv;
        redi[warp * 256 + ch] = mi;
        reds[warp * 256 + ch] = sv;
      }
      __syncthreads();
      {
        float bv = best[chunk], ss = 0.f;
        int bi = besti[chunk];
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          const float v = redv[w * 256 + tid];
          if (v > bv) { bv = v; bi = redi[w * 256 + tid]; }
          ss += reds[w * 256 + tid];
        }
        best[chunk] = bv; besti[chunk] = bi; sum3[chunk] += ss;
      }
      __syncthreads();
    }
  }
  // ---- pooled outputs (bias after the pool: max_n(h+b) = max_n(h)+b)
#pragma unroll
  for (int chunk = 0; chunk < 2; ++chunk) {
    const int ch = chunk * 256 + tid;
    const float bb = P.b3[ch];
    feat[(int64_t)b * ldf + ch] = best[chunk] + bb;
    if (feat_mean) feat_mean[(int64_t)b * ldf + ch] = sum3[chunk] / (float)N + bb;
    if (argmax) argmax[(int64_t)b * 512 + ch] = besti[chunk];
  }
  if (h2mean) h2mean[(int64_t)b * 256 + tid] = sum2 / (float)N;
}

// ---------------------------------------------------------------- in-place centring (network.py:172-173)
__global__ void __launch_bounds__(256)
center_kernel(float* __restrict__ x, int64_t ldx, int N, int C) {
  __shared__ float sred[32];
  float* xb = x + (int64_t)blockIdx.x * ldx;
  float m[3];
  for (int a = 0; a < 3; ++a) {
    float t = 0.f;
    for (int n = threadIdx.x; n < N; n += blockDim.x) t += xb[(int64_t)n * C + a];
    m[a] = pm_block_sum(t, sred) / (float)N;
  }
  __syncthreads();
  for (int n = threadIdx.x; n < N; n += blockDim.x)
    for (int a = 0; a < 3; ++a) xb[(int64_t)n * C + a] -= m[a];
}

// ---------------------------------------------------------------- critical-point compaction
// pass 1: number of rows each cloud contributes (unique argmax points, or all N points when `all_points`)
__global__ void __launch_bounds__(256)
crit_count_kernel(const int32_t* __restrict__ argmax, int N, int all_points, int32_t* __restrict__ ucount) {
  extern __shared__ int cnt[];
  __shared__ float sred[32];
  const int b = blockIdx.x;
  if (all_points) { if (threadIdx.x == 0) ucount[b] = N; return; }
  for (int n = threadIdx.x; n < N; n += blockDim.x) cnt[n] = 0;
  __syncthreads();
  for (int c = threadIdx.x; c < 512; c += blockDim.x) atomicAdd(&cnt[argmax[(int64_t)b * 512 + c]], 1);
  __syncthreads();
  float u = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) u += (cnt[n] > 0) ? 1.f : 0.f;
  u = pm_block_sum(u, sred);
  if (threadIdx.x == 0) ucount[b] = (int)u;
}

// pass 2: exclusive scan over clouds (B <= a few thousand: one CTA, serial over chunks)
__global__ void __launch_bounds__(1024)
crit_scan_kernel(const int32_t* __restrict__ ucount, int B, int32_t* __restrict__ rowoff, int32_t* __restrict__ r_dev) {
  __shared__ int sm[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < B; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = (i < B) ? ucount[i] : 0;
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int t = (threadIdx.x >= o) ? sm[threadIdx.x - o] : 0;
      __syncthreads();
      sm[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < B) rowoff[i] = carry + sm[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += sm[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) { rowoff[B] = carry; *r_dev = carry; }
}

// pass 3: per cloud, rank its critical points (ascending point index), list each row's channels
// (ascending channel index — keeps every later summation order deterministic) and gather the inputs.
// A channel's place inside its point's segment is the number of lower channels that chose the same point, counted directly
// (512 broadcast reads per thread) — no atomics on the fill order and no per-segment sort (the earlier version insertion-sorted
// every segment in global memory, one thread per point: 120 us per minibatch, most of it the long segments' dependent loads).
__global__ void __launch_bounds__(256)
crit_fill_kernel(const float* __restrict__ x, int64_t ldx, int N, int C, const int32_t* __restrict__ argmax,
                 int all_points, const int32_t* __restrict__ rowoff, int32_t* __restrict__ row_b,
                 int32_t* __restrict__ row_cbeg, int32_t* __restrict__ row_ccnt, int32_t* __restrict__ chan_sorted,
                 int32_t* __restrict__ slot, float* __restrict__ Xc) {
  extern __shared__ int sm[];
  int* cnt = sm;            // [N] channels per point
  int* rank = sm + N;       // [N] row rank of point
  int* cbeg = sm + 2 * N;   // [N] start of the point's channel segment
  int* am = sm + 3 * N;     // [512] this cloud's argmax row
  __shared__ int wsum_r[8], wsum_c[8];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int n = tid; n < N; n += blockDim.x) cnt[n] = 0;
  for (int c = tid; c < 512; c += blockDim.x) am[c] = argmax[(int64_t)b * 512 + c];
  __syncthreads();
  for (int c = tid; c < 512; c += blockDim.x) atomicAdd(&cnt[am[c]], 1);      // counts only: the order of the adds is irrelevant
  __syncthreads();
  // exclusive scans over n of (point is live, channels of the point): each thread owns a run of consecutive points
  const int per = (N + 255) / 256;
  const int n0 = tid * per;
  int fr = 0, fc = 0;
  for (int i = 0; i < per; ++i) {
    const int n = n0 + i;
    if (n < N) { fr += (all_points || cnt[n] > 0) ? 1 : 0; fc += cnt[n]; }
  }
  int ir = fr, ic = fc;                                                       // inclusive warp scans of the thread totals
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int tr = __shfl_up_sync(0xffffffffu, ir, o), tc = __shfl_up_sync(0xffffffffu, ic, o);
    if (lane >= o) { ir += tr; ic += tc; }
  }
  if (lane == 31) { wsum_r[warp] = ir; wsum_c[warp] = ic; }
  __syncthreads();
  int base_r = ir - fr, base_c = ic - fc;
  for (int w = 0; w < warp; ++w) { base_r += wsum_r[w]; base_c += wsum_c[w]; }
  for (int i = 0; i < per; ++i) {
    const int n = n0 + i;
    if (n < N) {
      rank[n] = base_r; cbeg[n] = base_c;
      base_r += (all_points || cnt[n] > 0) ? 1 : 0; base_c += cnt[n];
    }
  }
  __syncthreads();
  const int r0 = rowoff[b];
  for (int n = tid; n < N; n += blockDim.x) {
    if (all_points || cnt[n] > 0) {
      const int r = r0 + rank[n];
      row_b[r] = b;
      row_cbeg[r] = b * 512 + cbeg[n];
      row_ccnt[r] = cnt[n];
      for (int c = 0; c < C; ++c) Xc[(int64_t)r * C + c] = x[(int64_t)b * ldx + (int64_t)n * C + c];
    }
  }
  for (int c = tid; c < 512; c += blockDim.x) {
    const int n = am[c];
    int pos = 0;
    for (int c2 = 0; c2 < c; ++c2) pos += (am[c2] == n) ? 1 : 0;              // warp-uniform address: a broadcast read
    chan_sorted[b * 512 + cbeg[n] + pos] = c;
    slot[b * 512 + c] = r0 + rank[n];
  }
}

// H1c[r,:] = act(W1 x_r + b1)       (rows r < *r_dev).  256 threads = 2 row-lanes x 128 channels, four rows in flight per thread
// (the one-row-at-a-time loop was bound by the latency of its dependent Xc loads: 150 us for 180 k rows).
__global__ void __launch_bounds__(256)
crit_layer1_kernel(const float* __restrict__ Xc, int C, const float* __restrict__ W1, const float* __restrict__ b1,
                   int act, const int32_t* __restrict__ r_dev, float* __restrict__ H1c) {
  const int R = *r_dev, ch = threadIdx.x & 127, half = threadIdx.x >> 7;
  float w[CMAX];
  for (int c = 0; c < CMAX; ++c) w[c] = (c < C) ? W1[ch * C + c] : 0.f;
  const float bb = b1[ch];
  const int stride = gridDim.x * 2;
  for (int r0 = blockIdx.x * 2 + half; r0 < R; r0 += 4 * stride) {
    float a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + u * stride;
      a[u] = bb;
      if (r < R)
        for (int c = 0; c < C; ++c) a[u] = fmaf(__ldg(Xc + (int64_t)r * C + c), w[c], a[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + u * stride;
      if (r < R) H1c[(int64_t)r * 128 + ch] = pm_act_fwd_fast(act, a[u]);
    }
  }
}

// dPre2[r,k] = (sum_{c in chan(r)} dfeat[b,c] W3[c,k]  [+ gm[b,k]]) * act'(H2c[r,k]);
// four rows per CTA (64 threads x float4 over k per row) so that each SM keeps many W3-row gathers (L2) in flight; per element the
// channels are summed in ascending order (deterministic)
__global__ void __launch_bounds__(256)
crit_dh2_v4_kernel(const float* __restrict__ dfeat, int64_t lddf, const float* __restrict__ W3,
                   const float* __restrict__ H2c, const int32_t* __restrict__ row_b, const int32_t* __restrict__ row_cbeg,
                   const int32_t* __restrict__ row_ccnt, const int32_t* __restrict__ chan_sorted,
                   const float* __restrict__ gm, int act, const int32_t* __restrict__ r_dev, float* __restrict__ dPre2) {
  const int R = *r_dev, sub = threadIdx.x >> 6, l = threadIdx.x & 63;
  for (int r = blockIdx.x * 4 + sub; r < R; r += gridDim.x * 4) {
    const int b = row_b[r], beg = row_cbeg[r], q = row_ccnt[r];
    float4 a = gm ? __ldg(reinterpret_cast<const float4*>(gm + (int64_t)b * 256) + l) : make_float4(0.f, 0.f, 0.f, 0.f);
    int j = 0;
    for (; j + 2 <= q; j += 2) {
      const int c0 = chan_sorted[beg + j], c1 = chan_sorted[beg + j + 1];
      const float g0 = dfeat[(int64_t)b * lddf + c0], g1 = dfeat[(int64_t)b * lddf + c1];
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(W3 + (int64_t)c0 * 256) + l);
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(W3 + (int64_t)c1 * 256) + l);
      a.x = fmaf(g0, w0.x, a.x); a.y = fmaf(g0, w0.y, a.y); a.z = fmaf(g0, w0.z, a.z); a.w = fmaf(g0, w0.w, a.w);
      a.x = fmaf(g1, w1.x, a.x); a.y = fmaf(g1, w1.y, a.y); a.z = fmaf(g1, w1.z, a.z); a.w = fmaf(g1, w1.w, a.w);
    }
    if (j < q) {
      const int c0 = chan_sorted[beg + j];
      const float g0 = dfeat[(int64_t)b * lddf + c0];
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(W3 + (int64_t)c0 * 256) + l);
      a.x = fmaf(g0, w0.x, a.x); a.y = fmaf(g0, w0.y, a.y); a.z = fmaf(g0, w0.z, a.z); a.w = fmaf(g0, w0.w, a.w);
    }
    const float4 h = *(reinterpret_cast<const float4*>(H2c + (int64_t)r * 256) + l);
    *(reinterpret_cast<float4*>(dPre2 + (int64_t)r * 256) + l) =
        make_float4(a.x * pm_act_bwd(act, h.x), a.y * pm_act_bwd(act, h.y), a.z * pm_act_bwd(act, h.z), a.w * pm_act_bwd(act, h.w));
  }
}

// layer-1 gradients over the compacted rows: part[slab][ch][0..C-1] = sum_r dPre1[r,ch] * Xc[r,c], part[slab][ch][CMAX] = sum_r dPre1[r,ch].
// 256 threads = 2 row-lanes x 128 h1 channels; rows dealt to CTAs in contiguous slabs; four rows in flight per thread.
constexpr int DW1_SLABS = 8 * PM_NUM_SMS;
__global__ void __launch_bounds__(256)
crit_dw1_kernel(const float* __restrict__ dPre1, const float* __restrict__ Xc, int C, const int32_t* __restrict__ r_dev,
                float* __restrict__ part) {
  __shared__ float red[128][CMAX + 1];
  const int R = *r_dev, ch = threadIdx.x & 127, half = threadIdx.x >> 7;
  const int per = ((R + (int)gridDim.x - 1) / (int)gridDim.x + 7) / 8 * 8;
  const int r0 = min(R, (int)blockIdx.x * per), r1 = min(R, r0 + per);
  float acc[CMAX + 1];
#pragma unroll
  for (int c = 0; c <= CMAX; ++c) acc[c] = 0.f;
  int r = r0 + half;
  for (; r + 6 < r1; r += 8) {                                  // rows r, r+2, r+4, r+6 of this row-lane
    float d[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) d[u] = dPre1[(int64_t)(r + 2 * u) * 128 + ch];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) acc[c] = fmaf(d[u], __ldg(Xc + (int64_t)(r + 2 * u) * C + c), acc[c]);
      acc[CMAX] += d[u];
    }
  }
  for (; r < r1; r += 2) {
    const float d = dPre1[(int64_t)r * 128 + ch];
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) acc[c] = fmaf(d, __ldg(Xc + (int64_t)r * C + c), acc[c]);
    acc[CMAX] += d;
  }
  if (half == 1) {
#pragma unroll
    for (int c = 0; c <= CMAX; ++c) red[ch][c] = acc[c];
  }
  __syncthreads();
  if (half == 0) {
    float* out = part + ((int64_t)blockIdx.x * 128 + ch) * (CMAX + 1);
#pragma unroll
    for (int c = 0; c <= CMAX; ++c) out[c] = acc[c] + red[ch][c];
  }
}
// fixed-order sum of the slabs -> dW1 (128, C) and db1 (128): one warp per output, lanes stride the slabs, shuffle tree
__global__ void __launch_bounds__(256)
crit_dw1_reduce_kernel(const float* __restrict__ part, int slabs, int C, float* __restrict__ dW1, float* __restrict__ db1) {
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= 128 * (CMAX + 1)) return;
  const int ch = i / (CMAX + 1), c = i % (CMAX + 1);
  if (c >= C && c != CMAX) return;
  float t = 0.f;
  for (int s = lane; s < slabs; s += 32) t += part[((int64_t)s * 128 + ch) * (CMAX + 1) + c];
  t = pm_warp_sum(t);
  if (lane == 0) { if (c == CMAX) db1[ch] = t; else dW1[ch * C + c] = t; }
}

// dW3 partial over a slab of clouds: part[slab][c][k] = sum_{b in slab} dfeat[b,c] * H2c[slot[b,c]][k]
constexpr int DW3_CH = 4;   // channels per CTA
__global__ void __launch_bounds__(256)
crit_dw3_kernel(const float* __restrict__ dfeat, int64_t lddf, const float* __restrict__ H2c,
                const int32_t* __restrict__ slot, int B, int b_per_slab, float* __restrict__ part,
                float* __restrict__ dbpart) {
  const int k = threadIdx.x;
  const int c0 = blockIdx.x * DW3_CH;
  const int b0 = blockIdx.y * b_per_slab, b1 = min(B, b0 + b_per_slab);
  float acc[DW3_CH], dbs[DW3_CH];
#pragma unroll
  for (int j = 0; j < DW3_CH; ++j) { acc[j] = 0.f; dbs[j] = 0.f; }
  for (int b = b0; b < b1; ++b) {
#pragma unroll
    for (int j = 0; j < DW3_CH; ++j) {
      const float g = dfeat[(int64_t)b * lddf + c0 + j];
      const int r = slot[b * 512 + c0 + j];
      acc[j] = fmaf(g, H2c[(int64_t)r * 256 + k], acc[j]);
      dbs[j] += g;
    }
  }
#pragma unroll
  for (int j = 0; j < DW3_CH; ++j) {
    part[((int64_t)blockIdx.y * 512 + c0 + j) * 256 + k] = acc[j];
    if (k == 0) dbpart[(int64_t)blockIdx.y * 512 + c0 + j] = dbs[j];
  }
}

__global__ void reduce_slabs_kernel(const float* __restrict__ part, int slabs, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float t = 0.f;
  for (int s = 0; s < slabs; ++s) t += part[(int64_t)s * n + i];
  out[i] = t;
}

// mean-pool branch (max_mean=True, network.py:180-182): every point receives dfeat_mean[b,c]/N through layer 3
// gm[b,k] = (1/N) sum_c dfeat_mean[b,c] W3[c,k]   — the part of dH2 that is common to all points of cloud b
__global__ void __launch_bounds__(256)
mean_gm_kernel(const float* __restrict__ dfm, int64_t lddf, const float* __restrict__ W3, float inv_n, float* __restrict__ gm) {
  __shared__ float sd[512];
  const int b = blockIdx.x, k = threadIdx.x;
  for (int c = k; c < 512; c += 256) sd[c] = dfm[(int64_t)b * lddf + c];
  __syncthreads();
  float a = 0.f;
  for (int c = 0; c < 512; ++c) a = fmaf(sd[c], __ldg(W3 + (int64_t)c * 256 + k), a);
  gm[(int64_t)b * 256 + k] = a * inv_n;
}
// dW3[c,k] += sum_b dfeat_mean[b,c] h2mean[b,k] ; db3[c] += sum_b dfeat_mean[b,c]   (mean_n h3 = W3 mean_n h2 + b3)
__global__ void __launch_bounds__(256)
mean_dw3_kernel(const float* __restrict__ dfm, int64_t lddf, const float* __restrict__ h2mean, int B, float* __restrict__ dW3,
                float* __restrict__ db3) {
  const int k = threadIdx.x, c0 = blockIdx.x * DW3_CH;
  float acc[DW3_CH], dbs[DW3_CH];
#pragma unroll
  for (int j = 0; j < DW3_CH; ++j) { acc[j] = 0.f; dbs[j] = 0.f; }
  for (int b = 0; b < B; ++b) {
    const float h = h2mean[(int64_t)b * 256 + k];
#pragma unroll
    for (int j = 0; j < DW3_CH; ++j) {
      const float g = dfm[(int64_t)b * lddf + c0 + j];
      acc[j] = fmaf(g, h, acc[j]);
      dbs[j] += g;
    }
  }
#pragma unroll
  for (int j = 0; j < DW3_CH; ++j) {
    dW3[(int64_t)(c0 + j) * 256 + k] += acc[j];
    if (k == 0) db3[c0 + j] += dbs[j];
  }
}

struct BwdWs {
  int32_t *ucount, *rowoff, *r_dev, *row_b, *row_cbeg, *row_ccnt, *chan_sorted, *slot;
  float *Xc, *H1c, *H2c, *dPre2, *dPre1, *dw3part, *db3part, *gm, *lin;
  size_t lin_bytes, total;
  int slabs, b_per_slab;
  int64_t rmax;
};

inline BwdWs carve_bwd(void* ws, int B, int N, int C, int with_mean) {
  BwdWs w{};
  const int64_t rmax = (int64_t)B * (with_mean ? N : (N < 512 ? N : 512));
  w.rmax = rmax;
  w.slabs = B >= 64 ? 8 : 1;
  w.b_per_slab = pm_cdiv(B, w.slabs);
  size_t off = 0;
  char* base = reinterpret_cast<char*>(ws);
  auto take = [&](size_t bytes) { void* p = base ? base + off : nullptr; off += pm_align_up(bytes, 256); return p; };
  w.ucount = (int32_t*)take((size_t)B * 4);
  w.rowoff = (int32_t*)take((size_t)(B + 1) * 4);
  w.r_dev = (int32_t*)take(4);
  w.row_b = (int32_t*)take((size_t)rmax * 4);
  w.row_cbeg = (int32_t*)take((size_t)rmax * 4);
  w.row_ccnt = (int32_t*)take((size_t)rmax * 4);
  w.chan_sorted = (int32_t*)take((size_t)B * 512 * 4);
  w.slot = (int32_t*)take((size_t)B * 512 * 4);
  w.Xc = (float*)take((size_t)rmax * C * 4);
  w.H1c = (float*)take((size_t)rmax * 128 * 4);
  w.H2c = (float*)take((size_t)rmax * 256 * 4);
  w.dPre2 = (float*)take((size_t)rmax * 256 * 4);
  w.dPre1 = (float*)take((size_t)rmax * 128 * 4);
  w.dw3part = (float*)take((size_t)w.slabs * 512 * 256 * 4);
  w.db3part = (float*)take((size_t)w.slabs * 512 * 4);
  w.gm = (float*)take(with_mean ? (size_t)B * 256 * 4 : 0);
  const int rm = (int)(rmax > INT32_MAX ? INT32_MAX : rmax);
  size_t l1 = pm_linear_backward_ws_bytes(rm, 256, 128);
  size_t l2 = pm_linear_backward_ws_bytes(rm, 128, C);
  size_t l3 = pm_linear_backward_tc_ws_bytes(rm, 256, 128);
  w.lin_bytes = l1 > l2 ? l1 : l2;
  if (l3 > w.lin_bytes) w.lin_bytes = l3;
  const size_t l4 = (size_t)DW1_SLABS * 128 * (CMAX + 1) * sizeof(float);
  if (l4 > w.lin_bytes) w.lin_bytes = l4;
  w.lin = (float*)take(w.lin_bytes);
  w.total = off;
  return w;
}

}  // namespace

extern "C" {

int pm_pointnet_center(float* x, int64_t ldx, int B, int N, int C, pm_stream_t s) {
  PM_REQUIRE(x && B > 0 && N > 0 && C >= 3, PM_ERR_ARG, "pm_pointnet_center: need C>=3 (xyz), got B=%d N=%d C=%d", B, N, C);
  center_kernel<<<B, 256, 0, pm_st(s)>>>(x, ldx, N, C);
  PM_CHECK_LAUNCH("pm_pointnet_center");
  return PM_OK;
}

size_t pm_pointnet_encode_forward_ws_bytes(int B, int N, int C, int precision) {
  if (precision == PM_PREC_BF16) return pm_pointnet_encode_forward_tc_ws_bytes(B, N, C);
  if (precision == PM_PREC_FP32 && pm_pointnet_encode_forward_tc3_supported(N, C)) return pm_pointnet_encode_forward_tc3_ws_bytes(B, N, C);
  return 0;
}

int pm_pointnet_encode_forward(const float* x, int64_t ldx, int B, int N, int C, const pm_encoder_params* p, int act,
                               int precision, float* feat, float* feat_mean, int64_t ldf, int32_t* argmax,
                               float* h2mean, void* ws, size_t ws_bytes, pm_stream_t s) {
  PM_REQUIRE(x && p && feat, PM_ERR_ARG, "pm_pointnet_encode_forward: null pointer");
  PM_REQUIRE(B > 0 && N > 0 && C >= 1 && C <= CMAX, PM_ERR_SHAPE, "pm_pointnet_encode_forward: B=%d N=%d C=%d (C<=%d)", B, N, C, CMAX);
  PM_REQUIRE(ldx >= (int64_t)N * C && ldf >= 512, PM_ERR_SHAPE, "pm_pointnet_encode_forward: bad strides");
  PM_REQUIRE(act >= PM_ACT_NONE && act <= PM_ACT_SIGMOID, PM_ERR_ARG, "pm_pointnet_encode_forward: activation %d", act);
  PM_REQUIRE(pm_aligned(p->W2, 16) && pm_aligned(p->W3, 16), PM_ERR_ALIGN, "pm_pointnet_encode_forward: W2/W3 must be 16-byte aligned");
  if (precision == PM_PREC_BF16) {
    PM_REQUIRE(!feat_mean && !h2mean, PM_ERR_UNSUPPORTED, "bf16 encoder: max_mean pooling runs in PM_PREC_FP32 only");
    return pm_pointnet_encode_forward_tc(x, ldx, B, N, C, p, act, feat, ldf, argmax, ws, ws_bytes, s);
  }
  PM_REQUIRE(precision == PM_PREC_FP32 || precision == PM_PREC_FP32_FFMA, PM_ERR_ARG, "pm_pointnet_encode_forward: precision %d", precision);
  // fp32 parity mode: the split-fp16 tcgen05 kernel where it applies (max pooling only, N % 256 == 0, C <= 4), else CUDA cores
  if (precision == PM_PREC_FP32 && !feat_mean && !h2mean && pm_pointnet_encode_forward_tc3_supported(N, C))
    return pm_pointnet_encode_forward_tc3(x, ldx, B, N, C, p, act, feat, ldf, argmax, ws, ws_bytes, s);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(encoder_fwd_fp32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FwdSmem));
    if (e != cudaSuccess) PM_FAIL(PM_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  encoder_fwd_fp32<<<B, FT, sizeof(FwdSmem), pm_st(s)>>>(x, ldx, N, C, *p, act, feat, feat_mean, ldf, argmax, h2mean);
  PM_CHECK_LAUNCH("pm_pointnet_encode_forward");
  return PM_OK;
}

size_t pm_pointnet_encode_backward_ws_bytes(int B, int N, int C, int with_mean, int precision) {
  if (precision == PM_PREC_BF16 && !with_mean) return pm_pointnet_encode_backward_tc_ws_bytes(B, N, C);
  return carve_bwd(nullptr, B, N, C, with_mean).total;
}

int pm_pointnet_encode_backward(const float* x, int64_t ldx, int B, int N, int C, const pm_encoder_params* p, int act,
                                int precision, const float* dfeat, const float* dfeat_mean, int64_t lddf,
                                const int32_t* argmax, const float* h2mean, const pm_encoder_grads* g, void* ws,
                                size_t ws_bytes, pm_stream_t s) {
  PM_REQUIRE(x && p && dfeat && argmax && g && ws, PM_ERR_ARG, "pm_pointnet_encode_backward: null pointer");
  PM_REQUIRE(B > 0 && N > 0 && N <= 8192 && C >= 1 && C <= CMAX, PM_ERR_SHAPE, "pm_pointnet_encode_backward: B=%d N=%d C=%d", B, N, C);
  PM_REQUIRE(act >= PM_ACT_NONE && act <= PM_ACT_SIGMOID, PM_ERR_ARG, "pm_pointnet_encode_backward: activation %d", act);
  if (precision == PM_PREC_BF16) {
    PM_REQUIRE(!dfeat_mean, PM_ERR_UNSUPPORTED, "bf16 encoder backward: max_mean pooling runs in PM_PREC_FP32 only");
    return pm_pointnet_encode_backward_tc(x, ldx, B, N, C, p, act, dfeat, lddf, argmax, g, ws, ws_bytes, s);
  }
  PM_REQUIRE(precision == PM_PREC_FP32 || precision == PM_PREC_FP32_FFMA, PM_ERR_ARG, "pm_pointnet_encode_backward: precision %d", precision);
  const bool tc = precision == PM_PREC_FP32;     // the two 128<->256 layers on tcgen05 (three-term bf16 split, 1e-4 gate)
  PM_REQUIRE(!dfeat_mean || h2mean, PM_ERR_ARG, "pm_pointnet_encode_backward: the mean-pool branch needs h2mean from the forward");
  const int with_mean = dfeat_mean != nullptr;
  BwdWs w = carve_bwd(ws, B, N, C, with_mean);
  PM_REQUIRE(ws_bytes >= w.total, PM_ERR_ARG, "pm_pointnet_encode_backward: workspace %zu < %zu", ws_bytes, w.total);
  PM_REQUIRE(w.rmax <= INT32_MAX, PM_ERR_SHAPE, "pm_pointnet_encode_backward: too many rows");
  cudaStream_t st = pm_st(s);
  const int rmax = (int)w.rmax;
  int rc;
  // 1. compact the critical points
  crit_count_kernel<<<B, 256, (size_t)N * 4, st>>>(argmax, N, with_mean, w.ucount);
  crit_scan_kernel<<<1, 1024, 0, st>>>(w.ucount, B, w.rowoff, w.r_dev);
  crit_fill_kernel<<<B, 256, (size_t)N * 12 + 2048, st>>>(x, ldx, N, C, argmax, with_mean, w.rowoff, w.row_b, w.row_cbeg,
                                                   w.row_ccnt, w.chan_sorted, w.slot, w.Xc);
  // 2. recompute their activations
  const int grid_rows = 8 * PM_NUM_SMS;
  crit_layer1_kernel<<<grid_rows, 256, 0, st>>>(w.Xc, C, p->W1, p->b1, act, w.r_dev, w.H1c);
  if (tc) rc = pm_linear_forward_tc(w.H1c, 128, p->W2, p->b2, w.H2c, 256, rmax, 256, 128, act, PM_PREC_FP32, w.r_dev, s);
  else rc = pm_linear_forward(w.H1c, 128, p->W2, p->b2, w.H2c, 256, rmax, 256, 128, act, w.r_dev, s);
  if (rc) return rc;
  // 3. layer 3: dPre2 rows and dW3/db3
  if (with_mean) mean_gm_kernel<<<B, 256, 0, st>>>(dfeat_mean, lddf, p->W3, 1.f / (float)N, w.gm);
  crit_dh2_v4_kernel<<<grid_rows, 256, 0, st>>>(dfeat, lddf, p->W3, w.H2c, w.row_b, w.row_cbeg, w.row_ccnt, w.chan_sorted,
                                                 with_mean ? w.gm : nullptr, act, w.r_dev, w.dPre2);
  crit_dw3_kernel<<<dim3(512 / DW3_CH, w.slabs), 256, 0, st>>>(dfeat, lddf, w.H2c, w.slot, B, w.b_per_slab, w.dw3part,
                                                                w.db3part);
  reduce_slabs_kernel<<<pm_cdiv(512 * 256, 256), 256, 0, st>>>(w.dw3part, w.slabs, 512 * 256, g->W3);
  reduce_slabs_kernel<<<2, 256, 0, st>>>(w.db3part, w.slabs, 512, g->b3);
  if (with_mean) mean_dw3_kernel<<<512 / DW3_CH, 256, 0, st>>>(dfeat_mean, lddf, h2mean, B, g->W3, g->b3);
  PM_CHECK_LAUNCH("pm_pointnet_encode_backward/crit");
  // 4. layer 2: dW2, db2, dPre1 = (dPre2 W2) * act'(H1c)
  if (tc) rc = pm_linear_backward_tc(w.H1c, 128, p->W2, w.dPre2, 256, g->W2, g->b2, w.dPre1, 128, rmax, 256, 128, act, PM_PREC_FP32,
                                     w.r_dev, w.lin, s);
  else rc = pm_linear_backward(w.H1c, 128, p->W2, w.dPre2, 256, g->W2, g->b2, w.dPre1, 128, rmax, 256, 128, act, w.r_dev, w.lin, s);
  if (rc) return rc;
  // 5. layer 1: dW1, db1 (no dx: the cloud is an input)
  if (tc) {
    crit_dw1_kernel<<<DW1_SLABS, 256, 0, st>>>(w.dPre1, w.Xc, C, w.r_dev, w.lin);
    crit_dw1_reduce_kernel<<<pm_cdiv(128 * (CMAX + 1), 8), 256, 0, st>>>(w.lin, DW1_SLABS, C, g->W1, g->b1);
    PM_CHECK_LAUNCH("pm_pointnet_encode_backward/dw1");
  } else if ((rc = pm_linear_backward(w.Xc, C, p->W1, w.dPre1, 128, g->W1, g->b1, nullptr, 0, rmax, 128, C, PM_ACT_NONE,
                                      w.r_dev, w.lin, s))) return rc;
  return PM_OK;
}

}  // extern "C"
