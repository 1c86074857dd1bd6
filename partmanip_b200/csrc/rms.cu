// K6 — running mean/std observation normaliser (reference: algorithms/algo_utils/RMS.py:10-18, 40-45).
// HBM-bound streaming kernels: x is (E, D) fp32 row-major (50 MB at E=4096, D=3072).
//   colsum / colsqdev : deterministic two-stage column reduction (slab partials, then fixed-order sum)
//   update            : the reference's non-standard S recurrence, op order preserved
//   normalize         : (x - mean) / std, true division, no epsilon
#include "common.cuh"

namespace {

constexpr int CR_TX = 32;   // threads along columns
constexpr int CR_TY = 8;    // row lanes

// MODE 0: sum x ; MODE 1: sum (x - colsum/count)^2 ; VEC = 4 (float4) or 1
template <int MODE, int VEC>
__global__ void __launch_bounds__(CR_TX* CR_TY)
colreduce_partial(const float* __restrict__ x, int64_t ld, int E, int D, const float* __restrict__ colsum,
                  float count, int rows_per_slab, float* __restrict__ partial) {
  __shared__ float sm[CR_TY][CR_TX * VEC + 1];
  const int col0 = (blockIdx.x * CR_TX + threadIdx.x) * VEC;
  const int r_begin = blockIdx.y * rows_per_slab;
  const int r_end = min(E, r_begin + rows_per_slab);
  float acc[VEC];
  float mu[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    acc[v] = 0.f;
    mu[v] = 0.f;
  }
  if (col0 < D) {
    if (MODE == 1) {
#pragma unroll
      for (int v = 0; v < VEC; ++v)
        if (col0 + v < D) mu[v] = __fdiv_rn(colsum[col0 + v], count);
    }
    // four rows' loads are issued before the first add (HBM latency needs ~40 KB in flight per SM); the adds keep row order
    constexpr int UR = 4;
    for (int r0 = r_begin + threadIdx.y; r0 < r_end; r0 += CR_TY * UR) {
      float val[UR][VEC];
#pragma unroll
      for (int u = 0; u < UR; ++u) {
        const int r = r0 + u * CR_TY;
        if (r < r_end) {
          const float* p = x + (int64_t)r * ld + col0;
          if (VEC == 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p));
            val[u][0] = t.x; val[u][1 % VEC] = t.y; val[u][2 % VEC] = t.z; val[u][3 % VEC] = t.w;
          } else {
            val[u][0] = __ldg(p);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UR; ++u) {
        if (r0 + u * CR_TY < r_end) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            if (MODE == 0) acc[v] += val[u][v];
            else { const float d = val[u][v] - mu[v]; acc[v] = fmaf(d, d, acc[v]); }
          }
        }
      }
    }
  }
#pragma unroll
  for (int v = 0; v < VEC; ++v) sm[threadIdx.y][threadIdx.x * VEC + v] = acc[v];
  __syncthreads();
  // fixed-order reduction over the row lanes
  for (int c = threadIdx.y * CR_TX + threadIdx.x; c < CR_TX * VEC; c += CR_TX * CR_TY) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < CR_TY; ++y) t += sm[y][c];
    const int col = blockIdx.x * CR_TX * VEC + c;
    if (col < D) partial[(int64_t)blockIdx.y * D + col] = t;
  }
}

__global__ void colreduce_final(const float* __restrict__ partial, int slabs, int D, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  float t = 0.f;
  for (int s = 0; s < slabs; ++s) t += partial[(int64_t)s * D + c];
  out[c] = t;
}

__global__ void rms_update_kernel(float* mean, float* S, float* stdv, const float* __restrict__ colsum,
                                  const float* __restrict__ sqdev, float count, int n, int D) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  const float fn = (float)n;
  const float old_mean = mean[d];
  const float new_mean = __fdiv_rn(colsum[d], count);                       // x.mean(dim=0)
  const float m = __fadd_rn(old_mean, __fdiv_rn(__fsub_rn(new_mean, old_mean), fn));
  const float var_b = __fdiv_rn(sqdev[d], count);                           // (x-new_mean).pow(2).mean(0)
  const float dm = __fsub_rn(old_mean, new_mean);
  const float t = __fdiv_rn(__fmul_rn(__fmul_rn(dm, dm), (float)(n - 1)), fn);
  const float s = __fadd_rn(__fadd_rn(S[d], var_b), t);
  mean[d] = m;
  S[d] = s;
  stdv[d] = __fsqrt_rn(__fdiv_rn(s, fn));
}

template <int VEC>
__global__ void __launch_bounds__(256)
rms_normalize_kernel(const float* __restrict__ x, int64_t ldx, float* __restrict__ out, int64_t ldo, int E,
                     int D, const float* __restrict__ mean, const float* __restrict__ stdv) {
  // grid.x over column groups, grid.y strides rows: mean/std for this thread's columns stay in registers
  const int col0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (col0 >= D) return;
  float mu[VEC], sd[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    mu[v] = (col0 + v < D) ? mean[col0 + v] : 0.f;
    sd[v] = (col0 + v < D) ? stdv[col0 + v] : 1.f;
  }
  if (VEC == 4) {                                           // two rows per step: both loads issued before the divisions
    for (int r = blockIdx.y; r < E; r += 2 * gridDim.y) {
      const int r2 = r + gridDim.y;
      float4 t = __ldg(reinterpret_cast<const float4*>(x + (int64_t)r * ldx + col0)), t2 = t;
      if (r2 < E) t2 = __ldg(reinterpret_cast<const float4*>(x + (int64_t)r2 * ldx + col0));
      t.x = __fdiv_rn(__fsub_rn(t.x, mu[0]), sd[0]);
      t.y = __fdiv_rn(__fsub_rn(t.y, mu[1 % VEC]), sd[1 % VEC]);
      t.z = __fdiv_rn(__fsub_rn(t.z, mu[2 % VEC]), sd[2 % VEC]);
      t.w = __fdiv_rn(__fsub_rn(t.w, mu[3 % VEC]), sd[3 % VEC]);
      *reinterpret_cast<float4*>(out + (int64_t)r * ldo + col0) = t;
      if (r2 < E) {
        t2.x = __fdiv_rn(__fsub_rn(t2.x, mu[0]), sd[0]);
        t2.y = __fdiv_rn(__fsub_rn(t2.y, mu[1 % VEC]), sd[1 % VEC]);
        t2.z = __fdiv_rn(__fsub_rn(t2.z, mu[2 % VEC]), sd[2 % VEC]);
        t2.w = __fdiv_rn(__fsub_rn(t2.w, mu[3 % VEC]), sd[3 % VEC]);
        *reinterpret_cast<float4*>(out + (int64_t)r2 * ldo + col0) = t2;
      }
    }
  } else {
    for (int r = blockIdx.y; r < E; r += gridDim.y)
      out[(int64_t)r * ldo + col0] = __fdiv_rn(__fsub_rn(__ldg(x + (int64_t)r * ldx + col0), mu[0]), sd[0]);
  }
}

inline int slabs_for(int E, int D, int vec) {
  const int col_blocks = pm_cdiv(D, CR_TX * vec);
  int s = pm_cdiv(4 * PM_NUM_SMS, col_blocks);
  const int max_s = pm_cdiv(E, CR_TY * 4);
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  return s;
}
inline bool can_vec4(const void* p, int64_t ld, int D) { return (D % 4 == 0) && (ld % 4 == 0) && pm_aligned(p, 16); }

template <int MODE>
int colreduce(const float* x, int64_t ldx, int E, int D, const float* colsum, float count, float* out, void* ws,
              cudaStream_t st) {
  PM_REQUIRE(E > 0 && D > 0 && x && out && ws, PM_ERR_ARG, "colreduce: bad args");
  // slab count must match pm_colreduce_ws_bytes (computed for vec=1, the larger of the two)
  const bool v4 = can_vec4(x, ldx, D);
  const int slabs = slabs_for(E, D, 1);
  const int rps = pm_cdiv(E, slabs);
  float* partial = reinterpret_cast<float*>(ws);
  dim3 blk(CR_TX, CR_TY);
  if (v4) {
    dim3 grd(pm_cdiv(D, CR_TX * 4), slabs);
    colreduce_partial<MODE, 4><<<grd, blk, 0, st>>>(x, ldx, E, D, colsum, count, rps, partial);
  } else {
    dim3 grd(pm_cdiv(D, CR_TX), slabs);
    colreduce_partial<MODE, 1><<<grd, blk, 0, st>>>(x, ldx, E, D, colsum, count, rps, partial);
  }
  colreduce_final<<<pm_cdiv(D, 256), 256, 0, st>>>(partial, slabs, D, out);
  PM_CHECK_LAUNCH("colreduce");
  return PM_OK;
}

}  // namespace

extern "C" {

size_t pm_colreduce_ws_bytes(int rows, int cols) {
  if (rows <= 0 || cols <= 0) return 0;
  return (size_t)slabs_for(rows, cols, 1) * cols * sizeof(float);
}

int pm_rms_colsum(const float* x, int64_t ldx, int E, int D, float* colsum, void* ws, pm_stream_t s) {
  return colreduce<0>(x, ldx, E, D, nullptr, 1.f, colsum, ws, pm_st(s));
}

int pm_rms_colsqdev(const float* x, int64_t ldx, int E, int D, const float* colsum, float count, float* sqdev,
                    void* ws, pm_stream_t s) {
  PM_REQUIRE(colsum && count > 0, PM_ERR_ARG, "pm_rms_colsqdev: bad args");
  return colreduce<1>(x, ldx, E, D, colsum, count, sqdev, ws, pm_st(s));
}

int pm_rms_update(float* mean, float* S, float* stdv, const float* colsum, const float* sqdev, float count, int n,
                  int D, pm_stream_t s) {
  PM_REQUIRE(mean && S && stdv && colsum && sqdev && n >= 1 && D > 0 && count > 0, PM_ERR_ARG, "pm_rms_update: bad args");
  rms_update_kernel<<<pm_cdiv(D, 256), 256, 0, pm_st(s)>>>(mean, S, stdv, colsum, sqdev, count, n, D);
  PM_CHECK_LAUNCH("pm_rms_update");
  return PM_OK;
}

int pm_rms_normalize(const float* x, int64_t ldx, float* out, int64_t ldo, int E, int D, const float* mean,
                     const float* stdv, pm_stream_t s) {
  PM_REQUIRE(x && out && mean && stdv && E > 0 && D > 0, PM_ERR_ARG, "pm_rms_normalize: bad args");
  const bool v4 = can_vec4(x, ldx, D) && can_vec4(out, ldo, D);
  const int vec = v4 ? 4 : 1;
  const int col_blocks = pm_cdiv(D, 256 * vec);
  int gy = pm_cdiv(8 * PM_NUM_SMS, col_blocks);
  if (gy > E) gy = E;
  dim3 grd(col_blocks, gy);
  if (v4) rms_normalize_kernel<4><<<grd, 256, 0, pm_st(s)>>>(x, ldx, out, ldo, E, D, mean, stdv);
  else rms_normalize_kernel<1><<<grd, 256, 0, pm_st(s)>>>(x, ldx, out, ldo, E, D, mean, stdv);
  PM_CHECK_LAUNCH("pm_rms_normalize");
  return PM_OK;
}

size_t pm_rms_forward_ws_bytes(int E, int D) {
  return pm_align_up((size_t)2 * D * sizeof(float), 256) + pm_colreduce_ws_bytes(E, D);
}

int pm_rms_forward(const float* x, int64_t ldx, float* out, int64_t ldo, int E, int D, float* mean, float* S,
                   float* stdv, int n_after, int update, void* scratch, pm_stream_t s) {
  if (update) {
    PM_REQUIRE(scratch, PM_ERR_ARG, "pm_rms_forward: scratch required");
    float* colsum = reinterpret_cast<float*>(scratch);
    float* sqdev = colsum + D;
    void* ws = reinterpret_cast<char*>(scratch) + pm_align_up((size_t)2 * D * sizeof(float), 256);
    int rc;
    if ((rc = pm_rms_colsum(x, ldx, E, D, colsum, ws, s))) return rc;
    if ((rc = pm_rms_colsqdev(x, ldx, E, D, colsum, (float)E, sqdev, ws, s))) return rc;
    if ((rc = pm_rms_update(mean, S, stdv, colsum, sqdev, (float)E, n_after, D, s))) return rc;
  }
  return pm_rms_normalize(x, ldx, out, ldo, E, D, mean, stdv, s);
}

}  // extern "C"
