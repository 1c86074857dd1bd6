// K8 — rollout-buffer helpers.  reference: algorithms/algo_utils/storage.py:43-56 (add_transitions),
// :84-91 (add_transitions_dagger) and the x[list] gathers of algorithms/ppo.py:317-324 (sampler=random).
// With the sequential sampler a minibatch is a contiguous slice of the (T*E, D) buffer — no kernel at all.
#include "common.cuh"

namespace {

template <int VEC>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ src, int64_t lds, const int64_t* __restrict__ idx,
                   float* __restrict__ out, int64_t ldo, int64_t n_rows, int width) {
  for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
    const int64_t sr = idx ? idx[r] : r;
    const float* p = src + sr * lds;
    float* q = out + r * ldo;
    if (VEC == 4) {                                        // four 16-byte loads in flight per thread before the first store
      for (int c0 = threadIdx.x * 4; c0 < width; c0 += blockDim.x * 16) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + u * blockDim.x * 4;
          if (c < width) v[u] = __ldg(reinterpret_cast<const float4*>(p + c));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + u * blockDim.x * 4;
          if (c < width) *reinterpret_cast<float4*>(q + c) = v[u];
        }
      }
    } else {
      for (int c = threadIdx.x; c < width; c += blockDim.x) q[c] = __ldg(p + c);
    }
  }
}

int launch_rows(const float* src, int64_t lds, const int64_t* idx, float* out, int64_t ldo, int64_t n_rows,
                int width, cudaStream_t st) {
  const bool v4 = (width % 4 == 0) && (lds % 4 == 0) && (ldo % 4 == 0) && pm_aligned(src, 16) && pm_aligned(out, 16);
  int64_t grid = n_rows < 16 * PM_NUM_SMS ? n_rows : 16 * PM_NUM_SMS;
  if (v4) gather_rows_kernel<4><<<(int)grid, 256, 0, st>>>(src, lds, idx, out, ldo, n_rows, width);
  else gather_rows_kernel<1><<<(int)grid, 256, 0, st>>>(src, lds, idx, out, ldo, n_rows, width);
  PM_CHECK_LAUNCH("rows");
  return PM_OK;
}

}  // namespace

extern "C" {

int pm_gather_rows(const float* src, int64_t lds, const int64_t* idx, float* out, int64_t ldo, int64_t n_rows,
                   int width, pm_stream_t s) {
  PM_REQUIRE(src && idx && out && n_rows > 0 && width > 0, PM_ERR_ARG, "pm_gather_rows: bad args");
  return launch_rows(src, lds, idx, out, ldo, n_rows, width, pm_st(s));
}

int pm_copy_rows(const float* src, int64_t lds, float* dst, int64_t ldd, int64_t n_rows, int width, pm_stream_t s) {
  PM_REQUIRE(src && dst && n_rows > 0 && width > 0, PM_ERR_ARG, "pm_copy_rows: bad args");
  return launch_rows(src, lds, nullptr, dst, ldd, n_rows, width, pm_st(s));
}

}  // extern "C"
