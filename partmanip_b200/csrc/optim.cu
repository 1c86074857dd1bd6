// K7 — global-L2 grad-norm clip + Adam on flat fp32 buffers.
// reference: nn.utils.clip_grad_norm_ + torch.optim.Adam defaults as called from algorithms/ppo.py:73-74,
// 351-353, 381-382 (third-party arithmetic restated; oracle/ppo_oracle.py:adam_step / clip_coef).
// 28 algorithmic bytes per parameter per step (read p,g,m,v; write p,m,v); HBM/L2-streaming.
// Step counter, lr and the KL-skip predicate live on the device so the step is CUDA-graph capturable.
#include "common.cuh"

namespace {

constexpr int SQ_THREADS = 256;
constexpr int SQ_MAX_BLOCKS = 592;

__global__ void __launch_bounds__(SQ_THREADS)
sumsq_partial_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ partial) {
  __shared__ double smd[32];
  double t = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = g[i];
    t += (double)v * (double)v;
  }
  t = pm_block_sum_d(t, smd);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

__global__ void adam_prepare_kernel(const double* __restrict__ partial, int nblk, float max_norm, float beta1,
                                    float beta2, float* __restrict__ opt_state, const int32_t* __restrict__ skip_flag) {
  __shared__ double smd[32];
  double t = 0.0;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) t += partial[i];
  t = pm_block_sum_d(t, smd);
  if (threadIdx.x != 0) return;
  const float total = (float)sqrt(t);
  float coef = 1.f;
  if (max_norm > 0.f) coef = fminf(max_norm / (total + 1e-6f), 1.0f);      // clip_grad_norm_: clamp(max/(total+1e-6), max=1)
  const int skip = skip_flag ? *skip_flag : 0;
  float step = opt_state[0];
  if (!skip) { step += 1.f; opt_state[0] = step; }
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  opt_state[2] = total;
  opt_state[3] = coef;
  opt_state[4] = (float)((double)opt_state[1] / bc1);                        // step_size = lr / bias_correction1
  opt_state[5] = (float)sqrt(bc2);                                           // bias_correction2_sqrt
  opt_state[6] = (float)skip;
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            int64_t n, int64_t n_clip, float beta1, float beta2, float eps, const float* __restrict__ opt_state) {
  if (opt_state[6] != 0.f) return;                                           // KL-skip: no optimizer step at all
  const float coef = opt_state[3], step_size = opt_state[4], bc2s = opt_state[5];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i];
    if (i < n_clip) gi *= coef;
    const float mi = m[i] + (gi - m[i]) * (1.f - beta1);                     // exp_avg.lerp_(grad, 1-beta1)
    const float vi = v[i] * beta2 + (1.f - beta2) * gi * gi;                 // exp_avg_sq.mul_(b2).addcmul_(g,g,1-b2)
    const float denom = sqrtf(vi) / bc2s + eps;
    p[i] = p[i] - step_size * (mi / denom);
    m[i] = mi;
    v[i] = vi;
  }
}

}  // namespace

extern "C" {

size_t pm_adam_ws_bytes(int64_t n) { (void)n; return SQ_MAX_BLOCKS * sizeof(double); }

int pm_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, int64_t n_clip,
                 float max_norm, float beta1, float beta2, float eps, float* opt_state, const int32_t* skip_flag,
                 void* ws, pm_stream_t s) {
  PM_REQUIRE(params && grads && exp_avg && exp_avg_sq && opt_state && ws, PM_ERR_ARG, "pm_adam_step: null pointer");
  PM_REQUIRE(n > 0 && n_clip >= 0 && n_clip <= n, PM_ERR_SHAPE, "pm_adam_step: n=%lld n_clip=%lld", (long long)n,
             (long long)n_clip);
  double* partial = reinterpret_cast<double*>(ws);
  int nblk = 1;
  if (max_norm > 0.f && n_clip > 0) {
    nblk = pm_cdiv(n_clip, SQ_THREADS * 4);
    if (nblk > SQ_MAX_BLOCKS) nblk = SQ_MAX_BLOCKS;
    sumsq_partial_kernel<<<nblk, SQ_THREADS, 0, pm_st(s)>>>(grads, n_clip, partial);
  } else {
    cudaMemsetAsync(partial, 0, sizeof(double), pm_st(s));
  }
  adam_prepare_kernel<<<1, 256, 0, pm_st(s)>>>(partial, nblk, max_norm, beta1, beta2, opt_state, skip_flag);
  int ablk = pm_cdiv(n, 256 * 2);
  if (ablk > 4 * PM_NUM_SMS) ablk = 4 * PM_NUM_SMS;
  adam_kernel<<<ablk, 256, 0, pm_st(s)>>>(params, grads, exp_avg, exp_avg_sq, n, n_clip, beta1, beta2, eps, opt_state);
  PM_CHECK_LAUNCH("pm_adam_step");
  return PM_OK;
}

}  // extern "C"
