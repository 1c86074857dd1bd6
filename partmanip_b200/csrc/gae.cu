// K5 — GAE reverse scan + returns + advantages (reference: algorithms/algo_utils/storage.py:96-114),
// and the unbiased-std normaliser used by whole_adv_norm / mini_adv_norm (storage.py:113-114, ppo.py:328-329).
// One thread per env walks t = T-1..0 (sequential in T, embarrassingly parallel in E); loads are
// coalesced across envs.  18 algorithmic bytes per (t,e) element; rounding follows the reference's
// op-by-op fp32 evaluation (no FMA contraction) so results are bit-exact.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
gae_kernel(const float* __restrict__ rew, const float* __restrict__ val, const uint8_t* __restrict__ done,
           const uint8_t* __restrict__ succ, const float* __restrict__ last, float* __restrict__ ret,
           float* __restrict__ adv, int T, int E, float gamma, float gamma_lam, int use_sv, float sv) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  float next_v = last[e];
  float a = 0.f;
  for (int t = T - 1; t >= 0; --t) {
    const int64_t i = (int64_t)t * E + e;
    const float v = val[i];
    const float nt = done[i] ? 0.f : 1.f;
    // delta = rewards + gamma * next_values - values
    const float delta = __fsub_rn(__fadd_rn(rew[i], __fmul_rn(gamma, next_v)), v);
    // advantage = not_terminal * (delta + gamma*lam * advantage)
    a = __fmul_rn(nt, __fadd_rn(delta, __fmul_rn(gamma_lam, a)));
    float r = __fadd_rn(a, v);
    if (use_sv) {
      const float s = succ[i] ? 1.f : 0.f;
      r = __fadd_rn(__fmul_rn(1.f - s, r), __fmul_rn(s, sv));
    }
    ret[i] = r;
    adv[i] = __fsub_rn(r, v);   // advantages = returns - values
    next_v = v;
  }
}

// single-CTA two-pass mean / unbiased std; n is small (T*E <= a few 100k)
__global__ void __launch_bounds__(1024)
normalize_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n, float* stats_out) {
  __shared__ double smd[32];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += (double)x[i];
  s = pm_block_sum_d(s, smd);
  const float mean = (float)(s / (double)n);
  double q = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = (double)x[i] - (double)mean;
    q += d * d;
  }
  q = pm_block_sum_d(q, smd);
  const float sd = (float)sqrt(q / (double)(n - 1));
  const float denom = sd + 1e-8f;
  if (stats_out && threadIdx.x == 0) {
    stats_out[0] = mean;
    stats_out[1] = denom;
  }
  if (out)
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) out[i] = __fdiv_rn(x[i] - mean, denom);
}

}  // namespace

extern "C" {

int pm_gae(const float* rewards, const float* values, const uint8_t* dones, const uint8_t* succs,
           const float* last_values, float* returns, float* advantages, int T, int E, float gamma, float gamma_lam,
           int use_succ_value, float succ_value, pm_stream_t s) {
  PM_REQUIRE(rewards && values && dones && last_values && returns && advantages, PM_ERR_ARG, "pm_gae: null pointer");
  PM_REQUIRE(T > 0 && E > 0, PM_ERR_SHAPE, "pm_gae: T=%d E=%d", T, E);
  PM_REQUIRE(!use_succ_value || succs, PM_ERR_ARG, "pm_gae: succs required with succ_value");
  gae_kernel<<<pm_cdiv(E, 256), 256, 0, pm_st(s)>>>(rewards, values, dones, succs, last_values, returns, advantages,
                                                      T, E, gamma, gamma_lam, use_succ_value, succ_value);
  PM_CHECK_LAUNCH("pm_gae");
  return PM_OK;
}

size_t pm_normalize_ws_bytes(int64_t n) { (void)n; return 256; }

int pm_normalize(const float* x, float* out, int64_t n, void* ws, pm_stream_t s) {
  PM_REQUIRE(x && n >= 2, PM_ERR_ARG, "pm_normalize: bad args");
  normalize_kernel<<<1, 1024, 0, pm_st(s)>>>(x, out, n, reinterpret_cast<float*>(ws));
  PM_CHECK_LAUNCH("pm_normalize");
  return PM_OK;
}

int pm_normalize_inplace(float* x, int64_t n, void* ws, pm_stream_t s) { return pm_normalize(x, x, n, ws, s); }

}  // extern "C"
