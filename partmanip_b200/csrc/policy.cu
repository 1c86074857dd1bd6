// K4 — Gaussian policy head and PPO losses.
// reference: algorithms/algo_utils/actor_critic.py:36-100, algorithms/ppo.py:326-374.
// Everything here is O(B*A) with A <= 32: latency-bound, so the point is one launch instead of the
// dozens MultivariateNormal issues, and device-side bookkeeping instead of .item() syncs.
#include "common.cuh"

namespace {

constexpr int MAXA = 32;
constexpr float HALF_LOG_2PI = 0.91893853320467274178f;

// ---------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
}

__global__ void randn_kernel(float* __restrict__ out, int64_t n, uint64_t seed, uint64_t offset) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // one Philox block -> 4 normals
  if (q * 4 >= n) return;
  const uint64_t ctr = offset + (uint64_t)q;
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
  uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
  for (int r = 0; r < 10; ++r) philox_round(c, k);
  float z[4];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const float u1 = ((float)c[2 * p] + 1.0f) * 2.3283064365386963e-10f;        // (0,1]
    const float u2 = (float)c[2 * p + 1] * 2.3283064365386963e-10f;             // [0,1)
    const float r = sqrtf(-2.f * logf(u1));
    float sn, cs;
    sincospif(2.f * u2, &sn, &cs);
    z[2 * p] = r * cs;
    z[2 * p + 1] = r * sn;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (q * 4 + j < n) out[q * 4 + j] = z[j];
}

// ---------------------------------------------------------------- sampling (actor_critic.py:36-47)
__global__ void policy_sample_kernel(const float* __restrict__ mu, const float* __restrict__ log_std,
                                     const float* __restrict__ eps, int E, int A, float max_action, int squash,
                                     float* __restrict__ actions, float* __restrict__ logp,
                                     float* __restrict__ sigma) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  float m = 0.f, hld = 0.f;
  for (int a = 0; a < A; ++a) {
    const float ls = log_std[a];
    const float el = expf(ls);
    const float sd = el * el;                       // scale_tril diagonal = exp(ls)*exp(ls)  (Q1)
    const float ep = eps[(int64_t)e * A + a];
    const float mean = mu[(int64_t)e * A + a];
    const float raw = mean + sd * ep;               // loc + scale_tril @ eps
    const float z = (raw - mean) / sd;
    m += z * z;
    hld += logf(sd);
    if (actions) actions[(int64_t)e * A + a] = squash ? tanhf(raw) * max_action : raw;
    if (sigma) sigma[(int64_t)e * A + a] = ls;      // "sigma" is log_std repeated (Q2)
  }
  if (logp) logp[e] = -0.5f * ((float)A * (2.f * HALF_LOG_2PI) + m) - hld;
}

__global__ void action_activation_kernel(const float* __restrict__ mu, float* __restrict__ out, int64_t n,
                                         float max_action, int squash) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = squash ? tanhf(mu[i]) * max_action : mu[i];
}

// log-prob / entropy of stored (squashed) actions (actor_critic.py:71-82), no gradient
__global__ void policy_logprob_kernel(const float* __restrict__ mu, int64_t ldmu, const float* __restrict__ log_std,
                                      const float* __restrict__ actions, int B, int A, float max_action, int squash,
                                      float* __restrict__ logp, float* __restrict__ entropy) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float m = 0.f, hld = 0.f;
  for (int a = 0; a < A; ++a) {
    const float el = expf(log_std[a]);
    const float sd = el * el;
    float raw = actions[(int64_t)b * A + a];
    if (squash) raw = atanhf(fminf(fmaxf(raw / max_action, -0.99999f), 0.99999f));
    const float z = (raw - mu[(int64_t)b * ldmu + a]) / sd;
    m += z * z;
    hld += logf(sd);
  }
  if (logp) logp[b] = -0.5f * ((float)A * (2.f * HALF_LOG_2PI) + m) - hld;
  if (entropy) entropy[b] = 0.5f * (float)A * (1.f + 2.f * HALF_LOG_2PI) + hld;
}

// ---------------------------------------------------------------- actor loss (ppo.py:326-344)
__global__ void __launch_bounds__(256)
actor_loss_kernel(const float* __restrict__ mu, int64_t ldmu, const float* __restrict__ log_std,
                  const float* __restrict__ actions, const float* __restrict__ logp_old,
                  const float* __restrict__ mu_old, const float* __restrict__ sigma_old,
                  const float* __restrict__ adv, const float* __restrict__ adv_stats, int B, int A,
                  float inv_batch, float clip_lo, float clip_hi, float max_action, int squash,
                  float* __restrict__ dmu, int64_t lddmu, float* __restrict__ logp_out,
                  float* __restrict__ partial /* [gridDim.x][2 + A] */) {
  __shared__ float sred[32];
  __shared__ float s_ls[MAXA], s_sd[MAXA], s_logsd[MAXA], s_e2[MAXA];
  if (threadIdx.x < A) {
    const float ls = log_std[threadIdx.x];
    const float el = expf(ls);
    s_ls[threadIdx.x] = ls;
    s_sd[threadIdx.x] = el * el;
    s_logsd[threadIdx.x] = logf(el * el);
    s_e2[threadIdx.x] = el * el;                    // torch.square(sigma.exp()) in the KL uses exp(ls)^2 too
  }
  __syncthreads();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  float surr = 0.f, kl = 0.f, g = 0.f;
  float z[MAXA];
  if (b < B) {
    float m = 0.f, hld = 0.f;
    for (int a = 0; a < A; ++a) {
      const float act = actions[(int64_t)b * A + a];
      float raw = act;
      if (squash) {                                 // atanh(clamp(a/max, +-(1-1e-5)))  (Q4)
        float t = act / max_action;
        t = fminf(fmaxf(t, -0.99999f), 0.99999f);
        raw = atanhf(t);
      }
      const float mean = mu[(int64_t)b * ldmu + a];
      const float zz = (raw - mean) / s_sd[a];
      z[a] = zz;
      m += zz * zz;
      hld += s_logsd[a];
      const float lso = sigma_old[(int64_t)b * A + a];
      const float eo = expf(lso);
      const float dm = mu_old[(int64_t)b * A + a] - mean;
      kl += s_ls[a] - lso + (eo * eo + dm * dm) / (2.0f * s_e2[a]) - 0.5f;
    }
    const float logp = -0.5f * ((float)A * (2.f * HALF_LOG_2PI) + m) - hld;
    if (logp_out) logp_out[b] = logp;
    float ad = adv[b];
    if (adv_stats) ad = (ad - adv_stats[0]) / adv_stats[1];
    const float ratio = expf(logp - logp_old[b]);
    const float rc = fminf(fmaxf(ratio, clip_lo), clip_hi);
    const float s1 = -ad * ratio, s2 = -ad * rc;
    surr = fmaxf(s1, s2);
    // d max(s1,s2)/d logp: torch.max splits ties 0.5/0.5; inside the clip range both paths carry -adv*ratio
    const bool in_range = (ratio >= clip_lo) && (ratio <= clip_hi);
    float w;
    if (in_range) w = 1.f;
    else w = (s1 > s2) ? 1.f : ((s1 == s2) ? 0.5f : 0.f);
    g = w * (-ad * ratio) * inv_batch;
    for (int a = 0; a < A; ++a) dmu[(int64_t)b * lddmu + a] = g * z[a] / s_sd[a];
  }
  // block partials: [0]=sum surrogate, [1]=sum kl, [2+a]=sum_b g*(2 z_a^2 - 2)
  float* out = partial + (int64_t)blockIdx.x * (2 + A);
  float t = pm_block_sum(surr, sred);
  if (threadIdx.x == 0) out[0] = t;
  t = pm_block_sum(kl, sred);
  if (threadIdx.x == 0) out[1] = t;
  for (int a = 0; a < A; ++a) {
    const float v = (b < B) ? g * (2.f * z[a] * z[a] - 2.f) : 0.f;
    t = pm_block_sum(v, sred);
    if (threadIdx.x == 0) out[2 + a] = t;
  }
}

__global__ void actor_loss_reduce(const float* __restrict__ partial, int nblk, int A, float* __restrict__ stats,
                                  float* __restrict__ dlog_std) {
  const int j = threadIdx.x;
  if (j >= 2 + A) return;
  float t = 0.f;
  for (int i = 0; i < nblk; ++i) t += partial[(int64_t)i * (2 + A) + j];
  if (j < 2) stats[j] = t;
  else dlog_std[j - 2] = t;
}

__global__ void actor_finalize_kernel(const float* __restrict__ stats, float inv_batch, float desired_kl,
                                      float* __restrict__ acc, int32_t* __restrict__ skip_flag) {
  const float kl_mean = stats[1] * inv_batch;
  if (kl_mean > acc[3]) acc[3] = kl_mean;           // kl_max tracks skipped minibatches too (ppo.py:335-336)
  const int skip = kl_mean > desired_kl;            // ppo.py:337-338
  *skip_flag = skip;
  if (!skip) {
    acc[0] += stats[0] * inv_batch;                 // mean_surrogate_loss += loss.item()
    acc[1] += kl_mean;
    acc[2] += 1.f;
  }
}

// ---------------------------------------------------------------- value loss (ppo.py:368-374)
__global__ void __launch_bounds__(256)
value_loss_kernel(const float* __restrict__ v, int64_t ldv, const float* __restrict__ ret,
                  const float* __restrict__ old_v, const float* __restrict__ clip_delta, int B, float inv_batch,
                  float* __restrict__ dv, int64_t lddv, float* __restrict__ partial) {
  __shared__ float sred[32];
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  float l = 0.f;
  if (b < B) {
    float target = ret[b];
    if (clip_delta) {
      const float d = *clip_delta, ov = old_v[b];
      target = ov + fminf(fmaxf(target - ov, -d), d);
    }
    const float diff = v[(int64_t)b * ldv] - target;
    l = diff * diff;
    dv[(int64_t)b * lddv] = 2.f * diff * inv_batch;
  }
  const float t = pm_block_sum(l, sred);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// DAgger behaviour-cloning loss (dagger.py:312-314): mean((tea_act - act(mu))^2) and d/dmu
__global__ void __launch_bounds__(256)
dagger_loss_kernel(const float* __restrict__ mu, int64_t ldmu, const float* __restrict__ tea_act, int64_t n, int A,
                   float max_action, int squash, float inv_count, float* __restrict__ dmu, int64_t lddmu,
                   float* __restrict__ partial) {
  __shared__ float sred[32];
  float l = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / A;
    const int a = (int)(i - b * A);
    const float m = mu[b * ldmu + a];
    const float t = squash ? tanhf(m) : m;
    const float stu = squash ? t * max_action : m;
    const float diff = tea_act[i] - stu;
    l += diff * diff;
    const float dact = squash ? max_action * (1.f - t * t) : 1.f;
    dmu[b * lddmu + a] = -2.f * diff * inv_count * dact;
  }
  l = pm_block_sum(l, sred);
  if (threadIdx.x == 0) partial[blockIdx.x] = l;
}

__global__ void sum_partials_kernel(const float* __restrict__ partial, int n, float scale, float* __restrict__ out) {
  __shared__ float sred[32];
  float t = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) t += partial[i];   // fixed assignment -> deterministic
  t = pm_block_sum(t, sred);
  if (threadIdx.x == 0) out[0] = t * scale;
}

__global__ void __launch_bounds__(256)
abs_partial_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ partial) {
  __shared__ float sred[32];
  float t = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    t += fabsf(x[i]);
  t = pm_block_sum(t, sred);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

__global__ void accumulate_kernel(const float* stats, float scale, float* acc, int idx) { acc[idx] += stats[0] * scale; }

}  // namespace

extern "C" {

int pm_randn(float* out, int64_t n, uint64_t seed, uint64_t offset, pm_stream_t s) {
  PM_REQUIRE(out && n >= 0, PM_ERR_ARG, "pm_randn: bad args");
  if (n == 0) return PM_OK;
  const int64_t q = (n + 3) / 4;
  randn_kernel<<<pm_cdiv(q, 256), 256, 0, pm_st(s)>>>(out, n, seed, offset);
  PM_CHECK_LAUNCH("pm_randn");
  return PM_OK;
}

int pm_policy_sample(const float* mu, const float* log_std, const float* eps, int E, int A, float max_action,
                     int squash, float* actions, float* logp, float* sigma, pm_stream_t s) {
  PM_REQUIRE(mu && log_std && eps, PM_ERR_ARG, "pm_policy_sample: null pointer");
  PM_REQUIRE(E > 0 && A > 0 && A <= MAXA, PM_ERR_SHAPE, "pm_policy_sample: E=%d A=%d (A<=%d)", E, A, MAXA);
  PM_REQUIRE(max_action > 0, PM_ERR_ARG, "pm_policy_sample: max_action must be > 0");
  policy_sample_kernel<<<pm_cdiv(E, 128), 128, 0, pm_st(s)>>>(mu, log_std, eps, E, A, max_action, squash, actions,
                                                                logp, sigma);
  PM_CHECK_LAUNCH("pm_policy_sample");
  return PM_OK;
}

int pm_action_activation(const float* mu, float* out, int64_t n, float max_action, int squash, pm_stream_t s) {
  PM_REQUIRE(mu && out && n > 0, PM_ERR_ARG, "pm_action_activation: bad args");
  action_activation_kernel<<<pm_cdiv(n, 256), 256, 0, pm_st(s)>>>(mu, out, n, max_action, squash);
  PM_CHECK_LAUNCH("pm_action_activation");
  return PM_OK;
}

int pm_policy_logprob(const float* mu, int64_t ldmu, const float* log_std, const float* actions, int B, int A,
                      float max_action, int squash, float* logp, float* entropy, pm_stream_t s) {
  PM_REQUIRE(mu && log_std && actions, PM_ERR_ARG, "pm_policy_logprob: null pointer");
  PM_REQUIRE(B > 0 && A > 0 && A <= MAXA, PM_ERR_SHAPE, "pm_policy_logprob: B=%d A=%d", B, A);
  policy_logprob_kernel<<<pm_cdiv(B, 128), 128, 0, pm_st(s)>>>(mu, ldmu, log_std, actions, B, A, max_action, squash,
                                                                logp, entropy);
  PM_CHECK_LAUNCH("pm_policy_logprob");
  return PM_OK;
}

size_t pm_ppo_actor_loss_ws_bytes(int B, int A) { return (size_t)pm_cdiv(B, 256) * (2 + A) * sizeof(float) + 256; }

int pm_ppo_actor_loss(const float* mu, int64_t ldmu, const float* log_std, const float* actions,
                      const float* logp_old, const float* mu_old, const float* sigma_old, const float* adv,
                      const float* adv_stats, int B, int A, float inv_batch, float eps_clip, float max_action,
                      int squash, float* stats, float* dmu, int64_t lddmu, float* dlog_std, float* logp_out,
                      void* ws, pm_stream_t s) {
  PM_REQUIRE(mu && log_std && actions && logp_old && mu_old && sigma_old && adv && stats && dmu && dlog_std && ws,
             PM_ERR_ARG, "pm_ppo_actor_loss: null pointer");
  PM_REQUIRE(B > 0 && A > 0 && A <= MAXA, PM_ERR_SHAPE, "pm_ppo_actor_loss: B=%d A=%d (A<=%d)", B, A, MAXA);
  const int nblk = pm_cdiv(B, 256);
  float* partial = reinterpret_cast<float*>(ws);
  const float lo = (float)(1.0 - (double)eps_clip), hi = (float)(1.0 + (double)eps_clip);
  actor_loss_kernel<<<nblk, 256, 0, pm_st(s)>>>(mu, ldmu, log_std, actions, logp_old, mu_old, sigma_old, adv,
                                                 adv_stats, B, A, inv_batch, lo, hi, max_action, squash, dmu, lddmu,
                                                 logp_out, partial);
  actor_loss_reduce<<<1, 64, 0, pm_st(s)>>>(partial, nblk, A, stats, dlog_std);
  PM_CHECK_LAUNCH("pm_ppo_actor_loss");
  return PM_OK;
}

int pm_ppo_actor_finalize(const float* stats, float inv_batch, float desired_kl, float* acc, int32_t* skip_flag,
                          pm_stream_t s) {
  PM_REQUIRE(stats && acc && skip_flag, PM_ERR_ARG, "pm_ppo_actor_finalize: null pointer");
  actor_finalize_kernel<<<1, 1, 0, pm_st(s)>>>(stats, inv_batch, desired_kl, acc, skip_flag);
  PM_CHECK_LAUNCH("pm_ppo_actor_finalize");
  return PM_OK;
}

int pm_value_loss(const float* v, int64_t ldv, const float* returns, const float* old_values,
                  const float* clip_delta, int B, float inv_batch, float* stats, float* dv, int64_t lddv, void* ws,
                  pm_stream_t s) {
  PM_REQUIRE(v && returns && stats && dv && ws && B > 0, PM_ERR_ARG, "pm_value_loss: bad args");
  PM_REQUIRE(!clip_delta || old_values, PM_ERR_ARG, "pm_value_loss: old_values required when clipped");
  const int nblk = pm_cdiv(B, 256);
  float* partial = reinterpret_cast<float*>(ws);
  value_loss_kernel<<<nblk, 256, 0, pm_st(s)>>>(v, ldv, returns, old_values, clip_delta, B, inv_batch, dv, lddv,
                                                 partial);
  sum_partials_kernel<<<1, 256, 0, pm_st(s)>>>(partial, nblk, 1.f, stats);
  PM_CHECK_LAUNCH("pm_value_loss");
  return PM_OK;
}

int pm_dagger_loss(const float* mu, int64_t ldmu, const float* tea_act, int B, int A, float max_action, int squash,
                   float inv_count, float* stats, float* dmu, int64_t lddmu, void* ws, pm_stream_t s) {
  PM_REQUIRE(mu && tea_act && stats && dmu && ws && B > 0 && A > 0, PM_ERR_ARG, "pm_dagger_loss: bad args");
  const int64_t n = (int64_t)B * A;
  int nblk = pm_cdiv(n, 256);
  if (nblk > 256) nblk = 256;
  float* partial = reinterpret_cast<float*>(ws);
  dagger_loss_kernel<<<nblk, 256, 0, pm_st(s)>>>(mu, ldmu, tea_act, n, A, max_action, squash, inv_count, dmu, lddmu,
                                                  partial);
  sum_partials_kernel<<<1, 256, 0, pm_st(s)>>>(partial, nblk, inv_count, stats);
  PM_CHECK_LAUNCH("pm_dagger_loss");
  return PM_OK;
}

int pm_abs_sum(const float* x, int64_t n, float scale, float* out, void* ws, pm_stream_t s) {
  PM_REQUIRE(x && out && ws && n > 0, PM_ERR_ARG, "pm_abs_sum: bad args");
  int nblk = pm_cdiv(n, 256 * 8);
  if (nblk > 256) nblk = 256;
  float* partial = reinterpret_cast<float*>(ws);
  abs_partial_kernel<<<nblk, 256, 0, pm_st(s)>>>(x, n, partial);
  sum_partials_kernel<<<1, 256, 0, pm_st(s)>>>(partial, nblk, scale, out);
  PM_CHECK_LAUNCH("pm_abs_sum");
  return PM_OK;
}

int pm_accumulate(const float* stats, float scale, float* acc, int idx, pm_stream_t s) {
  PM_REQUIRE(stats && acc && idx >= 0, PM_ERR_ARG, "pm_accumulate: bad args");
  accumulate_kernel<<<1, 1, 0, pm_st(s)>>>(stats, scale, acc, idx);
  PM_CHECK_LAUNCH("pm_accumulate");
  return PM_OK;
}

}  // extern "C"
