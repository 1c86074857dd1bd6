// K7f — ONE kernel per optimiser step for: cross-GPU gradient all-reduce (one-shot pull over NVLink peer memory) -> KL-skip
// decision -> global-L2-norm clip -> Adam.  Replaces, per minibatch step, NCCL all-reduce + pm_ppo_actor_finalize + the three
// launches of pm_adam_step.  reference: nn.utils.clip_grad_norm_ + torch.optim.Adam.step as called from
// algorithms/ppo.py:337-338 (KL skip), 351-353 (actor), 381-382 (critic); the reference has no distributed code — the
// all-reduce is the build's env-sharded data parallelism.
//
// Every rank's gradient buffer lives in SYMMETRIC memory (torch.distributed._symmetric_memory: the same allocation mapped into
// every rank's address space), so a rank reads its peers' gradients with plain global loads that travel over NVLink:
//   1. ready barrier : block 0 stores the launch's sequence number into flag R[rank] of every peer (st.release.sys); every CTA
//                      polls its OWN rank's R[0..world) (local memory) until all peers' gradients of this step are published;
//   2. pull + reduce : each CTA owns a slice of the flat buffer, sums it over ranks in rank order 0..world-1 (identical order on
//                      every rank => bit-identical replicas), keeps the reduced slice in a local buffer and its sum of squares;
//   3. grid barrier  : (atomic counter) — then block 0 publishes "done reading" flags D[rank] to every peer;
//   4. every CTA forms the global norm from the per-CTA partials in fixed order, the KL-skip predicate from the reduced
//      [sum surrogate, sum KL] tail, the bias corrections, and applies clip + Adam to its slice;
//   5. block 0 waits for all peers' D flags before the kernel ends: kernel completion implies that nobody still reads this
//      rank's gradient buffer, so the next backward may overwrite it (stream order does the rest).
// world == 1 skips 1, 3b and 5 (one launch instead of four).  All waits are bounded: a peer that never arrives surfaces as an
// error word (and the sticky flag), not as a hung GPU.
#include "common.cuh"

int32_t* pm_tc_sticky_word();
int pm_sm_count();

namespace {

constexpr int FS_THREADS = 512;
constexpr int FS_MAX_WORLD = 8;
constexpr int FLAG_DONE = 32;             // uint32 index of D[0] inside a rank's flag buffer (R[0..8) at 0)

struct FusedStepP {
  float* params; float* exp_avg; float* exp_avg_sq;
  int64_t n, n_clip;
  int n_tail;
  float max_norm, beta1, beta2, eps;
  float* opt_state;
  const float* grad_local;
  const float* const* grad_peers;         // device array [world] (entry `rank` == grad_local) or null
  uint32_t* const* flag_peers;            // device array [world] of flag buffers or null
  volatile uint32_t* flags_local;
  int rank, world;
  float* gred;                            // [n + n_tail] reduced gradient (local)
  int finalize;                           // 1: actor (KL-skip from the tail)
  float inv_batch, desired_kl;
  float* acc; int32_t* skip_flag;
  uint32_t* sync;                         // [0] grid-barrier counter, [1] launch sequence number
  double* partial;                        // [gridDim.x]
  int32_t* err_last; int32_t* err_sticky;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const volatile uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// peer gradients: system-scope loads that bypass L1 (the line may live in another GPU's memory, reached over NVLink)
__device__ __forceinline__ float4 ld_peer4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer1(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void report(const FusedStepP& p, int code) {
  if (p.err_last) atomicExch(p.err_last, code);
  if (p.err_sticky) atomicCAS(p.err_sticky, 0, code);
}
// seq numbers wrap after 2^32 launches: compare as signed distance
__device__ __forceinline__ bool reached(uint32_t v, uint32_t s) { return (int32_t)(v - s) >= 0; }

__global__ void __launch_bounds__(FS_THREADS)
fused_step_kernel(const FusedStepP p) {
  __shared__ double smd[32];
  __shared__ float s_bc[8];
  __shared__ int s_ok;
  const int tid = threadIdx.x, G = gridDim.x;
  const uint32_t s = ld_acquire_gpu(p.sync + 1) + 1;              // this launch's sequence number (block 0 stores it at the very end)
  const float step_old = p.opt_state[0], lr = p.opt_state[1];
  const int64_t total = p.n + p.n_tail;
  if (tid == 0) s_ok = 1;
  __syncthreads();
  // ---- 1. ready barrier across ranks
  if (p.world > 1) {
    if (blockIdx.x == 0 && tid < p.world) {
      __threadfence_system();
      st_release_sys(p.flag_peers[tid] + p.rank, s);
    }
    if (tid < p.world) {
      bool ok = false;
      for (uint32_t spin = 0; spin < (1u << 26); ++spin)
        if (reached(ld_acquire_sys(p.flags_local + tid), s)) { ok = true; break; }
      if (!ok) { report(p, 701); s_ok = 0; }
    }
    __syncthreads();
  }
  // ---- 2. pull + reduce this CTA's slice (multiples of 4 floats; the buffers are 16-byte aligned)
  const int64_t per = ((total + G - 1) / G + 3) / 4 * 4;
  const int64_t i0 = min(total, (int64_t)blockIdx.x * per), i1 = min(total, i0 + per);
  double sq = 0.0;
  if (s_ok) {
    // peer pointers into registers once (the table itself sits in device memory)
    const float* peer[FS_MAX_WORLD];
#pragma unroll
    for (int r = 0; r < FS_MAX_WORLD; ++r) peer[r] = (p.world > 1 && r < p.world) ? p.grad_peers[r] : p.grad_local;
    for (int64_t i = i0 + 4 * tid; i < i1; i += 4 * FS_THREADS) {
      float4 g;
      if (i + 4 <= i1) {
        if (p.world > 1) {
          // all peers' loads of this position are issued before the first add (NVLink round trip ~2 us: latency, not bandwidth,
          // bounds the pull), then summed in rank order
          float4 h[FS_MAX_WORLD];
#pragma unroll
          for (int r = 0; r < FS_MAX_WORLD; ++r)
            if (r < p.world) h[r] = ld_peer4(peer[r] + i);
          g = h[0];
#pragma unroll
          for (int r = 1; r < FS_MAX_WORLD; ++r)
            if (r < p.world) { g.x += h[r].x; g.y += h[r].y; g.z += h[r].z; g.w += h[r].w; }
        } else {
          g = *reinterpret_cast<const float4*>(p.grad_local + i);
        }
        *reinterpret_cast<float4*>(p.gred + i) = g;
      } else {
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        for (int e = 0; e < 4 && i + e < i1; ++e) {
          float a = 0.f;
          if (p.world > 1) { for (int r = 0; r < p.world; ++r) a += ld_peer1(p.grad_peers[r] + i + e); } else a = p.grad_local[i + e];
          p.gred[i + e] = a;
          t[e] = a;
        }
        g = make_float4(t[0], t[1], t[2], t[3]);
      }
      if (i < p.n_clip) sq += (double)g.x * g.x;
      if (i + 1 < p.n_clip) sq += (double)g.y * g.y;
      if (i + 2 < p.n_clip) sq += (double)g.z * g.z;
      if (i + 3 < p.n_clip) sq += (double)g.w * g.w;
    }
  }
  sq = pm_block_sum_d(sq, smd);
  if (tid == 0) p.partial[blockIdx.x] = sq;
  // ---- 3. grid barrier (all CTAs are co-resident: grid <= number of SMs, tiny footprint)
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    atomicAdd(p.sync, 1u);
    const uint32_t target = (uint32_t)G * s;
    bool ok = false;
    for (uint32_t spin = 0; spin < (1u << 26); ++spin)
      if (reached(ld_acquire_gpu(p.sync), target)) { ok = true; break; }
    if (!ok) { report(p, 702); s_ok = 0; }
  }
  __syncthreads();
  if (p.world > 1 && blockIdx.x == 0 && tid < p.world) {
    __threadfence_system();
    st_release_sys(p.flag_peers[tid] + FLAG_DONE + p.rank, s);      // this rank no longer reads peer `tid`'s gradients
  }
  // ---- 4. scalars (every CTA computes the same values from the same inputs, in the same order)
  if (tid == 0) {
    double t = 0.0;
    for (int j = 0; j < G; ++j) t += p.partial[j];
    const float total_norm = (float)sqrt(t);
    float coef = 1.f;
    if (p.max_norm > 0.f) coef = fminf(p.max_norm / (total_norm + 1e-6f), 1.0f);
    int skip = 0;
    float kl_mean = 0.f, sur = 0.f;
    if (p.finalize) {
      sur = p.gred[p.n] * p.inv_batch;
      kl_mean = p.gred[p.n + 1] * p.inv_batch;
      skip = kl_mean > p.desired_kl;                                   // ppo.py:337-338
    }
    if (!s_ok) skip = 1;                                               // a failed barrier must not step the weights
    const float step = skip ? step_old : step_old + 1.f;
    const double bc1 = 1.0 - pow((double)p.beta1, (double)step);
    const double bc2 = 1.0 - pow((double)p.beta2, (double)step);
    s_bc[0] = coef;
    s_bc[1] = (float)((double)lr / bc1);                               // step_size
    s_bc[2] = (float)sqrt(bc2);
    s_bc[3] = (float)skip;
    if (blockIdx.x == 0) {
      if (p.finalize) {
        if (kl_mean > p.acc[3]) p.acc[3] = kl_mean;                    // kl_max tracks skipped minibatches too (ppo.py:335-336)
        *p.skip_flag = skip;
        if (!skip) { p.acc[0] += sur; p.acc[1] += kl_mean; p.acc[2] += 1.f; }
      }
      p.opt_state[0] = step;
      p.opt_state[2] = total_norm; p.opt_state[3] = coef; p.opt_state[4] = s_bc[1]; p.opt_state[5] = s_bc[2]; p.opt_state[6] = (float)skip;
    }
  }
  __syncthreads();
  const float coef = s_bc[0], step_size = s_bc[1], bc2s = s_bc[2];
  if (s_bc[3] == 0.f) {
    const int64_t e1 = min(i1, p.n);
    for (int64_t i = i0 + tid; i < e1; i += FS_THREADS) {
      float gi = p.gred[i];
      if (i < p.n_clip) gi *= coef;
      const float mi = p.exp_avg[i] + (gi - p.exp_avg[i]) * (1.f - p.beta1);
      const float vi = p.exp_avg_sq[i] * p.beta2 + (1.f - p.beta2) * gi * gi;
      const float denom = sqrtf(vi) / bc2s + p.eps;
      p.params[i] = p.params[i] - step_size * (mi / denom);
      p.exp_avg[i] = mi;
      p.exp_avg_sq[i] = vi;
    }
  }
  // ---- 5. nobody reads this rank's gradients any more once all peers' D flags arrived
  if (blockIdx.x == 0) {
    if (p.world > 1 && tid < p.world) {
      bool ok = false;
      for (uint32_t spin = 0; spin < (1u << 26); ++spin)
        if (reached(ld_acquire_sys(p.flags_local + FLAG_DONE + tid), s)) { ok = true; break; }
      if (!ok) report(p, 703);
    }
    __syncthreads();
    if (tid == 0) { __threadfence(); atomicExch(p.sync + 1, s); }
  }
}

inline int fs_grid(int64_t total) {
  int g = (int)((total + 4 * FS_THREADS - 1) / (4 * FS_THREADS));
  if (g > PM_NUM_SMS) g = PM_NUM_SMS;
  const int sms = pm_sm_count();                 // the grid barrier needs every CTA resident at once
  if (g > sms) g = sms;
  if (g < 1) g = 1;
  return g;
}

}  // namespace

extern "C" {

// workspace: [sync (256 B) | err (256 B) | partial (148 doubles) | reduced gradient (n + n_tail floats)]
size_t pm_fused_step_ws_bytes(int64_t n, int n_tail) {
  return 512 + pm_align_up(PM_NUM_SMS * sizeof(double), 256) + pm_align_up((size_t)(n + n_tail) * sizeof(float), 256);
}

int pm_fused_step(float* params, float* exp_avg, float* exp_avg_sq, int64_t n, int64_t n_clip, int n_tail, float max_norm, float beta1,
                  float beta2, float eps, float* opt_state, const float* grad_local, const float* const* grad_peers_dev,
                  uint32_t* const* flag_peers_dev, uint32_t* flags_local, int rank, int world, int finalize, float inv_batch,
                  float desired_kl, float* acc, int32_t* skip_flag, void* ws, pm_stream_t s) {
  PM_REQUIRE(params && exp_avg && exp_avg_sq && opt_state && grad_local && ws, PM_ERR_ARG, "pm_fused_step: null pointer");
  PM_REQUIRE(n > 0 && n_clip >= 0 && n_clip <= n && n_tail >= 0 && n_tail <= 64, PM_ERR_SHAPE, "pm_fused_step: n=%lld n_clip=%lld n_tail=%d",
             (long long)n, (long long)n_clip, n_tail);
  PM_REQUIRE(world >= 1 && world <= FS_MAX_WORLD && rank >= 0 && rank < world, PM_ERR_ARG, "pm_fused_step: rank %d of %d", rank, world);
  PM_REQUIRE(world == 1 || (grad_peers_dev && flag_peers_dev && flags_local), PM_ERR_ARG, "pm_fused_step: peer tables required for world > 1");
  PM_REQUIRE(!finalize || (acc && skip_flag && n_tail >= 2), PM_ERR_ARG, "pm_fused_step: finalize needs acc, skip_flag and a 2-float tail");
  PM_REQUIRE(pm_aligned(grad_local, 16) && pm_aligned(ws, 256), PM_ERR_ALIGN, "pm_fused_step: gradient buffer / workspace alignment");
  uint8_t* w = reinterpret_cast<uint8_t*>(ws);
  FusedStepP p{};
  p.params = params; p.exp_avg = exp_avg; p.exp_avg_sq = exp_avg_sq; p.n = n; p.n_clip = n_clip; p.n_tail = n_tail;
  p.max_norm = max_norm; p.beta1 = beta1; p.beta2 = beta2; p.eps = eps; p.opt_state = opt_state;
  p.grad_local = grad_local; p.grad_peers = world > 1 ? grad_peers_dev : nullptr; p.flag_peers = world > 1 ? flag_peers_dev : nullptr;
  p.flags_local = flags_local; p.rank = rank; p.world = world;
  p.finalize = finalize; p.inv_batch = inv_batch; p.desired_kl = desired_kl; p.acc = acc; p.skip_flag = skip_flag;
  p.sync = reinterpret_cast<uint32_t*>(w);
  p.err_last = reinterpret_cast<int32_t*>(w + 256);
  p.err_sticky = pm_tc_sticky_word();
  p.partial = reinterpret_cast<double*>(w + 512);
  p.gred = reinterpret_cast<float*>(w + 512 + pm_align_up(PM_NUM_SMS * sizeof(double), 256));
  fused_step_kernel<<<fs_grid(n + n_tail), FS_THREADS, 0, pm_st(s)>>>(p);
  PM_CHECK_LAUNCH("pm_fused_step");
  return PM_OK;
}

}  // extern "C"
