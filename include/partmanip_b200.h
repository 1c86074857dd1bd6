/*
 * partmanip_b200.h — C-ABI of the B200-native PartManip PPO hot path (libpartmanip_b200.so).
 *
 * The reference (PKU-EPIC/PartManip) has no FFI: its hot path is Python calling PyTorch.  This
 * header is the seam the build introduces UNDER the reference's Python classes:
 * every entry point replaces a span of reference Python/PyTorch code, cited as file:line relative
 * to the reference root.  Conventions:
 *   - plain C: raw DEVICE pointers, sizes, strides (in elements), a cudaStream_t passed as void*;
 *     no torch types.  All tensors are fp32 row-major unless stated; "ld" = leading dimension
 *     (row stride, elements).  Bool tensors are uint8 (torch.bool storage).
 *   - every call is stream-ordered and asynchronous; none synchronises or allocates.  Workspaces
 *     are caller-owned; their sizes come from the matching pm_*_ws_bytes() query.
 *   - return value: PM_OK (0) or a negative pm_status.  pm_last_error() gives a message.
 *   - device-side scalars (optimizer step, skip flag, counts) live in caller-owned device memory
 *     so that whole minibatch steps can be captured in CUDA graphs without host round trips.
 */
#ifndef PARTMANIP_B200_H
#define PARTMANIP_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* pm_stream_t; /* cudaStream_t */

typedef enum {
  PM_OK = 0,
  PM_ERR_SHAPE = -1,
  PM_ERR_ARG = -2,
  PM_ERR_ALIGN = -3,
  PM_ERR_CUDA = -4,
  PM_ERR_UNSUPPORTED = -5
} pm_status;

/* algorithms/algo_utils/network.py:7-24 get_activation(); "crelu" maps to PM_ACT_RELU there. */
typedef enum {
  PM_ACT_NONE = 0,
  PM_ACT_TANH = 1,
  PM_ACT_RELU = 2,
  PM_ACT_ELU = 3,
  PM_ACT_SELU = 4,
  PM_ACT_LRELU = 5,
  PM_ACT_SIGMOID = 6
} pm_act;

/* arithmetic mode of the encoder's two GEMM layers */
typedef enum {
  PM_PREC_FP32 = 0,     /* the reference's precision (parity gate 1e-4) on the tensor cores: operands split into fp16 / bf16 terms,
                           several tcgen05 MMAs per product, fp32 accumulate in TMEM; shapes the tcgen05 kernels do not take
                           (N % 256 != 0, C > 4, max_mean pooling) run on the CUDA-core kernels */
  PM_PREC_BF16 = 1,     /* tcgen05 bf16 operands, fp32 accumulate in TMEM: parity gate 1e-2 */
  PM_PREC_FP32_FFMA = 2 /* fp32 on CUDA cores only (FFMA): parity gate 1e-4 */
} pm_precision;

const char* pm_last_error(void);
int pm_version(void);
/* 1 if the tcgen05 (sm_100a) encoder path was compiled in */
int pm_has_tcgen05(void);
/* The tcgen05 kernels bound every mbarrier wait; a wait that times out (protocol bug, or a pathologically slow device) records
 * its code in a per-device word that no launch clears.  Returns the first code since the last clear on the CURRENT device
 * (0 = none) and clears it when clear != 0.  Synchronises the device.  The algorithm classes call it once per iteration and
 * raise: features / gradients of a launch that reported an error are garbage (the reference has no equivalent: PyTorch
 * kernels cannot time out). */
int pm_tc_sticky_error(int clear);

/* ------------------------------------------------------------------------------------------
 * K6  running mean / std observation normaliser
 * replaces algorithms/algo_utils/RMS.py:10-18 (RunningMeanStd.update) and :40-45
 * (Normalization.__call__).  Split in four so a multi-GPU caller can all-reduce colsum / sqdev
 * between the calls; pm_rms_forward chains them for one GPU.
 * ------------------------------------------------------------------------------------------ */
size_t pm_colreduce_ws_bytes(int rows, int cols);
/* colsum[d] = sum_e x[e,d]                                   (x.mean(dim=0) numerator, RMS.py:14) */
int pm_rms_colsum(const float* x, int64_t ldx, int E, int D, float* colsum, void* ws, pm_stream_t s);
/* sqdev[d] = sum_e (x[e,d] - colsum[d]/count)^2               ((x-new_mean).pow(2).mean numerator, RMS.py:16)
 * count = number of rows behind colsum (E, or the global env count when colsum was all-reduced) */
int pm_rms_colsqdev(const float* x, int64_t ldx, int E, int D, const float* colsum, float count,
                    float* sqdev, void* ws, pm_stream_t s);
/* mean,S,std update with n = update counter AFTER the increment  (RMS.py:12-17) */
int pm_rms_update(float* mean, float* S, float* std, const float* colsum, const float* sqdev,
                  float count, int n, int D, pm_stream_t s);
/* out = (x - mean) / std, no epsilon                            (RMS.py:44) */
int pm_rms_normalize(const float* x, int64_t ldx, float* out, int64_t ldo, int E, int D,
                     const float* mean, const float* std, pm_stream_t s);
/* single-GPU chain; `scratch` holds 2*D floats + pm_colreduce_ws_bytes(E,D) */
size_t pm_rms_forward_ws_bytes(int E, int D);
int pm_rms_forward(const float* x, int64_t ldx, float* out, int64_t ldo, int E, int D, float* mean,
                   float* S, float* std, int n_after, int update, void* scratch, pm_stream_t s);

/* ------------------------------------------------------------------------------------------
 * K5  GAE / returns / advantages
 * replaces algorithms/algo_utils/storage.py:96-114 (RolloutStorage.compute_returns).
 * rewards, values, returns, advantages: (T,E); dones, succs: (T,E) uint8; last_values: (E).
 * use_succ_value=0 reproduces default_succ_value=None.  Bit-exact with the reference's op order.
 * ------------------------------------------------------------------------------------------ */
int pm_gae(const float* rewards, const float* values, const uint8_t* dones, const uint8_t* succs,
           const float* last_values, float* returns, float* advantages, int T, int E, float gamma,
           float gamma_lam, int use_succ_value, float succ_value, pm_stream_t s);
/* x <- (x - mean(x)) / (std_unbiased(x) + 1e-8) over n elements  (storage.py:113-114, ppo.py:328-329) */
size_t pm_normalize_ws_bytes(int64_t n);
int pm_normalize_inplace(float* x, int64_t n, void* ws, pm_stream_t s);
/* out <- normalised copy of x (mini_adv_norm works on the gathered minibatch copy) */
int pm_normalize(const float* x, float* out, int64_t n, void* ws, pm_stream_t s);

/* ------------------------------------------------------------------------------------------
 * K4  Gaussian policy head and PPO losses
 * ------------------------------------------------------------------------------------------ */
/* standard-normal draws, Philox4x32-10 + Box-Muller; counter = (offset + element index) */
int pm_randn(float* out, int64_t n, uint64_t seed, uint64_t offset, pm_stream_t s);

/* replaces actor_critic.py:36-47 (random_act_cri) after the actor forward:
 *   raw = mu + exp(log_std)^2 * eps ; actions = tanh(raw)*max_action (squash=1) or raw (squash=0);
 *   logp = MultivariateNormal(mu, scale_tril=diag(exp(log_std)^2)).log_prob(raw);
 *   sigma = log_std broadcast to (E,A)  (SURVEY Q1, Q2).  actions/logp/sigma may be NULL. */
int pm_policy_sample(const float* mu, const float* log_std, const float* eps, int E, int A,
                     float max_action, int squash, float* actions, float* logp, float* sigma,
                     pm_stream_t s);
/* replaces actor_critic.py:58-65,84-91 (act / act_cri post-processing): out = tanh(mu)*max_action */
int pm_action_activation(const float* mu, float* out, int64_t n, float max_action, int squash,
                         pm_stream_t s);

/* replaces actor_critic.py:71-82 (update_act_cri's distribution part, no gradient): log-prob of the stored
 * squashed actions (Q4) and the Gaussian entropy.  logp / entropy: (B), either may be NULL. */
int pm_policy_logprob(const float* mu, int64_t ldmu, const float* log_std, const float* actions, int B, int A,
                      float max_action, int squash, float* logp, float* entropy, pm_stream_t s);

/* replaces actor_critic.py:71-82 (log-prob of stored squashed actions, Q4) + ppo.py:326-344:
 *   raw   = atanh(clamp(actions/max_action, +-(1-1e-5)))             (actor_critic.py:93-95)
 *   logp  = log N(raw; mu, exp(2 log_std))
 *   kl    = sum_a(ls - ls_old + (exp(ls_old)^2 + (mu_old-mu)^2)/(2 exp(ls)^2) - 0.5)   (ppo.py:332-333)
 *   loss  = mean_b max(-adv*r, -adv*clamp(r,1-eps,1+eps)), r = exp(logp - logp_old)    (ppo.py:341-344)
 * writes stats[0] = sum_b surrogate_b, stats[1] = sum_b kl_b (LOCAL sums; caller all-reduces),
 * dmu[B,A] and dlog_std[A] = d(loss)/d(.) with the mean taken over 1/inv_batch samples.
 * adv_stats: NULL, or device {mean, std+1e-8} (from pm_normalize) applied on the fly (mini_adv_norm). */
size_t pm_ppo_actor_loss_ws_bytes(int B, int A);
int pm_ppo_actor_loss(const float* mu, int64_t ldmu, const float* log_std, const float* actions,
                      const float* logp_old, const float* mu_old, const float* sigma_old,
                      const float* adv, const float* adv_stats, int B, int A, float inv_batch,
                      float eps_clip, float max_action, int squash, float* stats, float* dmu,
                      int64_t lddmu, float* dlog_std, float* logp_out, void* ws, pm_stream_t s);

/* device-side replacement of ppo.py:334-338,355-357 bookkeeping (the host `continue`):
 *   kl_mean = stats[1]*inv_batch; acc[3] = max(acc[3], kl_mean); skip = kl_mean > desired_kl;
 *   if !skip: acc[0] += stats[0]*inv_batch; acc[1] += kl_mean; acc[2] += 1.
 * skip_flag (int32) gates pm_adam_step. */
int pm_ppo_actor_finalize(const float* stats, float inv_batch, float desired_kl, float* acc,
                          int32_t* skip_flag, pm_stream_t s);

/* replaces ppo.py:368-374: value loss (plain or clipped) forward + d/dv.  stats[0] += nothing;
 * writes stats[0] = sum_b (v-target)^2 (LOCAL).  clip_delta: device scalar = mean|eps*old_v|
 * (from pm_abs_mean) when clipped, else NULL. */
int pm_value_loss(const float* v, int64_t ldv, const float* returns, const float* old_values,
                  const float* clip_delta, int B, float inv_batch, float* stats, float* dv,
                  int64_t lddv, void* ws, pm_stream_t s);
/* replaces dagger.py:312-319 loss + backward to the student mean: stats[0] = mean((tea_act - act(mu))^2) over
 * 1/inv_count elements (LOCAL sum * inv_count), dmu = d(loss)/d(mu).  tea_act: (B,A) contiguous. */
int pm_dagger_loss(const float* mu, int64_t ldmu, const float* tea_act, int B, int A, float max_action, int squash,
                   float inv_count, float* stats, float* dmu, int64_t lddmu, void* ws, pm_stream_t s);
/* out[0] = scale * sum|x| (deterministic); used for delta_value_clipped (ppo.py:370) */
int pm_abs_sum(const float* x, int64_t n, float scale, float* out, void* ws, pm_stream_t s);
/* acc[idx] += stats[0]*scale  (mean_value_loss accumulation, ppo.py:384) */
int pm_accumulate(const float* stats, float scale, float* acc, int idx, pm_stream_t s);

/* ------------------------------------------------------------------------------------------
 * K3  dense layers (MLP network.py:27-54; PointNet head network.py:152-159), fp32
 * ------------------------------------------------------------------------------------------ */
/* y[M,N] = act(x[M,K] W[N,K]^T + b[N]).  m_dev: NULL or device int32 row count (<= M). */
int pm_linear_forward(const float* x, int64_t ldx, const float* W, const float* b, float* y,
                      int64_t ldy, int M, int N, int K, int act, const int32_t* m_dev, pm_stream_t s);
/* backward of one layer given dpre[M,N] = dL/d(pre-activation of this layer):
 *   dW[N,K] = dpre^T x ; db[N] = colsum(dpre) ;
 *   dx[M,K] = (dpre W) * act'(x) where x is the PREVIOUS layer's activated output and
 *   act_prev its activation (PM_ACT_NONE: plain dx).  dx may be NULL (first layer). */
size_t pm_linear_backward_ws_bytes(int M, int N, int K);
int pm_linear_backward(const float* x, int64_t ldx, const float* W, const float* dpre, int64_t lddpre,
                       float* dW, float* db, float* dx, int64_t lddx, int M, int N, int K,
                       int act_prev, const int32_t* m_dev, void* ws, pm_stream_t s);

/* ------------------------------------------------------------------------------------------
 * K3b  fused PointNet head  Linear(F,128)-act-Linear(128,32)-act-Linear(32,out)   (network.py:152-159, 186-198)
 * One launch forward (16 batch rows per CTA through all three layers in shared memory; h1[B,128] and
 * h2[B,32] are kept for the backward), three launches backward.  out <= 32.  fp32.
 * dfeat (may be NULL) receives d/d(feat) for the first dfeat_cols columns (the pooled part; the proprio
 * tail is an input).  Gradients are OVERWRITTEN.
 * PM_PREC_BF16: the three products with the F-wide dimension (feat.W0^T, dPre1.W0, dPre1^T.feat) and the 128->32 layer
 * run as tcgen05 MMAs with bf16 operands / fp32 accumulation (1e-2 gate); PM_PREC_FP32: FFMA kernels (1e-4 gate).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const float *W0, *b0; /* (128,F),(128)    final_mlp.0 */
  const float *W1, *b1; /* (32,128),(32)    final_mlp.2 */
  const float *W2, *b2; /* (out,32),(out)   final_mlp.4 */
} pm_head_params;
typedef struct {
  float *W0, *b0, *W1, *b1, *W2, *b2;
} pm_head_grads;
int pm_pointnet_head_forward(const float* feat, int64_t ldf, int B, int F, const pm_head_params* p, int out_dim,
                             int act, int precision, float* h1, float* h2, float* out, int64_t ldo, pm_stream_t s);
size_t pm_pointnet_head_backward_ws_bytes(int B, int F);
int pm_pointnet_head_backward(const float* feat, int64_t ldf, int B, int F, const pm_head_params* p, int out_dim,
                              int act, int precision, const float* h1, const float* h2, const float* dout, int64_t lddo,
                              const pm_head_grads* g, float* dfeat, int64_t lddf, int dfeat_cols, void* ws,
                              size_t ws_bytes, pm_stream_t s);

/* ------------------------------------------------------------------------------------------
 * K1/K2  PointNet encoder: per-point MLP C->128->256->512 + symmetric pooling
 * replaces network.py:165-182 (forward up to the pooled feature) and its autograd backward.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const float *W1, *b1; /* (128,C),(128)   mlp.0 */
  const float *W2, *b2; /* (256,128),(256) mlp.2 */
  const float *W3, *b3; /* (512,256),(512) mlp.4 */
} pm_encoder_params;
typedef struct {
  float *W1, *b1, *W2, *b2, *W3, *b3;
} pm_encoder_grads;

/* in-place xyz centring through the caller's tensor (network.py:172-173) */
int pm_pointnet_center(float* x, int64_t ldx, int B, int N, int C, pm_stream_t s);

/* feat[b, 0:512] = max_n h3[b,n,:]; if feat_mean: feat_mean[b, 0:512] = mean_n h3[b,n,:]
 * (both with row stride ldf; the caller lays them out as cat(max, mean, proprio)).
 * argmax[b,512] (int32, may be NULL) = first point index attaining the max (needed by backward).
 * h2mean[b,256] (may be NULL) = mean_n h2[b,n,:] (needed by the mean-branch backward). */
int pm_pointnet_encode_forward(const float* x, int64_t ldx, int B, int N, int C,
                               const pm_encoder_params* p, int act, int precision, float* feat,
                               float* feat_mean, int64_t ldf, int32_t* argmax, float* h2mean,
                               void* ws, size_t ws_bytes, pm_stream_t s);
size_t pm_pointnet_encode_forward_ws_bytes(int B, int N, int C, int precision);
/* diagnostic for PM_PREC_BF16: the device-side protocol error word the last launch left in `ws`
 * (0 = clean; 1xx = a bounded mbarrier wait expired).  Synchronises the stream. */
int pm_pointnet_tc_last_error(const void* ws, pm_stream_t s);
/* the same diagnostic for a PM_PREC_FP32 launch that ran on the split-fp16 tcgen05 kernel */
int pm_pointnet_tc3_last_error(const void* ws, pm_stream_t s);

/* Backward through max-pool + per-point MLP.  The max-pool routes dfeat[b,c] to the single point
 * argmax[b,c], so only the unique "critical" points of each cloud carry gradient: their
 * activations are recomputed from the 4C-byte inputs and layers 3..1 are back-propagated over
 * those rows only.  Identical to autograd's result.  dfeat_mean != NULL adds the
 * dense mean-pool branch (max_mean=True).  Gradients are OVERWRITTEN.
 * PM_PREC_BF16: one fused tcgen05 kernel that treats every (cloud, channel) pair as a row (bf16 operands,
 * fp32 accumulation in TMEM / registers), then a fixed-order reduce of per-CTA partials. */
size_t pm_pointnet_encode_backward_ws_bytes(int B, int N, int C, int with_mean, int precision);
int pm_pointnet_encode_backward(const float* x, int64_t ldx, int B, int N, int C,
                                const pm_encoder_params* p, int act, int precision, const float* dfeat,
                                const float* dfeat_mean, int64_t lddf, const int32_t* argmax,
                                const float* h2mean, const pm_encoder_grads* g, void* ws,
                                size_t ws_bytes, pm_stream_t s);
/* diagnostic for PM_PREC_BF16 backward: protocol error word of the last launch (0 = clean).  Synchronises. */
int pm_pointnet_bwd_tc_last_error(const void* ws, pm_stream_t s);

/* ------------------------------------------------------------------------------------------
 * K3t  the same dense layers on the tensor cores (tcgen05.mma, fp32 accumulators in TMEM) — the state policy
 * MLP 53 -> 512^3 -> 10 that `--algocfg ppo --taskcfg open_drawer` trains (cfg/algos/ppo.yaml:44-47, network.py:27-54),
 * DAgger's teacher, and the dense layers of the fp32 critical-point encoder backward.  Tensors stay fp32 in HBM; tiles are
 * converted on the fly.  precision: PM_PREC_BF16 = bf16 operands (1e-2 gate); PM_PREC_FP32 = every value split into three
 * bf16 terms (24 mantissa bits, fp32 exponent range) and six term-pair MMAs per product (1e-4 gate).  Same argument
 * meaning as pm_linear_forward / pm_linear_backward.
 * ------------------------------------------------------------------------------------------ */
int pm_linear_forward_tc(const float* x, int64_t ldx, const float* W, const float* b, float* y, int64_t ldy, int M, int N, int K,
                         int act, int precision, const int32_t* m_dev, pm_stream_t s);
size_t pm_linear_backward_tc_ws_bytes(int M, int N, int K);
int pm_linear_backward_tc(const float* x, int64_t ldx, const float* W, const float* dpre, int64_t lddpre, float* dW, float* db,
                          float* dx, int64_t lddx, int M, int N, int K, int act_prev, int precision, const int32_t* m_dev,
                          void* ws, pm_stream_t s);

/* ------------------------------------------------------------------------------------------
 * K7  grad-norm clip + Adam on flat buffers
 * replaces nn.utils.clip_grad_norm_ + torch.optim.Adam.step (ppo.py:351-353, 381-382).
 * The flat buffer is [clipped params (n_clip) | unclipped tail (log_std, Q8)].
 * opt_state (device, 8 floats): [0]=step (float, incremented here unless skipped), [1]=lr,
 *   [2..7] scratch written by pm_adam_prepare: total_norm, clip_coef, step_size, 1/sqrt(bc2), skipped.
 * ------------------------------------------------------------------------------------------ */
size_t pm_adam_ws_bytes(int64_t n);
int pm_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                 int64_t n_clip, float max_norm /* <=0: no clipping */, float beta1, float beta2,
                 float eps, float* opt_state, const int32_t* skip_flag, void* ws, pm_stream_t s);

/* K7f  ONE launch per optimiser step: cross-GPU gradient all-reduce (one-shot pull over NVLink peer memory) + KL-skip decision
 * (pm_ppo_actor_finalize's arithmetic, ppo.py:335-338) + clip + Adam (pm_adam_step's arithmetic).  grad_local: this rank's
 * gradients followed by n_tail extra floats that are summed over ranks too (finalize != 0: tail[0] = sum surrogate, tail[1] = sum
 * KL).  world > 1: grad_peers_dev / flag_peers_dev are DEVICE arrays of `world` pointers to every rank's gradient buffer /
 * 64-word flag buffer (symmetric memory mapped into this process; entry `rank` is the local one), flags_local = this rank's flag
 * buffer, zero before the first call; every rank must call in lockstep.  The workspace (pm_fused_step_ws_bytes) must be ZEROED
 * before the first call and then belong to this optimiser only (it carries the launch sequence number). */
size_t pm_fused_step_ws_bytes(int64_t n, int n_tail);
int pm_fused_step(float* params, float* exp_avg, float* exp_avg_sq, int64_t n, int64_t n_clip, int n_tail, float max_norm, float beta1,
                  float beta2, float eps, float* opt_state, const float* grad_local, const float* const* grad_peers_dev,
                  uint32_t* const* flag_peers_dev, uint32_t* flags_local, int rank, int world, int finalize, float inv_batch,
                  float desired_kl, float* acc, int32_t* skip_flag, void* ws, pm_stream_t s);

/* ------------------------------------------------------------------------------------------
 * K8  rollout-buffer helpers (storage.py:43-56 is cudaMemcpyAsync; the random sampler needs gathers)
 * ------------------------------------------------------------------------------------------ */
/* out[i,:] = src[idx[i],:]   (x[list] gathers of ppo.py:317-324 for sampler=random) */
int pm_gather_rows(const float* src, int64_t lds, const int64_t* idx, float* out, int64_t ldo,
                   int64_t n_rows, int width, pm_stream_t s);
/* dst[rows,width] <- src (strided 2-D copy; add_transitions / add_transitions_dagger) */
int pm_copy_rows(const float* src, int64_t lds, float* dst, int64_t ldd, int64_t n_rows, int width,
                 pm_stream_t s);

/* ------------------------------------------------------------------------------------------
 * NEXT ROW  depth -> world point cloud -> farthest-point subsample
 * replaces utils/depth2tsdf.py:146-159 (back-projection, camera->world, workspace mask: points outside the open box
 * (origin, origin+size) become (0,0,0)) and :160 (pytorch3d.ops.sample_farthest_points, start index 0, first index on ties).
 * cam_intr (3x3) and vol_origin (3) are HOST arrays; cam_pose is a device (M,4,4) row-major array.
 * ------------------------------------------------------------------------------------------ */
int pm_depth2pc_backproject(const float* depth, int E, int M, int H, int W, const float* cam_intr, const float* cam_pose_dev,
                            const float* vol_origin, float size, float* out /* (E, M*H*W, 3) */, pm_stream_t s);
/* The same, reading the E*M camera images in place through a device table of pointers (view index = e*M + m), with the stacking
 * step of tasks/hand_base.py:317-324 folded in: d = -image when negate != 0, then +-inf -> inf_value (100 there).  aligned16 = every
 * image pointer is 16-byte aligned (enables 16-byte loads). */
int pm_depth2pc_backproject_views(const float* const* depth_views_dev, int E, int M, int H, int W, int aligned16, int negate,
                                  float inf_value, const float* cam_intr, const float* cam_pose_dev, const float* vol_origin, float size,
                                  float* out /* (E, M*H*W, 3) */, pm_stream_t s);
size_t pm_fps_ws_bytes(int E, int P);
/* how many 8-CTA clusters (= clouds) of the cluster sampler the device keeps resident at once (cudaOccupancyMaxActiveClusters) */
int pm_fps_cluster_max_active(void);
/* compact != 0 (needs P % 4 == 0): the cloud is first compacted to its non-zero points + the first zero point (identical
 * picks: the masked points are exact duplicates of (0,0,0)), so the K passes touch only the valid points.
 * compact: 0 = none, 1 = auto (P > 48 Ki -> one 8-CTA cluster per cloud, else one CTA per cloud), 2 = cluster, 3 = one CTA. */
int pm_farthest_point_sample(const float* points /* (E,P,3) */, int E, int P, int K, int compact, float* out /* (E,K,3) */,
                             int64_t* out_idx /* (E,K) or NULL */, void* ws, size_t ws_bytes, pm_stream_t s);

/* ------------------------------------------------------------------------------------------
 * NEXT ROW (first half): depth images -> fused TSDF volume (`depth_tsdf` observations)
 * pm_tsdf_voxel_tables replaces utils/depth2tsdf.py:14-62 (TSDFVolume.__init__ voxel grid + register_camera's voxel -> pixel
 * projection): pix_off (M, R^3) = row * W + col of the pixel voxel v = x*R*R + y*R + z projects to in view m, or -1 when it falls
 * outside the image / behind the camera; pix_z (M, R^3) = its camera-space depth.  cam_intr (3x3), vol_origin (3): HOST arrays.
 * pm_tsdf_integrate replaces utils/depth2tsdf.py:68-86 (TSDFVolume.integrate): depth (E,M,H,W) -> out (E, R, R, R).
 * ------------------------------------------------------------------------------------------ */
int pm_tsdf_voxel_tables(const float* cam_pose_dev /* (M,4,4) */, int M, const float* cam_intr, int H, int W, float size, int resolution,
                         const float* vol_origin, int32_t* pix_off, float* pix_z, pm_stream_t s);
int pm_tsdf_integrate(const float* depth, int E, int M, int H, int W, const int32_t* pix_off, const float* pix_z, float size,
                      int resolution, float default_tsdf, float* out, pm_stream_t s);
/* replaces utils/depth2tsdf.py:103-119 (TSDFVolume.sparse_voxel after the fusion): voxels with lo < tsdf < hi (0.2 / -0.2 there) in
 * row-major order -> K farthest points on their integer coordinates (pytorch3d semantics: start at the first, first index on
 * ties) -> out (E, K, 4) = (x, y, z, tsdf).  An env whose band is empty returns voxel 0 K times (the reference fails there). */
size_t pm_tsdf_sparse_voxel_ws_bytes(int E, int resolution, int K);
int pm_tsdf_sparse_voxel(const float* tsdf /* (E,R,R,R) */, int E, int resolution, float lo, float hi, int K, float* out, void* ws,
                         size_t ws_bytes, pm_stream_t s);

/* ------------------------------------------------------------------------------------------
 * NEXT ROW (second half): the Conv3D student on the fused TSDF volume — algorithms/algo_utils/network.py:56-63
 * (conv_stride = nn.Conv3d(stride, padding = k // 2)), :67-97 (Conv3DNet), :119-135 (Encoder).  A convolution = pm_conv3d_im2col
 * + pm_linear_forward_tc (bias + activation fused) on CHANNELS-LAST activations ((sample, voxel) rows x channels); its backward =
 * pm_linear_backward_tc + pm_conv3d_col2im.  The patch column order is TAP-major, (kd, kh, kw, c): a tap's channels are contiguous in
 * the activations and in the patch row (16-byte vector moves when C % 4 == 0); pm_conv3d_weight_permute brings nn.Conv3d's
 * weight.view(Cout, C, k^3) into that order (to_tap_major != 0) and a weight gradient back (to_tap_major == 0).
 * ------------------------------------------------------------------------------------------ */
int pm_conv3d_out_dim(int Din, int k, int s);
/* cols[(b, od, oh, ow), (kd, kh, kw, c)] = in[b*sample_stride + voxel*ld_in + c] (0 outside the volume); row stride Kpad >= C*k^3,
 * padding columns zeroed.  in: voxel (id, ih, iw) of sample b at in + b*sample_stride + ((id*Din + ih)*Din + iw)*ld_in. */
int pm_conv3d_im2col(const float* in, int64_t ld_in, int64_t sample_stride, int B, int C, int Din, int k, int s, float* cols, int Kpad,
                     pm_stream_t st);
/* adjoint of im2col in gather form (deterministic), times act'(y): din[(b, voxel), c] = (sum of covering patch entries) * act'(y[...]) */
int pm_conv3d_col2im(const float* dcols, int Kpad, int B, int C, int Din, int k, int s, const float* y, int act, float* din,
                     pm_stream_t st);
int pm_conv3d_weight_permute(const float* src, int Cout, int C, int k, int to_tap_major, float* dst, pm_stream_t st);
/* The first layer, Conv3d(1, 16, k 5, stride `stride`, padding 2) + activation (network.py:70, 104, 122), directly on the volume rows
 * x (B, >= Din^3) in exact fp32 — no patch matrix (it would be 5 GB at 2048 samples): y ((b, voxel'), 16) channels-last.
 * pm_conv3d_first_backward: dW (16,1,5,5,5) from dpre ((b, voxel'), 16); its bias gradient is a column sum of dpre. */
int pm_conv3d_first_forward(const float* x, int64_t ldx, int B, int Din, int stride, const float* w, const float* bias, int act, float* y,
                            pm_stream_t st);
size_t pm_conv3d_first_backward_ws_bytes(void);
int pm_conv3d_first_backward(const float* x, int64_t ldx, int B, int Din, int stride, const float* dpre, float* dW, void* ws, pm_stream_t st);
/* nn.MaxPool3d(kernel_size = k) (stride k, no padding) of PoolConv3DNet (network.py:100-117) on channels-last rows y ((b, voxel), C):
 * out ((b, cell), C), argmax ((b, cell), C) = index of the winning voxel inside the sample (first maximum in (d, h, w) scan order).
 * backward: dpre ((b, voxel), C) = dout routed to the winning voxel, times act'(y) (y = the activation output that was pooled). */
int pm_maxpool3d_forward(const float* y, int B, int C, int Din, int k, float* out, int32_t* argmax, pm_stream_t st);
int pm_maxpool3d_backward(const float* dout, const int32_t* argmax, const float* y, int act, int B, int C, int Din, int k, float* dpre,
                          pm_stream_t st);
/* to_rows != 0: out[b*ld_row + c*P + pos] = in[(b*P + pos)*C + c] (channels-last -> x.reshape(batch, -1) order, network.py:92-96);
 * to_rows == 0: out[(b*P + pos)*C + c] = in[b*ld_row + c*P + pos] */
int pm_conv3d_flatten(const float* in, float* out, int B, int P, int C, int64_t ld_row, int to_rows, pm_stream_t st);

/* ------------------------------------------------------------------------------------------
 * NEXT ROW (third part): TSDF of the scene from per-part signed-distance grids (`mesh_tsdf` observations).
 * replaces utils/mesh2sdf.py:119-139 (TSDFfromMesh.query_tsdf_parallel) + :239-272 (triplet_interpolation_query_parallel):
 * sdf_field (M, field_stride) = the parts' grids padded with +1 to the common resolution (X, bbox_res_y, bbox_res_z) as
 * merge_sdf_field builds them (:169-198), sdf_res (M,3) int32 each part's own resolution, sdf_voxel (M), sdf_bbox_min (M,3);
 * pose_R (E,M,3,3), pose_T (E,M,3) part poses; init_tsdf (E, R^3) the volume the parts are min-ed into (metres, unscaled);
 * out (E, R, R, R) = clamp(min(...) / (4 * size / R), -1, 1).  vox_origin: HOST array of 3.
 * ------------------------------------------------------------------------------------------ */
int pm_mesh2sdf_query(const float* sdf_field, int64_t field_stride, const int32_t* sdf_res, const float* sdf_voxel, const float* sdf_bbox_min,
                      int M, int bbox_res_y, int bbox_res_z, const float* pose_R, const float* pose_T, const float* init_tsdf, int E,
                      int resolution, const float* vox_origin, float size, float* out, pm_stream_t st);

/* ------------------------------------------------------------------------------------------
 * NEXT ROW (second half): the env-side arithmetic of the open_drawer task around the physics step, one launch
 * per phase.  The simulator's flat state tensors are read IN PLACE through the task's own int64 index tables
 * (dof_state_mask (E, num_dofs + 1), rigid_body_mask (E, num_rigid_body + 2): tasks/open_drawer.py:58-70).
 *
 * pm_open_drawer_post_physics replaces tasks/open_drawer.py:240-281 (compute_observations, with franka.update_state
 *   tasks/load_robot.py:153-164) when do_obs, tasks/open_drawer.py:170-238 (compute_reward) when do_reward, and
 *   `progress_buf += 1` (tasks/hand_base.py:388) when advance_progress.  obs (E, 29 + 2 num_dofs) = the `normal_state` row;
 *   extras_f (6, E) = reaching_reward, close_reward, rot_reward, joint_state_reward, is_grasped, step_id;
 *   extras_b (3, E) = is_open, is_open_notgrasp, is_reached;  raw_reward aliases rew_buf as in the reference.
 * pm_franka_control replaces tasks/load_robot.py:96-118 (franka.control, driveMode 'pos' / 'ik', fixed or mobile base) and
 *   :142-151 (solve_ik: damped least squares on the mean of the two finger-tip Jacobians; the 6x6 system is solved by Cholesky
 *   instead of torch.inverse).  Joint positions: dof_state_mask != null -> qpos = the simulator's dof_state_all, read through the
 *   index table (mask_ld wide); dof_state_mask == null -> a strided view q[e, j] = qpos[e * qpos_row_stride + j * qpos_elem_stride]
 *   (franka.dof_qpos_raw as update_state leaves it).  jacobian (E, n_links, 6, num_dofs); default_root_quat: HOST array of 4 (x, y, z, w);
 *   jacobian_sum (device scalar or null) receives sum(j_eef), the value the reference tests against 1e-5 before exit(1).
 * pm_episode_flags replaces tasks/hand_base.py:367-377: train != 0 updates epis_max_step / epis_max_rew, writes reset_buf,
 *   reset_succ and succ_rate = sum(success) / max(sum(reset_buf), 1); train == 0 writes reset_buf = progress >= max_episode_length.
 *   counts3 (3 x int32, device): [0] = sum(success), [1] = sum(reset_buf) (the value `if self.reset_buf.sum() > 0` reads), [2] scratch.
 * pm_scatter_dof_targets replaces tasks/hand_base.py:382: pos_act_all[dof_state_mask[:, :num_dofs]] = pos_act.
 * ------------------------------------------------------------------------------------------ */
enum { PM_DRIVE_POS = 0, PM_DRIVE_IK = 1 };
int pm_open_drawer_obs_dim(int num_dofs);
int pm_open_drawer_post_physics(const float* dof_state_all, const float* rigid_body_all, const float* root_tensor, int n_actors,
                                int obj_actor, const int64_t* dof_state_mask, const int64_t* rigid_body_mask, int E, int num_dofs,
                                int num_rigid_body, int ltip_rb_index, int rtip_rb_index, const float* dof_lower, const float* dof_upper,
                                const float* part_bbox_init, const float* part_axis_dir_init, const float* part_joint_lower,
                                const float* part_joint_upper, const int64_t* obj_lstid, float suc_prop, int do_obs, int do_reward,
                                int advance_progress, int64_t* progress_buf, float* obs, float* part_bbox, float* dof_state,
                                float* rigid_body, float* tip_rb, float* tip_rot_9d, float* gripper_length, float* dof_qpos_normalized,
                                float* rew_buf, uint8_t* success, uint8_t* succ_objid, float* extras_f, uint8_t* extras_b, pm_stream_t s);
/* grasp_cube: replaces tasks/grasp_cube.py:118-138 (compute_observations, incl. utils/torch_jit_utils.py:412-425 deambiguity_rotation and
 * franka.update_state) when do_obs, tasks/grasp_cube.py:66-115 (compute_reward) when do_reward.  The simulator tensors are regular here:
 * dof_state (E, dofs_per_env, 2), rigid_body (E, bodies_per_env, 13), root_tensor (E, n_actors, 13).  pose_lower_limit / pose_upper_limit
 * (7), success_pos (3), obj_default_pos (3): HOST arrays.  obs (E, 19 + 2 num_dofs) = `normal_state`; proprio (E, 7 + 2 num_dofs) or null
 * = `proprio_state`; extras_f (7, E) = reaching_reward, close_reward, rot_reward, reaching_goal_reward, obj_movement, obj_height, step_id;
 * extras_b (2, E) = is_reached, obj_up_flag. */
int pm_grasp_cube_obs_dim(int num_dofs);
int pm_grasp_cube_post_physics(const float* dof_state, int dofs_per_env, const float* rigid_body, int bodies_per_env, const float* root_tensor,
                               int n_actors, int obj_actor, int E, int num_dofs, int ltip_rb_index, int rtip_rb_index, const float* dof_lower,
                               const float* dof_upper, const float* pose_lower_limit, const float* pose_upper_limit, const float* success_pos,
                               const float* obj_default_pos, float goal_thresh, int do_obs, int do_reward, int advance_progress,
                               int64_t* progress_buf, float* obs, float* proprio, float* tip_rb, float* tip_rot_9d, float* gripper_length,
                               float* dof_qpos_normalized, float* rew_buf, uint8_t* success, float* extras_f, uint8_t* extras_b, pm_stream_t s);
int pm_franka_control(const float* raw_output, int E, int num_dofs, int mobile, int drive_mode, const float* qpos, int64_t qpos_row_stride,
                      int64_t qpos_elem_stride, const int64_t* dof_state_mask, int mask_ld, const float* jacobian, int n_links, int ltip_rb_index,
                      int rtip_rb_index, const float* dof_lower, const float* dof_upper, const float* default_root_quat, float dt,
                      float damping, float* action_tensor, float* jacobian_sum, pm_stream_t s);
int pm_episode_flags(int E, int train, const float* rew_buf, const int64_t* progress_buf, const uint8_t* success, float* epis_max_rew,
                     int64_t* epis_max_step, int64_t explore_step, int64_t max_episode_length, uint8_t* reset_buf, uint8_t* reset_succ,
                     int32_t* counts3, float* succ_rate, pm_stream_t s);
int pm_scatter_dof_targets(const float* pos_act, const int64_t* dof_state_mask, int mask_ld, int E, int num_dofs, float* pos_act_all,
                           pm_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* PARTMANIP_B200_H */
