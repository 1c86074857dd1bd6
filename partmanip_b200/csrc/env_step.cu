// E1-E4 — env-side arithmetic of the open_drawer task, one launch per phase instead of ~150.
// reference:
//   E1  tasks/open_drawer.py:240-281 compute_observations (+ tasks/load_robot.py:153-164 franka.update_state)
//       tasks/open_drawer.py:170-238 compute_reward                       -> open_drawer_post_kernel (one thread per env)
//   E2  tasks/load_robot.py:96-151 franka.control ('ik' / 'pos', fixed or mobile base) + solve_ik
//                                                                           -> franka_control_kernel (one thread per env)
//   E3  tasks/hand_base.py:367-377 pre_physics_step's episode bookkeeping  -> episode_flags_kernel
//   E4  tasks/hand_base.py:382 pos_act_all[dof_state_mask[:, :num_dofs]] = pos_act -> scatter_targets_kernel
// All HBM-bound and tiny (about 1.7 KB read + 1.4 KB written per env): what they buy is launch count, not bandwidth.  The
// simulator's flat state tensors are read in place through the task's own index tables (dof_state_mask / rigid_body_mask are
// int64 index tensors despite the name, tasks/open_drawer.py:58-70), so the masked-gather copies disappear too.
// This file is compiled with -fmad=false: every elementwise chain rounds op by op like the reference's torch expressions.
#include "common.cuh"

namespace {

constexpr int ENV_THREADS = 64;
constexpr int COPY_PER_THREAD = 8;
constexpr int MAX_DOFS = 16;                     // robot dofs (franka: 9 fixed base, 12 mobile)
constexpr int MAX_ARM = 8;                       // dofs solved by the IK (franka: 7)

struct OpenDrawerP {
  // simulator state, read in place
  const float* dof_all;                          // [total_dofs, 2]
  const float* rb_all;                           // [total_rigid_bodies, 13]
  const float* root;                             // [E, n_actors, 13]
  const int64_t* dof_mask;                       // [E, nd + 1]   row indices into dof_all; last = the drawer joint
  const int64_t* rb_mask;                        // [E, nb + 2]   row indices into rb_all
  int E, nd, nb, n_actors, obj_actor, ltip, rtip;
  const float* dof_lower;                        // [nd]
  const float* dof_upper;                        // [nd]
  const float* bbox_init;                        // [E, 8, 3]
  const float* axis_dir;                         // [E, 3]
  const float* joint_lower;                      // [E]
  const float* joint_upper;                      // [E]
  const int64_t* obj_lstid;                      // [E]
  float suc_prop;
  int do_obs, do_reward, advance_progress;
  int64_t* progress;                             // [E]
  // observation outputs
  float* obs;                                    // [E, 29 + 2 nd]
  float* part_bbox;                              // [E, 8, 3]
  float* dof_state;                              // [E, nd + 1, 2]
  float* rb_state;                               // [E, nb + 2, 13]
  float* tip_rb;                                 // [E, 13]
  float* tip_rot;                                // [E, 9]
  float* gripper;                                // [E]
  float* qpos_norm;                              // [E, nd]
  // reward outputs
  float* rew;                                    // [E]
  uint8_t* success;                              // [E]
  uint8_t* succ_objid;                           // [num_objs]
  float* extras_f;                               // [6, E]: reaching, close, rot, joint_state, is_grasped, step_id
  uint8_t* extras_b;                             // [3, E]: is_open, is_open_notgrasp, is_reached
};

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float norm(V3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
__device__ __forceinline__ V3 ld3(const float* p) { return {p[0], p[1], p[2]}; }

// utils/torch_jit_utils.py:375-403, quaternion (x, y, z, w), not normalised first
__device__ __forceinline__ void quat_to_mat(const float* q, float* m /* 9, row-major */) {
  const float i = q[0], j = q[1], k = q[2], r = q[3];
  const float two_s = 2.0f / (i * i + j * j + k * k + r * r);
  m[0] = 1 - two_s * (j * j + k * k); m[1] = two_s * (i * j - k * r);     m[2] = two_s * (i * k + j * r);
  m[3] = two_s * (i * j + k * r);     m[4] = 1 - two_s * (i * i + k * k); m[5] = two_s * (j * k - i * r);
  m[6] = two_s * (i * k - j * r);     m[7] = two_s * (j * k + i * r);     m[8] = 1 - two_s * (i * i + j * j);
}

// isaacgym.torch_utils.quat_rotate on the basis vector e_axis (utils/torch_jit_utils.py:65-69 quat_axis):
// v (2 w^2 - 1) + 2 w (q_v x v) + 2 q_v (q_v . v)
__device__ __forceinline__ V3 quat_axis(const float* q, int axis) {
  const float w = q[3];
  const V3 qv = {q[0], q[1], q[2]};
  V3 v = {axis == 0 ? 1.f : 0.f, axis == 1 ? 1.f : 0.f, axis == 2 ? 1.f : 0.f};
  const V3 a = v * (2.0f * (w * w) - 1.0f);
  const V3 cr = {qv.y * v.z - qv.z * v.y, qv.z * v.x - qv.x * v.z, qv.x * v.y - qv.y * v.x};
  const V3 b = cr * w * 2.0f;
  const V3 c = qv * dot(qv, v) * 2.0f;
  return a + b + c;
}

// One launch, two kinds of CTA (independent: both read the simulator tensors, neither reads the other's output):
//   blockIdx.x <  n_env_ctas : one env per thread — robot state, handle frame, state row, reward
//   blockIdx.x >= n_env_ctas : the gathered copies the task keeps as attributes (dof_state_tensor, rigid_body_tensor), one element
//                              per thread — as a per-env loop they were 200 dependent loads per thread and 85 % of the kernel's time
__global__ void __launch_bounds__(ENV_THREADS)
open_drawer_post_kernel(const OpenDrawerP p, const int n_env_ctas) {
  extern __shared__ float s_obs[];               // [ENV_THREADS][obs_dim] (obs_dim is odd: conflict-free rows)
  const int tid = threadIdx.x;
  const int ndp = p.nd + 1, nbp = p.nb + 2;
  const int obs_dim = 29 + 2 * p.nd;

  if ((int)blockIdx.x >= n_env_ctas) {           // COPY_PER_THREAD independent gathers in flight per thread, 32-bit index arithmetic
    const uint32_t n_rb = (uint32_t)p.E * nbp * 13, n_dof = (uint32_t)p.E * ndp * 2;
    const uint32_t base = (uint32_t)(blockIdx.x - n_env_ctas) * (ENV_THREADS * COPY_PER_THREAD) + tid;
    float v[COPY_PER_THREAD];
#pragma unroll
    for (int u = 0; u < COPY_PER_THREAD; ++u) {
      const uint32_t i = base + u * ENV_THREADS;
      v[u] = 0.f;
      if (i < n_rb) {
        const uint32_t row = i / 13u;
        v[u] = p.rb_all[p.rb_mask[row] * 13 + (i - row * 13u)];
      } else if (i - n_rb < n_dof) {
        const uint32_t j = i - n_rb;
        v[u] = p.dof_all[p.dof_mask[j >> 1] * 2 + (j & 1u)];
      }
    }
#pragma unroll
    for (int u = 0; u < COPY_PER_THREAD; ++u) {
      const uint32_t i = base + u * ENV_THREADS;
      if (i < n_rb) p.rb_state[i] = v[u];
      else if (i - n_rb < n_dof) p.dof_state[i - n_rb] = v[u];
    }
    return;
  }
  const int e0 = blockIdx.x * ENV_THREADS;
  const int n_env = min(ENV_THREADS, p.E - e0);

  // ---- one env per thread
  const int e = e0 + tid;
  if (e < p.E) {
    const float* lt = p.rb_all + p.rb_mask[(int64_t)e * nbp + p.ltip] * 13;
    const float* rt = p.rb_all + p.rb_mask[(int64_t)e * nbp + p.rtip] * 13;
    float tip[13];
#pragma unroll
    for (int i = 0; i < 13; ++i) tip[i] = (lt[i] + rt[i]) / 2;
    const float gl = norm(ld3(lt) - ld3(rt));
    const float qd = p.dof_all[p.dof_mask[(int64_t)e * ndp + p.nd] * 2];          // drawer joint position
    const float* ro = p.root + ((int64_t)e * p.n_actors + p.obj_actor) * 13;
    float R[9];
    quat_to_mat(ro + 3, R);
    const V3 op = ld3(ro);
    const V3 ax = ld3(p.axis_dir + (int64_t)e * 3);
    // corners 0, 1, 3, 4, 6 are the ones the handle frame needs; all 8 go to part_bbox
    V3 c[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const V3 b = ld3(p.bbox_init + ((int64_t)e * 8 + k) * 3) + ax * qd;
      c[k] = {b.x * R[0] + b.y * R[1] + b.z * R[2] + op.x, b.x * R[3] + b.y * R[4] + b.z * R[5] + op.y,
              b.x * R[6] + b.y * R[7] + b.z * R[8] + op.z};
    }
    V3 h_out = c[0] - c[4], h_long = c[1] - c[0], h_short = c[3] - c[0];
    const V3 mid = (c[0] + c[6]) / 2;
    const float l_out = norm(h_out), l_long = norm(h_long), l_short = norm(h_short);
    h_out = h_out / l_out; h_long = h_long / l_long; h_short = h_short / l_short;

    if (p.do_obs) {
      float* o = s_obs + tid * obs_dim;
#pragma unroll
      for (int i = 0; i < 13; ++i) o[i] = tip[i];
      o[13] = mid.x; o[14] = mid.y; o[15] = mid.z;
      o[16] = h_out.x; o[17] = h_out.y; o[18] = h_out.z;
      o[19] = h_short.x; o[20] = h_short.y; o[21] = h_short.z;
      o[22] = h_long.x; o[23] = h_long.y; o[24] = h_long.z;
      o[25] = l_out; o[26] = l_long; o[27] = l_short;
      for (int j = 0; j < p.nd; ++j) {
        const float* d = p.dof_all + p.dof_mask[(int64_t)e * ndp + j] * 2;
        const float qn = 2 * (d[0] - p.dof_lower[j]) / (p.dof_upper[j] - p.dof_lower[j]) - 1;
        o[28 + j] = qn;
        o[28 + p.nd + j] = d[1];
        p.qpos_norm[(int64_t)e * p.nd + j] = qn;
      }
      o[28 + 2 * p.nd] = qd;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float* pb = p.part_bbox + ((int64_t)e * 8 + k) * 3;
        pb[0] = c[k].x; pb[1] = c[k].y; pb[2] = c[k].z;
      }
#pragma unroll
      for (int i = 0; i < 13; ++i) p.tip_rb[(int64_t)e * 13 + i] = tip[i];
      float TR[9];
      quat_to_mat(tip + 3, TR);
#pragma unroll
      for (int i = 0; i < 9; ++i) p.tip_rot[(int64_t)e * 9 + i] = TR[i];
      p.gripper[e] = gl;
    }

    int64_t prog = 0;
    if (p.progress) {
      prog = p.progress[e];
      if (p.advance_progress) p.progress[e] = ++prog;                              // hand_base.py:388
    }

    if (p.do_reward) {
      // reaching (open_drawer.py:184-193)
      const V3 delta = V3{tip[0], tip[1], tip[2]} - mid;
      const float dist = norm(delta);
      const bool r_out = fabsf(dot(delta, h_out)) < l_out / 2;
      const float s_l = dot(ld3(lt) - mid, h_short), s_r = dot(ld3(rt) - mid, h_short);
      const bool r_short = (s_l * s_r) < 0;
      const bool r_long = fabsf(dot(delta, h_long)) < l_long / 2;
      const bool reached = r_out && r_short && r_long;
      // bool + bool stays bool in torch: the bonus is 0.1 * (out OR short OR long)
      const float reaching = -dist + 0.1f * ((r_out || r_short || r_long) ? 1.f : 0.f);
      // rotation (open_drawer.py:195-204)
      const V3 grip = quat_axis(tip + 3, 2), sep = quat_axis(tip + 3, 1), down = quat_axis(tip + 3, 0);
      const float dot1 = dot(grip * -1.f, h_out);
      const float dot2 = fmaxf(dot(sep, h_short), dot(sep * -1.f, h_short));
      const float dot3 = fmaxf(dot(down, h_long), dot(down * -1.f, h_long));
      const float rot = dot1 + dot2 + dot3 - 3;
      // close / grasp / drawer (open_drawer.py:206-216)
      const float close = (0.1f - gl) * (reached ? 1.f : 0.f) + 0.1f * (gl - 0.1f) * (reached ? 0.f : 1.f);
      const bool grasp = reached && (gl < l_short + 0.01f) && (rot > -0.2f);
      const float jl = p.joint_lower[e], ju = p.joint_upper[e];
      const float frac = (qd - jl) / ju;
      const float joint = (grasp ? 1.f : 0.f) * (0.1f + fminf(frac, p.suc_prop));
      const bool open_ng = frac > 0.1f;
      float rew = reaching + 0.5f * rot + 5 * close + 5 * joint;
      rew = rew + fabsf(rew) * rot;
      const bool succ = grasp && ((qd - jl) >= p.suc_prop * ju);
      if (succ) p.succ_objid[p.obj_lstid[e]] = 1;                                  // every writer stores the same value
      rew += 2 * (succ ? 1.f : 0.f);
      p.rew[e] = rew;
      p.success[e] = succ;
      const int64_t E = p.E;
      p.extras_f[0 * E + e] = reaching;
      p.extras_f[1 * E + e] = close;
      p.extras_f[2 * E + e] = rot;
      p.extras_f[3 * E + e] = joint;
      p.extras_f[4 * E + e] = grasp ? 1.f : 0.f;
      p.extras_f[5 * E + e] = (float)prog;
      p.extras_b[0 * E + e] = grasp && open_ng;
      p.extras_b[1 * E + e] = open_ng;
      p.extras_b[2 * E + e] = reached;
    }
  }

  // ---- phase C: the 29 + 2 nd state rows, written coalesced
  if (p.do_obs) {
    __syncthreads();
    float* dst = p.obs + (int64_t)e0 * obs_dim;
    for (int i = tid; i < n_env * obs_dim; i += ENV_THREADS) dst[i] = s_obs[i];
  }
}

// ---------------------------------------------------------------------------------------------------------------- grasp_cube
struct GraspCubeP {
  const float* dof;                              // [E, nd_all, 2] (the simulator's dof tensor viewed per env; robot dofs first)
  const float* rb;                               // [E, nb, 13]
  const float* root;                             // [E, n_actors, 13]
  int E, nd, nd_all, nb, n_actors, obj_actor, ltip, rtip;
  const float* dof_lower;
  const float* dof_upper;
  float pose_lo[7], pose_hi[7];                  // pose_lower_limit / pose_upper_limit (grasp_cube.py:18-21)
  float success_pos[3], obj_default[3], goal_thresh;
  int do_obs, do_reward, advance_progress;
  int64_t* progress;
  float* obs;                                    // [E, 19 + 2 nd]
  float* proprio;                                // [E, 7 + 2 nd] or null
  float* tip_rb;                                 // [E, 13]
  float* tip_rot;                                // [E, 9]
  float* gripper;                                // [E]
  float* qpos_norm;                              // [E, nd]
  float* rew;
  uint8_t* success;
  float* extras_f;                               // [7, E]: reaching, close, rot, reaching_goal, obj_movement, obj_height, step_id
  uint8_t* extras_b;                             // [2, E]: is_reached, obj_up_flag
};

// utils/torch_jit_utils.py:412-425 deambiguity_rotation: of the 24 matrices built from ordered column pairs of R with the row
// sign patterns below (third column = cross product), the one closest to the identity (smallest acos((trace - 1) / 2); first
// index on ties, torch.argmin).
__device__ __forceinline__ void deambiguity_rotation(const float* q, float* best /* 9 */) {
  float R[9];
  quat_to_mat(q, R);
  const int ind[6][2] = {{0, 1}, {0, 2}, {1, 2}, {1, 0}, {2, 0}, {2, 1}};
  float best_rad = 1e30f;
#pragma unroll 1
  for (int k = 0; k < 24; ++k) {
    const int c0 = ind[k % 6][0], c1 = ind[k % 6][1];
    float a[3] = {R[0 * 3 + c0], R[1 * 3 + c0], R[2 * 3 + c0]};
    float b[3] = {R[0 * 3 + c1], R[1 * 3 + c1], R[2 * 3 + c1]};
    if (k < 12) { a[0] = -a[0]; b[0] = -b[0]; }              // all_r_mat_12[:, :12, 0]   : ROW 0 of both columns
    if (k >= 6 && k < 18) { a[1] = -a[1]; b[1] = -b[1]; }    // all_r_mat_12[:, 6:18, 1]  : ROW 1 of both columns
    const float c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    const float t = (a[0] + b[1] + c[2] - 1) / 2;
    const float rad = acosf(fminf(fmaxf(t, -1.f), 1.f));
    if (rad < best_rad) {
      best_rad = rad;
      best[0] = a[0]; best[1] = b[0]; best[2] = c[0];
      best[3] = a[1]; best[4] = b[1]; best[5] = c[1];
      best[6] = a[2]; best[7] = b[2]; best[8] = c[2];
    }
  }
}

__global__ void __launch_bounds__(ENV_THREADS)
grasp_cube_post_kernel(const GraspCubeP p) {
  extern __shared__ float s_obs[];
  const int tid = threadIdx.x;
  const int e0 = blockIdx.x * ENV_THREADS, e = e0 + tid;
  const int n_env = min(ENV_THREADS, p.E - e0);
  const int obs_dim = 19 + 2 * p.nd;
  if (e < p.E) {
    const float* lt = p.rb + ((int64_t)e * p.nb + p.ltip) * 13;
    const float* rt = p.rb + ((int64_t)e * p.nb + p.rtip) * 13;
    float tip[13];
#pragma unroll
    for (int i = 0; i < 13; ++i) tip[i] = (lt[i] + rt[i]) / 2;
    const float gl = norm(ld3(lt) - ld3(rt));
    const float* ro = p.root + ((int64_t)e * p.n_actors + p.obj_actor) * 13;
    const V3 opos = ld3(ro);
    float orot[9];
    deambiguity_rotation(ro + 3, orot);
    if (p.do_obs) {
      float* o = s_obs + tid * obs_dim;
      float* pr = p.proprio ? p.proprio + (int64_t)e * (7 + 2 * p.nd) : nullptr;
#pragma unroll
      for (int i = 0; i < 7; ++i) {                // grasp_cube.py:122
        o[i] = 2 * (tip[i] - p.pose_lo[i]) / (p.pose_hi[i] - p.pose_lo[i]) - 1;
        if (pr) pr[i] = o[i];
      }
      o[7] = 2 * (opos.x - p.pose_lo[0]) / (p.pose_hi[0] - p.pose_lo[0]) - 1;
      o[8] = 2 * (opos.y - p.pose_lo[1]) / (p.pose_hi[1] - p.pose_lo[1]) - 1;
      o[9] = 2 * (opos.z - p.pose_lo[2]) / (p.pose_hi[2] - p.pose_lo[2]) - 1;
#pragma unroll
      for (int i = 0; i < 9; ++i) o[10 + i] = orot[i];
      for (int j = 0; j < p.nd; ++j) {
        const float* d = p.dof + ((int64_t)e * p.nd_all + j) * 2;
        const float qn = 2 * (d[0] - p.dof_lower[j]) / (p.dof_upper[j] - p.dof_lower[j]) - 1;
        o[19 + j] = qn;
        o[19 + p.nd + j] = d[1];
        p.qpos_norm[(int64_t)e * p.nd + j] = qn;
        if (pr) { pr[7 + j] = qn; pr[7 + p.nd + j] = d[1]; }
      }
#pragma unroll
      for (int i = 0; i < 13; ++i) p.tip_rb[(int64_t)e * 13 + i] = tip[i];
      float TR[9];
      quat_to_mat(tip + 3, TR);
#pragma unroll
      for (int i = 0; i < 9; ++i) p.tip_rot[(int64_t)e * 9 + i] = TR[i];
      p.gripper[e] = gl;
    }
    int64_t prog = 0;
    if (p.progress) {
      prog = p.progress[e];
      if (p.advance_progress) p.progress[e] = ++prog;
    }
    if (p.do_reward) {                             // grasp_cube.py:70-114
      const float dist = norm(V3{tip[0], tip[1], tip[2]} - opos);
      const bool reached = dist < 0.02f;
      const float reaching = -dist;
      const float close = (0.1f - gl) * (reached ? 1.f : 0.f) + 0.1f * (gl - 0.1f) * (reached ? 0.f : 1.f);
      float H[9];
      quat_to_mat(tip + 3, H);
      const float down = -H[8];
      float par1 = 0.f, par2 = 0.f;                // sums over the row index of |H[:,0] O[:,0]| + |H[:,1] O[:,1]| (resp. crossed)
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        par1 += fabsf(H[r * 3 + 0] * orot[r * 3 + 0]) + fabsf(H[r * 3 + 1] * orot[r * 3 + 1]);
        par2 += fabsf(H[r * 3 + 0] * orot[r * 3 + 1]) + fabsf(H[r * 3 + 1] * orot[r * 3 + 0]);
      }
      const float rot = down + fmaxf(par1, par2) - 3;
      const float gdist = norm(opos - V3{p.success_pos[0], p.success_pos[1], p.success_pos[2]});
      const float goal = fmaxf(0.2f - gdist, 0.f) * (reached ? 1.f : 0.f);
      float rew = reaching + 0.5f * rot + 5 * close + 20 * goal;
      const bool succ = (gdist <= p.goal_thresh) && reached;
      rew += 3 * (succ ? 1.f : 0.f);
      p.rew[e] = rew;
      p.success[e] = succ;
      const int64_t E = p.E;
      p.extras_f[0 * E + e] = reaching;
      p.extras_f[1 * E + e] = close;
      p.extras_f[2 * E + e] = rot;
      p.extras_f[3 * E + e] = goal;
      p.extras_f[4 * E + e] = norm(opos - V3{p.obj_default[0], p.obj_default[1], p.obj_default[2]});
      p.extras_f[5 * E + e] = opos.z;
      p.extras_f[6 * E + e] = (float)prog;
      p.extras_b[0 * E + e] = reached;
      p.extras_b[1 * E + e] = opos.z > 0.1f;
    }
  }
  if (p.do_obs) {
    __syncthreads();
    float* dst = p.obs + (int64_t)e0 * obs_dim;
    for (int i = tid; i < n_env * obs_dim; i += ENV_THREADS) dst[i] = s_obs[i];
  }
}

struct FrankaP {
  const float* raw;                              // [E, na_in] policy output in [-1, 1]
  int E, nd, mobile, drive;                      // drive: 0 = 'pos', 1 = 'ik'
  const float* qpos;                             // current joint positions: dof_state_all rows through dof_mask (mask_ld wide), or,
  const int64_t* dof_mask; int mask_ld;          // with dof_mask == null, a strided view q[e, j] = qpos[e * row_stride + j * elem_stride]
  int64_t row_stride, elem_stride;
  const float* jac;                              // [E, n_links, 6, nd] (ik)
  int n_links, ltip, rtip;
  const float* dof_lower;
  const float* dof_upper;
  float root_rt[9];                              // quat_to_mat(default_root[3:7])
  float dt, damping2;
  float* action;                                 // [E, nd]
  float* jsum;                                   // [1] sum over every entry of the mean tip Jacobian (the reference's sanity check)
};

__global__ void __launch_bounds__(ENV_THREADS)
franka_control_kernel(const FrankaP p) {
  __shared__ float red[32];
  const int e = blockIdx.x * ENV_THREADS + threadIdx.x;
  float jpart = 0.f;
  if (e < p.E) {
    const int nd = p.nd, lo = p.mobile ? 3 : 0, na = nd - 2 - lo;
    const float* raw = p.raw + (int64_t)e * ((p.drive ? 7 : 8) + lo);
    float q[MAX_DOFS], act[MAX_DOFS];
    for (int j = 0; j < nd; ++j)
      q[j] = p.dof_mask ? p.qpos[p.dof_mask[(int64_t)e * p.mask_ld + j] * 2] : p.qpos[(int64_t)e * p.row_stride + j * p.elem_stride];
    float base[3] = {0.f, 0.f, 0.f};
    if (p.mobile) {                                // load_robot.py:97-101: base displacement in the robot's root frame
#pragma unroll
      for (int k = 0; k < 3; ++k) base[k] = raw[k] * 0.005f;
#pragma unroll
      for (int k = 0; k < 3; ++k)                  // bmm(R^T, d)[k] = sum_i R[i][k] d[i]
        act[k] = q[k] + (p.root_rt[0 * 3 + k] * base[0] + p.root_rt[1 * 3 + k] * base[1] + p.root_rt[2 * 3 + k] * base[2]);
      raw += 3;
    }
    if (p.drive == 0) {                            // 'pos' (load_robot.py:103-107)
      for (int j = 0; j < na; ++j) act[lo + j] = q[lo + j] + raw[j] * p.dt * 20;
      act[nd - 2] = q[nd - 2] + raw[na] * p.dt;
      act[nd - 1] = q[nd - 1] + raw[na] * p.dt;
    } else {                                       // 'ik' (load_robot.py:108-118, 142-151)
      float dp[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) dp[k] = raw[k] * 0.005f;
      if (p.mobile)
#pragma unroll
        for (int k = 0; k < 3; ++k) dp[k] -= base[k];
      float J[6][MAX_ARM];
      const float* jl = p.jac + ((int64_t)e * p.n_links + (p.ltip - 1)) * 6 * nd;
      const float* jr = p.jac + ((int64_t)e * p.n_links + (p.rtip - 1)) * 6 * nd;
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int k = 0; k < MAX_ARM; ++k) {
          J[r][k] = k < na ? (jl[r * nd + lo + k] + jr[r * nd + lo + k]) / 2 : 0.f;
          jpart += J[r][k];
        }
      // A = J J^T + damping^2 I (symmetric positive definite): Cholesky A = L L^T, solve L z = dp, L^T y = z, u = J^T y
      float L[6][6];
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) {
          float s = r == c ? p.damping2 : 0.f;
#pragma unroll
          for (int k = 0; k < MAX_ARM; ++k) s += J[r][k] * J[c][k];
#pragma unroll
          for (int k = 0; k < c; ++k) s -= L[r][k] * L[c][k];
          L[r][c] = r == c ? sqrtf(s) : s / L[c][c];
        }
      float z[6], y[6];
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        float s = dp[r];
#pragma unroll
        for (int k = 0; k < r; ++k) s -= L[r][k] * z[k];
        z[r] = s / L[r][r];
      }
#pragma unroll
      for (int r = 5; r >= 0; --r) {
        float s = z[r];
#pragma unroll
        for (int k = r + 1; k < 6; ++k) s -= L[k][r] * y[k];
        y[r] = s / L[r][r];
      }
#pragma unroll
      for (int k = 0; k < MAX_ARM; ++k) {
        if (k < na) {
          float u = 0.f;
#pragma unroll
          for (int r = 0; r < 6; ++r) u += J[r][k] * y[r];
          act[lo + k] = q[lo + k] + u;
        }
      }
      act[nd - 2] = q[nd - 2] + raw[6] * p.dt / 5;
      act[nd - 1] = q[nd - 1] + raw[6] * p.dt / 5;
    }
    for (int j = 0; j < nd; ++j)                   // tensor_clamp: max(min(t, upper), lower)
      p.action[(int64_t)e * nd + j] = fmaxf(fminf(act[j], p.dof_upper[j]), p.dof_lower[j]);
  }
  if (p.drive == 1 && p.jsum) {
    const float s = pm_block_sum(jpart, red);
    if (threadIdx.x == 0) atomicAdd(p.jsum, s);
  }
}

// hand_base.py:367-377; counts[0] = sum(success), counts[1] = sum(reset_buf), counts[2] = CTAs finished (all zero on entry)
__global__ void __launch_bounds__(256)
episode_flags_kernel(int E, int train, const float* __restrict__ rew, const int64_t* __restrict__ progress,
                     const uint8_t* __restrict__ success, float* __restrict__ epis_max_rew, int64_t* __restrict__ epis_max_step,
                     int64_t explore_step, int64_t max_episode_length, uint8_t* __restrict__ reset_buf, uint8_t* __restrict__ reset_succ,
                     int* __restrict__ counts, float* __restrict__ succ_rate) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  int s = 0, r = 0;
  if (e < E) {
    const int64_t prog = progress[e];
    if (train) {
      const float rw = rew[e], mr = epis_max_rew[e];
      const int64_t ms = rw < mr ? epis_max_step[e] : prog;
      epis_max_step[e] = ms;
      epis_max_rew[e] = fmaxf(rw, mr);
      s = success[e] ? 1 : 0;
      r = (prog >= ms + explore_step) || s;
      reset_succ[e] = (uint8_t)s;
    } else {
      r = prog >= max_episode_length;
    }
    reset_buf[e] = (uint8_t)r;
  }
  s = __syncthreads_count(s);                      // integer sums: order-independent, bit-exact
  r = __syncthreads_count(r);
  if (threadIdx.x == 0) {
    if (s) atomicAdd(counts + 0, s);
    if (r) atomicAdd(counts + 1, r);
    __threadfence();
    if (atomicAdd(counts + 2, 1) == (int)gridDim.x - 1) {
      const int ts = atomicAdd(counts + 0, 0), tr = atomicAdd(counts + 1, 0);
      if (succ_rate) *succ_rate = (float)ts / (float)max(tr, 1);
    }
  }
}

__global__ void scatter_targets_kernel(const float* __restrict__ pos_act, const int64_t* __restrict__ dof_mask, int mask_ld, int E,
                                       int nd, float* __restrict__ pos_act_all) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= E * nd) return;
  const int e = i / nd, j = i - e * nd;
  pos_act_all[dof_mask[(int64_t)e * mask_ld + j]] = pos_act[i];
}

}  // namespace

extern "C" {

int pm_open_drawer_obs_dim(int num_dofs) { return 29 + 2 * num_dofs; }

int pm_open_drawer_post_physics(const float* dof_state_all, const float* rigid_body_all, const float* root_tensor, int n_actors,
                                int obj_actor, const int64_t* dof_state_mask, const int64_t* rigid_body_mask, int E, int num_dofs,
                                int num_rigid_body, int ltip_rb_index, int rtip_rb_index, const float* dof_lower, const float* dof_upper,
                                const float* part_bbox_init, const float* part_axis_dir_init, const float* part_joint_lower,
                                const float* part_joint_upper, const int64_t* obj_lstid, float suc_prop, int do_obs, int do_reward,
                                int advance_progress, int64_t* progress_buf, float* obs, float* part_bbox, float* dof_state,
                                float* rigid_body, float* tip_rb, float* tip_rot_9d, float* gripper_length, float* dof_qpos_normalized,
                                float* rew_buf, uint8_t* success, uint8_t* succ_objid, float* extras_f, uint8_t* extras_b, pm_stream_t s) {
  PM_REQUIRE(dof_state_all && rigid_body_all && root_tensor && dof_state_mask && rigid_body_mask && part_bbox_init && part_axis_dir_init,
             PM_ERR_ARG, "pm_open_drawer_post_physics: null input");
  PM_REQUIRE(E > 0 && num_dofs > 2 && num_dofs <= MAX_DOFS && num_rigid_body > 0, PM_ERR_SHAPE,
             "pm_open_drawer_post_physics: E=%d num_dofs=%d num_rigid_body=%d", E, num_dofs, num_rigid_body);
  PM_REQUIRE(ltip_rb_index >= 0 && ltip_rb_index < num_rigid_body + 2 && rtip_rb_index >= 0 && rtip_rb_index < num_rigid_body + 2 &&
                 obj_actor >= 0 && obj_actor < n_actors,
             PM_ERR_SHAPE, "pm_open_drawer_post_physics: tip / actor index out of range");
  PM_REQUIRE(do_obs || do_reward, PM_ERR_ARG, "pm_open_drawer_post_physics: nothing to do");
  PM_REQUIRE((int64_t)E * ((num_rigid_body + 2) * 13 + (num_dofs + 1) * 2) < (1ll << 31), PM_ERR_SHAPE,
             "pm_open_drawer_post_physics: E=%d too large for 32-bit element indices", E);
  PM_REQUIRE(!do_obs || (dof_lower && dof_upper && obs && part_bbox && dof_state && rigid_body && tip_rb && tip_rot_9d && gripper_length &&
                         dof_qpos_normalized),
             PM_ERR_ARG, "pm_open_drawer_post_physics: null observation output");
  PM_REQUIRE(!do_reward || (part_joint_lower && part_joint_upper && obj_lstid && rew_buf && success && succ_objid && extras_f && extras_b),
             PM_ERR_ARG, "pm_open_drawer_post_physics: null reward input / output");
  PM_REQUIRE(!advance_progress || progress_buf, PM_ERR_ARG, "pm_open_drawer_post_physics: advance_progress without progress_buf");
  OpenDrawerP p;
  p.dof_all = dof_state_all; p.rb_all = rigid_body_all; p.root = root_tensor; p.dof_mask = dof_state_mask; p.rb_mask = rigid_body_mask;
  p.E = E; p.nd = num_dofs; p.nb = num_rigid_body; p.n_actors = n_actors; p.obj_actor = obj_actor; p.ltip = ltip_rb_index;
  p.rtip = rtip_rb_index; p.dof_lower = dof_lower; p.dof_upper = dof_upper; p.bbox_init = part_bbox_init; p.axis_dir = part_axis_dir_init;
  p.joint_lower = part_joint_lower; p.joint_upper = part_joint_upper; p.obj_lstid = obj_lstid; p.suc_prop = suc_prop;
  p.do_obs = do_obs; p.do_reward = do_reward; p.advance_progress = advance_progress; p.progress = progress_buf;
  p.obs = obs; p.part_bbox = part_bbox; p.dof_state = dof_state; p.rb_state = rigid_body; p.tip_rb = tip_rb; p.tip_rot = tip_rot_9d;
  p.gripper = gripper_length; p.qpos_norm = dof_qpos_normalized; p.rew = rew_buf; p.success = success; p.succ_objid = succ_objid;
  p.extras_f = extras_f; p.extras_b = extras_b;
  const size_t smem = do_obs ? (size_t)ENV_THREADS * (29 + 2 * num_dofs) * sizeof(float) : 0;
  const int n_env_ctas = pm_cdiv(E, ENV_THREADS);
  const int64_t n_copy = do_obs ? (int64_t)E * ((num_rigid_body + 2) * 13 + (num_dofs + 1) * 2) : 0;
  open_drawer_post_kernel<<<n_env_ctas + pm_cdiv(n_copy, ENV_THREADS * COPY_PER_THREAD), ENV_THREADS, smem, pm_st(s)>>>(p, n_env_ctas);
  PM_CHECK_LAUNCH("pm_open_drawer_post_physics");
  return PM_OK;
}

int pm_grasp_cube_obs_dim(int num_dofs) { return 19 + 2 * num_dofs; }

int pm_grasp_cube_post_physics(const float* dof_state, int dofs_per_env, const float* rigid_body, int bodies_per_env, const float* root_tensor,
                               int n_actors, int obj_actor, int E, int num_dofs, int ltip_rb_index, int rtip_rb_index, const float* dof_lower,
                               const float* dof_upper, const float* pose_lower_limit, const float* pose_upper_limit, const float* success_pos,
                               const float* obj_default_pos, float goal_thresh, int do_obs, int do_reward, int advance_progress,
                               int64_t* progress_buf, float* obs, float* proprio, float* tip_rb, float* tip_rot_9d, float* gripper_length,
                               float* dof_qpos_normalized, float* rew_buf, uint8_t* success, float* extras_f, uint8_t* extras_b, pm_stream_t s) {
  PM_REQUIRE(dof_state && rigid_body && root_tensor && pose_lower_limit && pose_upper_limit && success_pos && obj_default_pos, PM_ERR_ARG,
             "pm_grasp_cube_post_physics: null input");
  PM_REQUIRE(E > 0 && num_dofs > 2 && num_dofs <= MAX_DOFS && dofs_per_env >= num_dofs && bodies_per_env > 0, PM_ERR_SHAPE,
             "pm_grasp_cube_post_physics: E=%d num_dofs=%d dofs_per_env=%d bodies_per_env=%d", E, num_dofs, dofs_per_env, bodies_per_env);
  PM_REQUIRE(ltip_rb_index >= 0 && ltip_rb_index < bodies_per_env && rtip_rb_index >= 0 && rtip_rb_index < bodies_per_env && obj_actor >= 0 &&
                 obj_actor < n_actors,
             PM_ERR_SHAPE, "pm_grasp_cube_post_physics: tip / actor index out of range");
  PM_REQUIRE(do_obs || do_reward, PM_ERR_ARG, "pm_grasp_cube_post_physics: nothing to do");
  PM_REQUIRE(!do_obs || (dof_lower && dof_upper && obs && tip_rb && tip_rot_9d && gripper_length && dof_qpos_normalized), PM_ERR_ARG,
             "pm_grasp_cube_post_physics: null observation output");
  PM_REQUIRE(!do_reward || (rew_buf && success && extras_f && extras_b), PM_ERR_ARG, "pm_grasp_cube_post_physics: null reward output");
  PM_REQUIRE(!advance_progress || progress_buf, PM_ERR_ARG, "pm_grasp_cube_post_physics: advance_progress without progress_buf");
  GraspCubeP p;
  p.dof = dof_state; p.rb = rigid_body; p.root = root_tensor; p.E = E; p.nd = num_dofs; p.nd_all = dofs_per_env; p.nb = bodies_per_env;
  p.n_actors = n_actors; p.obj_actor = obj_actor; p.ltip = ltip_rb_index; p.rtip = rtip_rb_index; p.dof_lower = dof_lower; p.dof_upper = dof_upper;
  for (int i = 0; i < 7; ++i) { p.pose_lo[i] = pose_lower_limit[i]; p.pose_hi[i] = pose_upper_limit[i]; }
  for (int i = 0; i < 3; ++i) { p.success_pos[i] = success_pos[i]; p.obj_default[i] = obj_default_pos[i]; }
  p.goal_thresh = goal_thresh; p.do_obs = do_obs; p.do_reward = do_reward; p.advance_progress = advance_progress; p.progress = progress_buf;
  p.obs = obs; p.proprio = proprio; p.tip_rb = tip_rb; p.tip_rot = tip_rot_9d; p.gripper = gripper_length; p.qpos_norm = dof_qpos_normalized;
  p.rew = rew_buf; p.success = success; p.extras_f = extras_f; p.extras_b = extras_b;
  const size_t smem = do_obs ? (size_t)ENV_THREADS * (19 + 2 * num_dofs) * sizeof(float) : 0;
  grasp_cube_post_kernel<<<pm_cdiv(E, ENV_THREADS), ENV_THREADS, smem, pm_st(s)>>>(p);
  PM_CHECK_LAUNCH("pm_grasp_cube_post_physics");
  return PM_OK;
}

int pm_franka_control(const float* raw_output, int E, int num_dofs, int mobile, int drive_mode, const float* qpos, int64_t qpos_row_stride,
                      int64_t qpos_elem_stride, const int64_t* dof_state_mask, int mask_ld, const float* jacobian, int n_links, int ltip_rb_index,
                      int rtip_rb_index, const float* dof_lower, const float* dof_upper, const float* default_root_quat, float dt,
                      float damping, float* action_tensor, float* jacobian_sum, pm_stream_t s) {
  PM_REQUIRE(raw_output && qpos && dof_lower && dof_upper && action_tensor, PM_ERR_ARG, "pm_franka_control: null pointer");
  PM_REQUIRE(drive_mode == PM_DRIVE_POS || drive_mode == PM_DRIVE_IK, PM_ERR_ARG, "pm_franka_control: drive_mode %d (0 = pos, 1 = ik)",
             drive_mode);
  const int lo = mobile ? 3 : 0, na = num_dofs - 2 - lo;
  PM_REQUIRE(E > 0 && num_dofs <= MAX_DOFS && na >= 1 && na <= MAX_ARM && (!dof_state_mask || mask_ld >= num_dofs), PM_ERR_SHAPE,
             "pm_franka_control: E=%d num_dofs=%d arm dofs=%d mask_ld=%d", E, num_dofs, na, mask_ld);
  PM_REQUIRE(!mobile || default_root_quat, PM_ERR_ARG, "pm_franka_control: mobile base needs default_root_quat");
  FrankaP p;
  p.raw = raw_output; p.E = E; p.nd = num_dofs; p.mobile = mobile; p.drive = drive_mode; p.qpos = qpos;
  p.dof_mask = dof_state_mask; p.mask_ld = mask_ld; p.row_stride = qpos_row_stride; p.elem_stride = qpos_elem_stride; p.jac = jacobian; p.n_links = n_links; p.ltip = ltip_rb_index; p.rtip = rtip_rb_index;
  p.dof_lower = dof_lower; p.dof_upper = dof_upper; p.dt = dt; p.damping2 = damping * damping; p.action = action_tensor;
  p.jsum = jacobian_sum;
  for (int i = 0; i < 9; ++i) p.root_rt[i] = i % 4 == 0 ? 1.f : 0.f;
  if (mobile) {                                  // quat_to_mat(default_root[3:7]) (host pointer: 4 floats of configuration)
    const float i = default_root_quat[0], j = default_root_quat[1], k = default_root_quat[2], r = default_root_quat[3];
    const float two_s = 2.0f / (i * i + j * j + k * k + r * r);
    const float m[9] = {1 - two_s * (j * j + k * k), two_s * (i * j - k * r),     two_s * (i * k + j * r),
                        two_s * (i * j + k * r),     1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                        two_s * (i * k - j * r),     two_s * (j * k + i * r),     1 - two_s * (i * i + j * j)};
    for (int q = 0; q < 9; ++q) p.root_rt[q] = m[q];
  }
  if (drive_mode == PM_DRIVE_IK) {
    PM_REQUIRE(jacobian && n_links > 0 && ltip_rb_index >= 1 && ltip_rb_index - 1 < n_links && rtip_rb_index >= 1 && rtip_rb_index - 1 < n_links,
               PM_ERR_SHAPE, "pm_franka_control: ik needs the Jacobian and tip links inside it");
    if (jacobian_sum) {
      cudaError_t e = cudaMemsetAsync(jacobian_sum, 0, sizeof(float), pm_st(s));
      if (e != cudaSuccess) PM_FAIL(PM_ERR_CUDA, "pm_franka_control: %s", cudaGetErrorString(e));
    }
  }
  franka_control_kernel<<<pm_cdiv(E, ENV_THREADS), ENV_THREADS, 0, pm_st(s)>>>(p);
  PM_CHECK_LAUNCH("pm_franka_control");
  return PM_OK;
}

int pm_episode_flags(int E, int train, const float* rew_buf, const int64_t* progress_buf, const uint8_t* success, float* epis_max_rew,
                     int64_t* epis_max_step, int64_t explore_step, int64_t max_episode_length, uint8_t* reset_buf, uint8_t* reset_succ,
                     int32_t* counts3, float* succ_rate, pm_stream_t s) {
  PM_REQUIRE(E > 0 && progress_buf && reset_buf && counts3, PM_ERR_ARG, "pm_episode_flags: bad args");
  PM_REQUIRE(!train || (rew_buf && success && epis_max_rew && epis_max_step && reset_succ && succ_rate), PM_ERR_ARG,
             "pm_episode_flags: train mode needs rew_buf, success, epis_max_*, reset_succ, succ_rate");
  cudaError_t e = cudaMemsetAsync(counts3, 0, 3 * sizeof(int32_t), pm_st(s));
  if (e != cudaSuccess) PM_FAIL(PM_ERR_CUDA, "pm_episode_flags: %s", cudaGetErrorString(e));
  episode_flags_kernel<<<pm_cdiv(E, 256), 256, 0, pm_st(s)>>>(E, train, rew_buf, progress_buf, success, epis_max_rew, epis_max_step,
                                                               explore_step, max_episode_length, reset_buf, reset_succ, counts3, succ_rate);
  PM_CHECK_LAUNCH("pm_episode_flags");
  return PM_OK;
}

int pm_scatter_dof_targets(const float* pos_act, const int64_t* dof_state_mask, int mask_ld, int E, int num_dofs, float* pos_act_all,
                           pm_stream_t s) {
  PM_REQUIRE(pos_act && dof_state_mask && pos_act_all && E > 0 && num_dofs > 0 && mask_ld >= num_dofs, PM_ERR_ARG,
             "pm_scatter_dof_targets: bad args");
  scatter_targets_kernel<<<pm_cdiv((long long)E * num_dofs, 256), 256, 0, pm_st(s)>>>(pos_act, dof_state_mask, mask_ld, E, num_dofs,
                                                                                       pos_act_all);
  PM_CHECK_LAUNCH("pm_scatter_dof_targets");
  return PM_OK;
}

}  // extern "C"
