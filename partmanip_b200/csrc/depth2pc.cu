// NEXT ROW — depth images -> world-frame point cloud -> farthest-point subsample: the observation producer
// immediately upstream of the encoder.  reference: utils/depth2tsdf.py:136-173 (TSDFVolume.depth2pc); the FPS step is
// pytorch3d.ops.sample_farthest_points there (not vendored; algorithm restated in oracle/depth2pc_oracle.py).
//
//   backproject_kernel  one thread per pixel: ((col - cx) * d / fx, (row - cy) * d / fy, d) -> R . p + t -> zeroed unless strictly
//                       inside the workspace box.  Streaming, HBM-bound: 4 B read + 12 B written per pixel.
//   fps_kernel          one CTA per cloud, K greedy picks: every thread owns a strided slice of the points, keeps the running
//                       min squared distance to the selected set (shared memory when the cloud fits, else global) and its
//                       local arg-max; a block arg-max (first index on ties) selects the next point.  Squared distances are
//                       (dx*dx + dy*dy) + dz*dz with explicit round-to-nearest mul/add (no FMA contraction) so picks match
//                       the oracle bit for bit.  First correct version: cost is O(K * P) per cloud; compacting the valid
//                       points (the zeroed ones are all duplicates of (0,0,0)) is the next step.
#include "common.cuh"
#include "tc_common.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace {

__global__ void __launch_bounds__(256)
backproject_kernel(const float* __restrict__ depth, int E, int M, int HW, int W, float cx, float cy, float fx, float fy,
                   const float* __restrict__ cam_pose /* (M,4,4) row-major */, float ox, float oy, float oz, float size,
                   float* __restrict__ out /* (E, M*HW, 3) */) {
  const int64_t total = (int64_t)E * M * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const int m = (int)((i / HW) % M);
    const float d = __ldg(depth + i);
    const float col = (float)(p % W), row = (float)(p / W);
    const float p0 = __fdiv_rn(__fmul_rn(col - cx, d), fx);
    const float p1 = __fdiv_rn(__fmul_rn(row - cy, d), fy);
    const float* T = cam_pose + m * 16;
    float w[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) w[a] = fmaf(d, T[a * 4 + 2], fmaf(p1, T[a * 4 + 1], p0 * T[a * 4 + 0])) + T[a * 4 + 3];
    const bool valid = w[0] < size + ox && w[1] < size + oy && w[2] < size + oz && w[0] > ox && w[1] > oy && w[2] > oz;
    float* o = out + i * 3;
    o[0] = valid ? w[0] : 0.f; o[1] = valid ? w[1] : 0.f; o[2] = valid ? w[2] : 0.f;
  }
}

// 4 consecutive pixels per thread (HW % 4 == 0): one 16-byte depth load, three 16-byte stores (48 contiguous bytes)
__global__ void __launch_bounds__(256)
backproject4_kernel(const float* __restrict__ depth, int E, int M, int HW, int W, float cx, float cy, float fx, float fy,
                    const float* __restrict__ cam_pose, float ox, float oy, float oz, float size, float* __restrict__ out) {
  const int64_t total4 = (int64_t)E * M * HW / 4;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total4; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i0 = q * 4;
    const int p0i = (int)(i0 % HW);
    const int m = (int)((i0 / HW) % M);
    const float4 d4 = __ldg(reinterpret_cast<const float4*>(depth + i0));
    const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
    const float* T = cam_pose + m * 16;
    float o[12];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int p = p0i + j;
      const float d = dd[j];
      const float col = (float)(p % W), row = (float)(p / W);
      const float a0 = __fdiv_rn(__fmul_rn(col - cx, d), fx);
      const float a1 = __fdiv_rn(__fmul_rn(row - cy, d), fy);
      float w[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) w[a] = fmaf(d, T[a * 4 + 2], fmaf(a1, T[a * 4 + 1], a0 * T[a * 4 + 0])) + T[a * 4 + 3];
      const bool valid = w[0] < size + ox && w[1] < size + oy && w[2] < size + oz && w[0] > ox && w[1] > oy && w[2] > oz;
      o[3 * j] = valid ? w[0] : 0.f; o[3 * j + 1] = valid ? w[1] : 0.f; o[3 * j + 2] = valid ? w[2] : 0.f;
    }
    float4* dst = reinterpret_cast<float4*>(out + i0 * 3);
    dst[0] = make_float4(o[0], o[1], o[2], o[3]);
    dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    dst[2] = make_float4(o[8], o[9], o[10], o[11]);
  }
}

// The same back-projection reading the E*M camera images where the simulator left them (a device table of pointers), with the
// reference's stacking step folded in: d = -image (negate), +-inf -> inf_value (tasks/hand_base.py:317-324).  The stacked
// (E,M,H,W) tensor and the negated / isinf / where temporaries never exist.
template <bool VEC4>
__global__ void __launch_bounds__(256)
backproject_views_kernel(const float* const* __restrict__ views /* E*M pointers to (H,W) images */, int E, int M, int HW, int W, int negate,
                         float inf_value, float cx, float cy, float fx, float fy, const float* __restrict__ cam_pose, float ox, float oy,
                         float oz, float size, float* __restrict__ out) {
  constexpr int V = VEC4 ? 4 : 1;
  const int64_t totalv = (int64_t)E * M * HW / V;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < totalv; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i0 = q * V;
    const int p0i = (int)(i0 % HW);
    const int view = (int)(i0 / HW);
    const int m = view % M;
    const float* img = views[view];
    float dd[V];
    if (VEC4) {
      const float4 d4 = __ldg(reinterpret_cast<const float4*>(img + p0i));
      dd[0] = d4.x; dd[V > 1 ? 1 : 0] = d4.y; dd[V > 2 ? 2 : 0] = d4.z; dd[V > 3 ? 3 : 0] = d4.w;
    } else {
      dd[0] = __ldg(img + p0i);
    }
    const float* T = cam_pose + m * 16;
    float o[3 * V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int p = p0i + j;
      float d = negate ? -dd[j] : dd[j];
      if (isinf(d)) d = inf_value;
      const float col = (float)(p % W), row = (float)(p / W);
      const float a0 = __fdiv_rn(__fmul_rn(col - cx, d), fx);
      const float a1 = __fdiv_rn(__fmul_rn(row - cy, d), fy);
      float w[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) w[a] = fmaf(d, T[a * 4 + 2], fmaf(a1, T[a * 4 + 1], a0 * T[a * 4 + 0])) + T[a * 4 + 3];
      const bool valid = w[0] < size + ox && w[1] < size + oy && w[2] < size + oz && w[0] > ox && w[1] > oy && w[2] > oz;
      o[3 * j] = valid ? w[0] : 0.f; o[3 * j + 1] = valid ? w[1] : 0.f; o[3 * j + 2] = valid ? w[2] : 0.f;
    }
    if (VEC4) {
      float4* dst = reinterpret_cast<float4*>(out + i0 * 3);
      dst[0] = make_float4(o[0], o[1], o[2], o[V > 1 ? 3 : 0]);
      dst[1] = make_float4(o[V > 1 ? 4 : 0], o[V > 1 ? 5 : 0], o[V > 2 ? 6 : 0], o[V > 2 ? 7 : 0]);
      dst[2] = make_float4(o[V > 2 ? 8 : 0], o[V > 3 ? 9 : 0], o[V > 3 ? 10 : 0], o[V > 3 ? 11 : 0]);
    } else {
      out[i0 * 3] = o[0]; out[i0 * 3 + 1] = o[1]; out[i0 * 3 + 2] = o[2];
    }
  }
}

constexpr int FPS_THREADS = 1024;
constexpr int FPS_SMEM_POINTS = 48 * 1024;   // running min-distances kept in shared memory up to this many points (192 KB)

__device__ __forceinline__ float sqdist(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = ax - bx, dy = ay - by, dz = az - bz;
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

template <bool SMEM_DIST, bool VEC4>
__global__ void __launch_bounds__(FPS_THREADS, 1)
fps_kernel(const float* __restrict__ pts /* (E,P,3) */, int P, int K, float* __restrict__ mind_g /* (E,P) scratch */,
           float* __restrict__ out /* (E,K,3) */, int64_t* __restrict__ out_idx /* (E,K) or null */) {
  extern __shared__ float mind_s[];
  __shared__ float red_v[32];
  __shared__ int red_i[32];
  __shared__ int s_last;
  const int e = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* p = pts + (int64_t)e * P * 3;
  float* mind = SMEM_DIST ? mind_s : mind_g + (int64_t)e * P;
  for (int i = tid; i < P; i += FPS_THREADS) mind[i] = 3.402823466e+38f;
  int last = 0;
  if (tid == 0) {
    out[(int64_t)e * K * 3 + 0] = p[0]; out[(int64_t)e * K * 3 + 1] = p[1]; out[(int64_t)e * K * 3 + 2] = p[2];
    if (out_idx) out_idx[(int64_t)e * K] = 0;
  }
  __syncthreads();
  for (int k = 1; k < K; ++k) {
    const float lx = __ldg(p + (int64_t)last * 3), ly = __ldg(p + (int64_t)last * 3 + 1), lz = __ldg(p + (int64_t)last * 3 + 2);
    float bv = -1.f;
    int bi = 0x7fffffff;
    if (VEC4) {                                               // 4 points per step: three 16-byte loads, one 16-byte min-distance update
      for (int g4 = tid; g4 < P / 4; g4 += FPS_THREADS) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p) + 3 * g4), b = __ldg(reinterpret_cast<const float4*>(p) + 3 * g4 + 1),
                     c = __ldg(reinterpret_cast<const float4*>(p) + 3 * g4 + 2);
        float4 m4 = reinterpret_cast<float4*>(mind)[g4];
        m4.x = fminf(m4.x, sqdist(a.x, a.y, a.z, lx, ly, lz));
        m4.y = fminf(m4.y, sqdist(a.w, b.x, b.y, lx, ly, lz));
        m4.z = fminf(m4.z, sqdist(b.z, b.w, c.x, lx, ly, lz));
        m4.w = fminf(m4.w, sqdist(c.y, c.z, c.w, lx, ly, lz));
        reinterpret_cast<float4*>(mind)[g4] = m4;
        if (m4.x > bv) { bv = m4.x; bi = 4 * g4; }            // ascending index: strict > keeps the first index
        if (m4.y > bv) { bv = m4.y; bi = 4 * g4 + 1; }
        if (m4.z > bv) { bv = m4.z; bi = 4 * g4 + 2; }
        if (m4.w > bv) { bv = m4.w; bi = 4 * g4 + 3; }
      }
    } else {
      for (int i = tid; i < P; i += FPS_THREADS) {
        const float d = sqdist(__ldg(p + (int64_t)i * 3), __ldg(p + (int64_t)i * 3 + 1), __ldg(p + (int64_t)i * 3 + 2), lx, ly, lz);
        const float m = fminf(mind[i], d);
        mind[i] = m;
        if (m > bv) { bv = m; bi = i; }                       // ascending i: strict > keeps the first index
      }
    }
    // block arg-max, first index on ties
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      bv = red_v[lane]; bi = red_i[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) {
        s_last = bi;
        float* o3 = out + ((int64_t)e * K + k) * 3;
        o3[0] = p[(int64_t)bi * 3]; o3[1] = p[(int64_t)bi * 3 + 1]; o3[2] = p[(int64_t)bi * 3 + 2];
        if (out_idx) out_idx[(int64_t)e * K + k] = bi;
      }
    }
    __syncthreads();
    last = s_last;
  }
}

// ---- order-preserving compaction: keep every non-zero point and the FIRST zero point.  FPS is invariant under removing exact
//      duplicates as long as first occurrences stay in order (a duplicate of a selected point has min-distance 0, a duplicate
//      of an unselected one ties with it and the first index wins), and the masked points are all duplicates of (0,0,0).
//      The compacted cloud is padded to a multiple of 4 with copies of its point 0 (never selected before a distinct point).
__global__ void __launch_bounds__(FPS_THREADS, 1)
fps_compact_kernel(const float* __restrict__ pts, int P, float* __restrict__ cp /* (E, P+4, 3) */, int32_t* __restrict__ om /* (E, P+4) */,
                   int32_t* __restrict__ counts /* (E) */) {
  __shared__ int s_warp[32];
  __shared__ int s_carry, s_z0;
  const int e = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* p = pts + (int64_t)e * P * 3;
  float* c = cp + (int64_t)e * (P + 4) * 3;
  int32_t* o = om + (int64_t)e * (P + 4);
  if (tid == 0) { s_carry = 0; s_z0 = 0x7fffffff; }
  __syncthreads();
  int z0 = 0x7fffffff;
  for (int i = tid; i < P; i += FPS_THREADS)
    if (z0 == 0x7fffffff && p[(int64_t)i * 3] == 0.f && p[(int64_t)i * 3 + 1] == 0.f && p[(int64_t)i * 3 + 2] == 0.f) z0 = i;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) z0 = min(z0, __shfl_xor_sync(0xffffffffu, z0, s));
  if (lane == 0 && z0 != 0x7fffffff) atomicMin(&s_z0, z0);
  __syncthreads();
  z0 = s_z0;
  for (int base = 0; base < P; base += FPS_THREADS) {
    const int i = base + tid;
    float x = 0.f, y = 0.f, z = 0.f;
    bool keep = false;
    if (i < P) {
      x = p[(int64_t)i * 3]; y = p[(int64_t)i * 3 + 1]; z = p[(int64_t)i * 3 + 2];
      keep = (x != 0.f || y != 0.f || z != 0.f) || i == z0;
    }
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[warp] = __popc(mask);
    __syncthreads();
    int woff = 0, total = 0;
    for (int w = 0; w < 32; ++w) { const int n = s_warp[w]; if (w < warp) woff += n; total += n; }
    const int carry = s_carry;
    if (keep) {
      const int pos = carry + woff + __popc(mask & ((1u << lane) - 1u));
      c[(int64_t)pos * 3] = x; c[(int64_t)pos * 3 + 1] = y; c[(int64_t)pos * 3 + 2] = z;
      o[pos] = i;
    }
    __syncthreads();
    if (tid == 0) s_carry = carry + total;
    __syncthreads();
  }
  if (tid == 0) {
    int n = s_carry;
    counts[e] = n;
    while (n & 3) { c[(int64_t)n * 3] = c[0]; c[(int64_t)n * 3 + 1] = c[1]; c[(int64_t)n * 3 + 2] = c[2]; o[n] = o[0]; ++n; }
  }
}

// FPS over the compacted clouds (per-cloud point count in device memory; min distances in shared memory when they fit)
__global__ void __launch_bounds__(FPS_THREADS, 1)
fps_compacted_kernel(const float* __restrict__ cp, const int32_t* __restrict__ om, const int32_t* __restrict__ counts, int P, int K,
                     float* __restrict__ mind_g /* (E, P+4) */, float* __restrict__ out, int64_t* __restrict__ out_idx) {
  extern __shared__ __align__(16) float mind_s[];
  __shared__ float red_v[32];
  __shared__ int red_i[32];
  __shared__ int s_last;
  const int e = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* p = cp + (int64_t)e * (P + 4) * 3;
  const int32_t* o = om + (int64_t)e * (P + 4);
  const int n4 = (counts[e] + 3) >> 2;                                  // groups of 4 points (padded)
  float* mind = (n4 * 4 <= FPS_SMEM_POINTS) ? mind_s : mind_g + (int64_t)e * (P + 4);
  for (int i = tid; i < n4 * 4; i += FPS_THREADS) mind[i] = 3.402823466e+38f;
  int last = 0;
  if (tid == 0) {
    out[(int64_t)e * K * 3 + 0] = p[0]; out[(int64_t)e * K * 3 + 1] = p[1]; out[(int64_t)e * K * 3 + 2] = p[2];
    if (out_idx) out_idx[(int64_t)e * K] = o[0];
  }
  __syncthreads();
  for (int k = 1; k < K; ++k) {
    const float lx = p[(int64_t)last * 3], ly = p[(int64_t)last * 3 + 1], lz = p[(int64_t)last * 3 + 2];
    float bv = -1.f;
    int bi = 0x7fffffff;
    for (int g4 = tid; g4 < n4; g4 += FPS_THREADS) {
      const float4 a = reinterpret_cast<const float4*>(p)[3 * g4], b = reinterpret_cast<const float4*>(p)[3 * g4 + 1],
                   c = reinterpret_cast<const float4*>(p)[3 * g4 + 2];
      float4 m4 = reinterpret_cast<float4*>(mind)[g4];
      m4.x = fminf(m4.x, sqdist(a.x, a.y, a.z, lx, ly, lz));
      m4.y = fminf(m4.y, sqdist(a.w, b.x, b.y, lx, ly, lz));
      m4.z = fminf(m4.z, sqdist(b.z, b.w, c.x, lx, ly, lz));
      m4.w = fminf(m4.w, sqdist(c.y, c.z, c.w, lx, ly, lz));
      reinterpret_cast<float4*>(mind)[g4] = m4;
      if (m4.x > bv) { bv = m4.x; bi = 4 * g4; }
      if (m4.y > bv) { bv = m4.y; bi = 4 * g4 + 1; }
      if (m4.z > bv) { bv = m4.z; bi = 4 * g4 + 2; }
      if (m4.w > bv) { bv = m4.w; bi = 4 * g4 + 3; }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, s);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, s);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      bv = red_v[lane]; bi = red_i[lane];
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, s);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, s);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) {
        if (bi == 0x7fffffff) bi = 0;
        s_last = bi;
        float* o3 = out + ((int64_t)e * K + k) * 3;
        o3[0] = p[(int64_t)bi * 3]; o3[1] = p[(int64_t)bi * 3 + 1]; o3[2] = p[(int64_t)bi * 3 + 2];
        if (out_idx) out_idx[(int64_t)e * K + k] = o[bi];
      }
    }
    __syncthreads();
    last = s_last;
  }
}

// ---- cluster version for large clouds: EIGHT CTAs (one thread-block cluster) per cloud, each owning a contiguous slice of the
//      compacted points.  The running min distances live in REGISTERS (8 groups of 4 points per thread = 32 Ki points per CTA,
//      256 Ki per cloud; anything beyond spills to the global scratch), the first 18 432 points of the slice in the CTA's shared
//      memory (216 KB), the rest streams from L2 (16 concurrent clouds' slices stay L2-resident) one group ahead of its use.
//      Per pick: every CTA reduces its slice's arg-max and pushes {coords, value, index} into all eight CTAs' shared memory with
//      st.async (DSMEM store that completes bytes on the destination's mbarrier); every warp waits on its own CTA's mbarrier
//      and re-reduces the eight candidates.  No cluster barrier and no fence in the loop; slots and mbarriers are double-
//      buffered by pick parity.  Picks are identical to fps_compacted_kernel (same arithmetic, first index on ties).
constexpr int FPS_CL = 8;                        // CTAs per cloud (portable cluster size)
constexpr int FPS_CL_THREADS = 512;              // 128 registers per thread: 64 of them hold min distances
constexpr int FPS_CL_G = 16;                     // register-resident groups of 4 points per thread (8192 groups per CTA)
constexpr int FPS_CL_S = 9;                      // groups 0..8 of a thread are read from shared memory, 9..15 from L2
constexpr int FPS_CL_NS4 = FPS_CL_S * FPS_CL_THREADS;   // 4608 groups of the slice held in shared memory (x 48 B = 216 KB)
constexpr uint32_t FPS_CL_TX = FPS_CL * 32;      // bytes pushed into every CTA per pick

#ifdef PM_FPS_TIMING   // debug build: clock64 stamps of pick 300, cloud 0, threads 0 and 480 of every rank, into the min-distance scratch
#define FPS_STAMP(i) if (e == 0 && k == 300 && (tid == 0 || tid == 480)) reinterpret_cast<long long*>(mind_g)[(rank * 2 + (tid != 0)) * 8 + (i)] = clock64();
#else
#define FPS_STAMP(i)
#endif

#define FPS_UPD(M, X, Y, Z, BV, BC, CODE)                                 \
  {                                                                       \
    const float m_ = fminf(M, sqdist(X, Y, Z, lx, ly, lz));               \
    M = m_;                                                               \
    if (m_ > BV) { BV = m_; BC = (CODE); }                                \
  }
#define FPS_UPD4(MD, A, B, C, BV, BC, CODE0)                              \
  FPS_UPD(MD.x, A.x, A.y, A.z, BV, BC, (CODE0))                           \
  FPS_UPD(MD.y, A.w, B.x, B.y, BV, BC, (CODE0) + 1)                       \
  FPS_UPD(MD.z, B.z, B.w, C.x, BV, BC, (CODE0) + 2)                       \
  FPS_UPD(MD.w, C.y, C.z, C.w, BV, BC, (CODE0) + 3)

__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, float a, float b, float c, float d, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               :: "r"(remote_addr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d)),
                  "r"(remote_bar) : "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}

__global__ void __cluster_dims__(FPS_CL, 1, 1) __launch_bounds__(FPS_CL_THREADS, 1)
fps_cluster_kernel(const float* __restrict__ cp, const int32_t* __restrict__ om, const int32_t* __restrict__ counts, int P, int K,
                   float* __restrict__ mind_g /* (E, P+4) */, float* __restrict__ out, int64_t* __restrict__ out_idx) {
  extern __shared__ __align__(16) float4 pts_s[];                       // [FPS_CL_NS4][3]
  __shared__ unsigned red_v[FPS_CL_THREADS / 32];
  __shared__ int red_i[FPS_CL_THREADS / 32];
  __shared__ __align__(16) float4 slot[2][FPS_CL][2];                   // [parity][source rank]{(x,y,z,value bits),(index,-,-,-)}
  __shared__ __align__(8) uint64_t mbar[2];                             // [parity]: 256 bytes of candidates per phase
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int e = blockIdx.x / FPS_CL, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* pe = cp + (int64_t)e * (P + 4) * 3;
  const int n4 = (counts[e] + 3) >> 2;                                  // groups of 4 points in the cloud (padded)
  const int n4c = (n4 + FPS_CL - 1) / FPS_CL;
  const int g0 = rank * n4c;                                            // first group of this CTA's slice
  const int ng = max(0, min(n4c, n4 - g0));
  const float4* gp = reinterpret_cast<const float4*>(pe) + 3 * (int64_t)g0;
  float4* mo = reinterpret_cast<float4*>(mind_g + (int64_t)e * (P + 4)) + g0;   // overflow min distances (groups >= 8192)
  if (tid == 0) {
    pmtc::mbar_init(pmtc::smem_u32(&mbar[0]), 1);
    pmtc::mbar_init(pmtc::smem_u32(&mbar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 3 * min(ng, FPS_CL_NS4); i += FPS_CL_THREADS) pts_s[i] = gp[i];
  float4 md[FPS_CL_G];
#pragma unroll
  for (int j = 0; j < FPS_CL_G; ++j) md[j] = make_float4(3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f);
  for (int lg = FPS_CL_G * FPS_CL_THREADS + tid; lg < ng; lg += FPS_CL_THREADS)
    mo[lg] = make_float4(3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f);
  float lx = pe[0], ly = pe[1], lz = pe[2];
  if (rank == 0 && tid == 32) {
    out[(int64_t)e * K * 3 + 0] = lx; out[(int64_t)e * K * 3 + 1] = ly; out[(int64_t)e * K * 3 + 2] = lz;
    if (out_idx) out_idx[(int64_t)e * K] = 0;                           // indices into the compacted cloud; translated at the end
  }
  cluster.sync();                                                       // every CTA's mbarriers are initialised before any push
  if (tid == 0 && K > 1)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(pmtc::smem_u32(&mbar[1])), "r"(FPS_CL_TX) : "memory");
  for (int k = 1; k < K; ++k) {
    const int par = k & 1;
    FPS_STAMP(0)
    // two trackers, each fed in ascending index order: groups 0..8 (shared memory) and groups 9.. (L2, fetched two groups
    // ahead of their use into q0 / q1)
    float bv = 0.f, bvg = 0.f;                                          // strict >: a slice whose distances are all 0 offers nothing
    int bc = -1, bcg = -1, bo = 0;
    float4 q0a, q0b, q0c, q1a, q1b, q1c;
#define FPS_FETCH(J, QA, QB, QC)                                                                          \
    {                                                                                                     \
      const int lg_ = (J) * FPS_CL_THREADS + tid;                                                         \
      if (lg_ < ng) { QA = __ldg(gp + 3 * lg_); QB = __ldg(gp + 3 * lg_ + 1); QC = __ldg(gp + 3 * lg_ + 2); } \
    }
#define FPS_SMEM_GROUP(J)                                                                                 \
    {                                                                                                     \
      const int lg_ = (J) * FPS_CL_THREADS + tid;                                                         \
      if (lg_ < ng) {                                                                                     \
        const float4 a_ = pts_s[3 * lg_], b_ = pts_s[3 * lg_ + 1], c_ = pts_s[3 * lg_ + 2];               \
        FPS_UPD4(md[J], a_, b_, c_, bv, bc, 4 * (J))                                                      \
      }                                                                                                   \
    }
    FPS_FETCH(FPS_CL_S, q0a, q0b, q0c)
    FPS_FETCH(FPS_CL_S + 1, q1a, q1b, q1c)
    FPS_SMEM_GROUP(0)
    FPS_SMEM_GROUP(1)
#pragma unroll
    for (int i = 0; i < FPS_CL_G - FPS_CL_S; ++i) {
      const int jg = FPS_CL_S + i;
      if (i & 1) {
        if (jg * FPS_CL_THREADS + tid < ng) { FPS_UPD4(md[jg], q1a, q1b, q1c, bvg, bcg, 4 * jg) }
        if (jg + 2 < FPS_CL_G) FPS_FETCH(jg + 2, q1a, q1b, q1c)
      } else {
        if (jg * FPS_CL_THREADS + tid < ng) { FPS_UPD4(md[jg], q0a, q0b, q0c, bvg, bcg, 4 * jg) }
        if (jg + 2 < FPS_CL_G) FPS_FETCH(jg + 2, q0a, q0b, q0c)
      }
      FPS_SMEM_GROUP(2 + i)
    }
#undef FPS_FETCH
#undef FPS_SMEM_GROUP
    for (int lg = FPS_CL_G * FPS_CL_THREADS + tid; lg < ng; lg += FPS_CL_THREADS) {   // beyond the register-resident part
      const float4 a = __ldg(gp + 3 * lg), b = __ldg(gp + 3 * lg + 1), c = __ldg(gp + 3 * lg + 2);
      float4 m4 = mo[lg];
      const int bc0 = bcg;
      bcg = -2;
      FPS_UPD4(m4, a, b, c, bvg, bcg, 0)
      mo[lg] = m4;
      if (bcg >= 0) { bo = 4 * lg + bcg; bcg = 1 << 20; } else bcg = bc0;
    }
    if (bvg > bv) { bv = bvg; bc = bcg; }                               // every index of the second tracker is above the first's
    // local point index -> index in the compacted cloud
    int bi = 0x7fffffff;
    if (bc >= 0) bi = 4 * g0 + ((bc & (1 << 20)) ? bo : 4 * ((bc >> 2) * FPS_CL_THREADS + tid) + (bc & 3));
    FPS_STAMP(1)
    unsigned vb = __float_as_uint(bv);                                  // distances are >= 0: their bit patterns order like the values
    unsigned vmax = __reduce_max_sync(0xffffffffu, vb);
    int imin = __reduce_min_sync(0xffffffffu, vb == vmax ? bi : 0x7fffffff);
    if (lane == 0) { red_v[warp] = vmax; red_i[warp] = imin; }
    FPS_STAMP(2)
    __syncthreads();
    FPS_STAMP(3)
    if (warp == 0) {
      // everyone is past the previous pick's wait: arm the other mbarrier for the next pick
      if (lane == 0 && k + 1 < K)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(pmtc::smem_u32(&mbar[par ^ 1])), "r"(FPS_CL_TX) : "memory");
      vb = lane < FPS_CL_THREADS / 32 ? red_v[lane] : 0u; bi = lane < FPS_CL_THREADS / 32 ? red_i[lane] : 0x7fffffff;
      vmax = __reduce_max_sync(0xffffffffu, vb);
      imin = __reduce_min_sync(0xffffffffu, vb == vmax ? bi : 0x7fffffff);
      float x = 0.f, y = 0.f, z = 0.f;
      if (imin != 0x7fffffff) {                                         // the CTA's candidate (uniform address: broadcast load)
        const int lp = imin - 4 * g0;
        if ((lp >> 2) < FPS_CL_NS4) {
          const float* ps = reinterpret_cast<const float*>(pts_s) + 3 * lp;
          x = ps[0]; y = ps[1]; z = ps[2];
        } else {
          x = pe[(int64_t)imin * 3]; y = pe[(int64_t)imin * 3 + 1]; z = pe[(int64_t)imin * 3 + 2];
        }
      }
      if (lane < FPS_CL) {                                              // lane r pushes the candidate into CTA r's slot
        const uint32_t rs = mapa_u32(pmtc::smem_u32(&slot[par][rank][0]), lane), rb = mapa_u32(pmtc::smem_u32(&mbar[par]), lane);
        st_async_v4(rs, x, y, z, __uint_as_float(vmax), rb);
        st_async_v4(rs + 16, __int_as_float(imin), 0.f, 0.f, 0.f, rb);
      }
    }
    FPS_STAMP(4)
    {                                                                   // all eight candidates have landed in this CTA's slots
      const uint32_t bar = pmtc::smem_u32(&mbar[par]), ph = ((k - 1) >> 1) & 1;
      uint32_t spin = 0;
      while (!pmtc::mbar_try_wait(bar, ph))
        if (++spin > (1u << 22)) __trap();                              // a lost push must not hang the GPU
    }
    FPS_STAMP(5)
    {
      const float4 s0 = slot[par][lane & (FPS_CL - 1)][0];
      const int si = __float_as_int(slot[par][lane & (FPS_CL - 1)][1].x);
      const unsigned sv = __float_as_uint(s0.w);
      vmax = __reduce_max_sync(0xffffffffu, sv);
      imin = __reduce_min_sync(0xffffffffu, sv == vmax ? si : 0x7fffffff);
      const int src = __ffs(__ballot_sync(0xffffffffu, sv == vmax && si == imin)) - 1;
      lx = __shfl_sync(0xffffffffu, s0.x, src); ly = __shfl_sync(0xffffffffu, s0.y, src); lz = __shfl_sync(0xffffffffu, s0.z, src);
      if (imin == 0x7fffffff) { imin = 0; lx = pe[0]; ly = pe[1]; lz = pe[2]; }    // every remaining point coincides with a pick
    }
    FPS_STAMP(6)
    if (rank == 0 && tid == 32) {                                       // fire-and-forget stores from a warp that never fences
      float* o3 = out + ((int64_t)e * K + k) * 3;
      o3[0] = lx; o3[1] = ly; o3[2] = lz;
      if (out_idx) out_idx[(int64_t)e * K + k] = imin;
    }
  }
  if (rank == 0 && out_idx) {                                           // compacted index -> index in the caller's cloud
    __syncthreads();
    const int32_t* o = om + (int64_t)e * (P + 4);
    for (int k = tid; k < K; k += FPS_CL_THREADS) out_idx[(int64_t)e * K + k] = o[out_idx[(int64_t)e * K + k]];
  }
  cluster.sync();                                                       // no CTA leaves while a peer could still address its memory
}
#undef FPS_UPD4
#undef FPS_UPD

// ---- pruned cluster version (exact): the same 8-CTA layout, but every warp keeps, for each of its 16 blocks of 128 consecutive
//      points (32 lanes x 4 points — a stretch of one depth-image row, so spatially compact), the block's bounding box, its
//      largest min distance and that point's position + coordinates (lane j of the warp owns block j's summary).  A new pick
//      can only lower a block's distances if the box is closer to it than the block's largest min distance; the lower bound is
//      computed with the same rounding sequence as the distances themselves (fp32 subtraction, multiplication and addition are
//      monotone), so skipping is exact — the picks stay bit-identical to the other kernels.  One 16-lane box test replaces the
//      whole scan; only the blocks near the pick are updated (a few per cent once a few dozen picks have been made), and the
//      arg-max comes from the block summaries without touching the points.  Consecutive blocks go to consecutive CTAs of the
//      cluster (block B -> CTA B % 8, warp (B / 8) % 16), because the blocks a pick reaches are neighbours in the image: the
//      work of one pick spreads over all 128 warps of the cluster instead of landing on one CTA.
constexpr int FPS_PR_NB = 32;                    // blocks per warp: 0..15 keep their min distances in registers, 16..31 in the global
                                                 // scratch (touched only when visited) => 512 blocks = 64 Ki points per CTA, 512 Ki per cloud
// group (J, warp, lane) of CTA `rank` = group ((J * 16 + warp) * 8 + rank) * 32 + lane of the compacted cloud
#define FPS_PR_GG(J) (((((J) * (FPS_CL_THREADS / 32) + warp) * FPS_CL + rank) << 5) + lane)
#define FPS_PR_LOAD(J, A, B, C)                                                                           \
  if ((J) < FPS_CL_S) {                                                                                   \
    A = pts_s[3 * ((J) * FPS_CL_THREADS + tid)]; B = pts_s[3 * ((J) * FPS_CL_THREADS + tid) + 1];         \
    C = pts_s[3 * ((J) * FPS_CL_THREADS + tid) + 2];                                                      \
  } else {                                                                                                \
    const float4* q_ = pc4 + 3 * (int64_t)min(FPS_PR_GG(J), n4 - 1);                                      \
    A = __ldg(q_); B = __ldg(q_ + 1); C = __ldg(q_ + 2);                                                  \
  }

__global__ void __cluster_dims__(FPS_CL, 1, 1) __launch_bounds__(FPS_CL_THREADS, 1)
fps_cluster_pruned_kernel(const float* __restrict__ cp, const int32_t* __restrict__ om, const int32_t* __restrict__ counts, int P, int K,
                          float* __restrict__ mind_g /* (E, P+4) */, float* __restrict__ out, int64_t* __restrict__ out_idx) {
  extern __shared__ __align__(16) float4 pts_s[];                       // [FPS_CL_NS4][3]
  __shared__ unsigned red_v[FPS_CL_THREADS / 32];
  __shared__ int red_i[FPS_CL_THREADS / 32];
  __shared__ __align__(16) float4 red_c[FPS_CL_THREADS / 32];
  __shared__ __align__(16) float4 slot[2][FPS_CL][2];                   // [parity][source rank]{(x,y,z,value bits),(index,-,-,-)}
  __shared__ __align__(8) uint64_t mbar[2];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int e = blockIdx.x / FPS_CL, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* pe = cp + (int64_t)e * (P + 4) * 3;
  const int n4 = (counts[e] + 3) >> 2;                                  // groups of 4 points in the cloud (>= 1)
  const int nB = (n4 + 31) >> 5;                                        // blocks of 32 groups in the cloud
  const int nbl = nB > rank ? (nB - rank + FPS_CL - 1) / FPS_CL : 0;    // blocks of this CTA (local block l = global block l * 8 + rank)
  const float4* pc4 = reinterpret_cast<const float4*>(pe);
  float4* mo = reinterpret_cast<float4*>(mind_g + (int64_t)e * (P + 4));        // min distances of local blocks >= 256, by cloud group
  constexpr float FMAX = 3.402823466e+38f, FINF = __builtin_huge_valf();
  if (tid == 0) {
    pmtc::mbar_init(pmtc::smem_u32(&mbar[0]), 1);
    pmtc::mbar_init(pmtc::smem_u32(&mbar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < min(nbl * 32, FPS_CL_NS4); i += FPS_CL_THREADS) {       // local group i = local block i / 32, lane i % 32
    const int gg = ((((i >> 5) * FPS_CL) + rank) << 5) + (i & 31);
    if (gg < n4) { pts_s[3 * i] = pc4[3 * (int64_t)gg]; pts_s[3 * i + 1] = pc4[3 * (int64_t)gg + 1]; pts_s[3 * i + 2] = pc4[3 * (int64_t)gg + 2]; }
  }
  for (int l = FPS_CL_G * (FPS_CL_THREADS / 32) + warp; l < nbl; l += FPS_CL_THREADS / 32) {
    const int gg = ((l * FPS_CL + rank) << 5) + lane;
    if (gg < n4) mo[gg] = make_float4(FMAX, FMAX, FMAX, FMAX);
  }
  __syncthreads();
  // block summaries: lane j owns block j of this warp
  float4 md[FPS_CL_G];
  float blx = 0.f, bly = 0.f, blz = 0.f, bhx = 0.f, bhy = 0.f, bhz = 0.f, bm = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
  int bp = 0;
#pragma unroll
  for (int j = 0; j < FPS_PR_NB; ++j) {
    const bool valid = FPS_PR_GG(j) < n4;
    const float init = valid ? FMAX : 0.f;                              // lanes beyond the cloud never win and never change
    if (j < FPS_CL_G) md[j < FPS_CL_G ? j : 0] = make_float4(init, init, init, init);
    if (j * (FPS_CL_THREADS / 32) + warp < nbl) {                       // the block holds at least one point (warp-uniform)
      float4 a, b, c;
      FPS_PR_LOAD(j, a, b, c)
      float mnx = valid ? fminf(fminf(a.x, a.w), fminf(b.z, c.y)) : FINF, mxx = valid ? fmaxf(fmaxf(a.x, a.w), fmaxf(b.z, c.y)) : -FINF;
      float mny = valid ? fminf(fminf(a.y, b.x), fminf(b.w, c.z)) : FINF, mxy = valid ? fmaxf(fmaxf(a.y, b.x), fmaxf(b.w, c.z)) : -FINF;
      float mnz = valid ? fminf(fminf(a.z, b.y), fminf(c.x, c.w)) : FINF, mxz = valid ? fmaxf(fmaxf(a.z, b.y), fmaxf(c.x, c.w)) : -FINF;
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, s)); mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, s));
        mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, s)); mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, s));
        mnz = fminf(mnz, __shfl_xor_sync(0xffffffffu, mnz, s)); mxz = fmaxf(mxz, __shfl_xor_sync(0xffffffffu, mxz, s));
      }
      if (lane == j) { blx = mnx; bly = mny; blz = mnz; bhx = mxx; bhy = mxy; bhz = mxz; bm = FMAX; }   // first pick visits it
    }
  }
  float lx = pe[0], ly = pe[1], lz = pe[2];
  if (rank == 0 && tid == 32) {
    out[(int64_t)e * K * 3 + 0] = lx; out[(int64_t)e * K * 3 + 1] = ly; out[(int64_t)e * K * 3 + 2] = lz;
    if (out_idx) out_idx[(int64_t)e * K] = 0;
  }
  cluster.sync();
  if (tid == 0 && K > 1)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(pmtc::smem_u32(&mbar[1])), "r"(FPS_CL_TX) : "memory");
  for (int k = 1; k < K; ++k) {
    const int par = k & 1;
    FPS_STAMP(0)
    // 1. which of this warp's blocks can the new pick reach?
    unsigned mask;
    {
      const float ex = fmaxf(fmaxf(blx - lx, lx - bhx), 0.f), ey = fmaxf(fmaxf(bly - ly, ly - bhy), 0.f), ez = fmaxf(fmaxf(blz - lz, lz - bhz), 0.f);
      const float lower = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
      mask = __ballot_sync(0xffffffffu, !(lower >= bm));
    }
    // 2. update those blocks and their summaries
#define FPS_PR_BLOCK(J)                                                                                   \
    if (mask & (1u << (J))) {                                                                             \
      float4 a_, b_, c_;                                                                                  \
      const int gg_ = FPS_PR_GG(J);                                                                       \
      FPS_PR_LOAD(J, a_, b_, c_)                                                                     \
      float4 m_;                                                                                          \
      if ((J) < FPS_CL_G) m_ = md[(J) < FPS_CL_G ? (J) : 0];                                              \
      else m_ = gg_ < n4 ? mo[gg_] : make_float4(0.f, 0.f, 0.f, 0.f);                                     \
      m_.x = fminf(m_.x, sqdist(a_.x, a_.y, a_.z, lx, ly, lz));                                           \
      m_.y = fminf(m_.y, sqdist(a_.w, b_.x, b_.y, lx, ly, lz));                                           \
      m_.z = fminf(m_.z, sqdist(b_.z, b_.w, c_.x, lx, ly, lz));                                           \
      m_.w = fminf(m_.w, sqdist(c_.y, c_.z, c_.w, lx, ly, lz));                                           \
      if ((J) < FPS_CL_G) md[(J) < FPS_CL_G ? (J) : 0] = m_;                                              \
      else if (gg_ < n4) mo[gg_] = m_;                                                                    \
      float gm_ = m_.x, gx_ = a_.x, gy_ = a_.y, gz_ = a_.z;                                               \
      int gc_ = 0;                                                                                        \
      if (m_.y > gm_) { gm_ = m_.y; gc_ = 1; gx_ = a_.w; gy_ = b_.x; gz_ = b_.y; }                        \
      if (m_.z > gm_) { gm_ = m_.z; gc_ = 2; gx_ = b_.z; gy_ = b_.w; gz_ = c_.x; }                        \
      if (m_.w > gm_) { gm_ = m_.w; gc_ = 3; gx_ = c_.y; gy_ = c_.z; gz_ = c_.w; }                        \
      const unsigned vm_ = __reduce_max_sync(0xffffffffu, __float_as_uint(gm_));                          \
      const int fl_ = __ffs(__ballot_sync(0xffffffffu, __float_as_uint(gm_) == vm_)) - 1;                 \
      const int pc_ = __shfl_sync(0xffffffffu, gc_, fl_);                                                 \
      const float px_ = __shfl_sync(0xffffffffu, gx_, fl_), py_ = __shfl_sync(0xffffffffu, gy_, fl_),     \
                  pz_ = __shfl_sync(0xffffffffu, gz_, fl_);                                               \
      if (lane == (J)) { bm = __uint_as_float(vm_); bp = fl_ * 4 + pc_; bx = px_; by = py_; bz = pz_; }   \
    }
    if (mask & 0x000fu) { FPS_PR_BLOCK(0) FPS_PR_BLOCK(1) FPS_PR_BLOCK(2) FPS_PR_BLOCK(3) }
    if (mask & 0x00f0u) { FPS_PR_BLOCK(4) FPS_PR_BLOCK(5) FPS_PR_BLOCK(6) FPS_PR_BLOCK(7) }
    if (mask & 0x0f00u) { FPS_PR_BLOCK(8) FPS_PR_BLOCK(9) FPS_PR_BLOCK(10) FPS_PR_BLOCK(11) }
    if (mask & 0xf000u) { FPS_PR_BLOCK(12) FPS_PR_BLOCK(13) FPS_PR_BLOCK(14) FPS_PR_BLOCK(15) }
    if (mask & 0xffff0000u) {                                           // blocks whose min distances live in the global scratch
      if (mask & 0x000f0000u) { FPS_PR_BLOCK(16) FPS_PR_BLOCK(17) FPS_PR_BLOCK(18) FPS_PR_BLOCK(19) }
      if (mask & 0x00f00000u) { FPS_PR_BLOCK(20) FPS_PR_BLOCK(21) FPS_PR_BLOCK(22) FPS_PR_BLOCK(23) }
      if (mask & 0x0f000000u) { FPS_PR_BLOCK(24) FPS_PR_BLOCK(25) FPS_PR_BLOCK(26) FPS_PR_BLOCK(27) }
      if (mask & 0xf0000000u) { FPS_PR_BLOCK(28) FPS_PR_BLOCK(29) FPS_PR_BLOCK(30) FPS_PR_BLOCK(31) }
    }
#undef FPS_PR_BLOCK
    FPS_STAMP(1)
    // 3. the warp's candidate from its block summaries (lowest block = lowest index on ties)
    {
      const unsigned bmb = __float_as_uint(bm);
      const unsigned vw = __reduce_max_sync(0xffffffffu, bmb);
      const int jw = __ffs(__ballot_sync(0xffffffffu, bmb == vw)) - 1;
      if (lane == jw) {
        red_v[warp] = vw;
        red_i[warp] = vw ? 4 * ((((jw * (FPS_CL_THREADS / 32) + warp) * FPS_CL) + rank) << 5) + bp : 0x7fffffff;
        red_c[warp] = make_float4(bx, by, bz, 0.f);
      }
      if (nbl > FPS_PR_NB * (FPS_CL_THREADS / 32)) {                    // blocks beyond the 512 summarised ones: streamed, unpruned
        float bvo = 0.f;
        int bio = 0x7fffffff;
        for (int l = FPS_PR_NB * (FPS_CL_THREADS / 32) + warp; l < nbl; l += FPS_CL_THREADS / 32) {     // ascending index per thread
          const int gg = ((l * FPS_CL + rank) << 5) + lane;
          if (gg < n4) {
            const float4 a = __ldg(pc4 + 3 * (int64_t)gg), b = __ldg(pc4 + 3 * (int64_t)gg + 1), c = __ldg(pc4 + 3 * (int64_t)gg + 2);
            float4 m4 = mo[gg];
            m4.x = fminf(m4.x, sqdist(a.x, a.y, a.z, lx, ly, lz)); m4.y = fminf(m4.y, sqdist(a.w, b.x, b.y, lx, ly, lz));
            m4.z = fminf(m4.z, sqdist(b.z, b.w, c.x, lx, ly, lz)); m4.w = fminf(m4.w, sqdist(c.y, c.z, c.w, lx, ly, lz));
            mo[gg] = m4;
            if (m4.x > bvo) { bvo = m4.x; bio = 4 * gg; }
            if (m4.y > bvo) { bvo = m4.y; bio = 4 * gg + 1; }
            if (m4.z > bvo) { bvo = m4.z; bio = 4 * gg + 2; }
            if (m4.w > bvo) { bvo = m4.w; bio = 4 * gg + 3; }
          }
        }
        const unsigned vo = __reduce_max_sync(0xffffffffu, __float_as_uint(bvo));
        const int io = __reduce_min_sync(0xffffffffu, __float_as_uint(bvo) == vo ? bio : 0x7fffffff);
        __syncwarp();
        if (lane == 0 && (vo > red_v[warp] || (vo == red_v[warp] && io < red_i[warp]))) {
          red_v[warp] = vo; red_i[warp] = io;
          red_c[warp] = make_float4(pe[(int64_t)io * 3], pe[(int64_t)io * 3 + 1], pe[(int64_t)io * 3 + 2], 0.f);
        }
      }
    }
    __syncthreads();
    FPS_STAMP(3)
    if (warp == 0) {
      if (lane == 0 && k + 1 < K)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(pmtc::smem_u32(&mbar[par ^ 1])), "r"(FPS_CL_TX) : "memory");
      const unsigned vb = lane < FPS_CL_THREADS / 32 ? red_v[lane] : 0u;
      const int bi = lane < FPS_CL_THREADS / 32 ? red_i[lane] : 0x7fffffff;
      const unsigned vmax = __reduce_max_sync(0xffffffffu, vb);
      const int imin = __reduce_min_sync(0xffffffffu, vb == vmax ? bi : 0x7fffffff);
      const int ww = __ffs(__ballot_sync(0xffffffffu, vb == vmax && bi == imin)) - 1;
      const float4 c = red_c[ww];
      if (lane < FPS_CL) {
        const uint32_t rs = mapa_u32(pmtc::smem_u32(&slot[par][rank][0]), lane), rb = mapa_u32(pmtc::smem_u32(&mbar[par]), lane);
        st_async_v4(rs, c.x, c.y, c.z, __uint_as_float(vmax), rb);
        st_async_v4(rs + 16, __int_as_float(imin), 0.f, 0.f, 0.f, rb);
      }
    }
    FPS_STAMP(4)
    {
      const uint32_t bar = pmtc::smem_u32(&mbar[par]), ph = ((k - 1) >> 1) & 1;
      uint32_t spin = 0;
      while (!pmtc::mbar_try_wait(bar, ph))
        if (++spin > (1u << 22)) __trap();
    }
    FPS_STAMP(5)
    int imin;
    {
      const float4 s0 = slot[par][lane & (FPS_CL - 1)][0];
      const int si = __float_as_int(slot[par][lane & (FPS_CL - 1)][1].x);
      const unsigned sv = __float_as_uint(s0.w);
      const unsigned vmax = __reduce_max_sync(0xffffffffu, sv);
      imin = __reduce_min_sync(0xffffffffu, sv == vmax ? si : 0x7fffffff);
      const int src = __ffs(__ballot_sync(0xffffffffu, sv == vmax && si == imin)) - 1;
      lx = __shfl_sync(0xffffffffu, s0.x, src); ly = __shfl_sync(0xffffffffu, s0.y, src); lz = __shfl_sync(0xffffffffu, s0.z, src);
      if (imin == 0x7fffffff) { imin = 0; lx = pe[0]; ly = pe[1]; lz = pe[2]; }
    }
    FPS_STAMP(6)
    if (rank == 0 && tid == 32) {
      float* o3 = out + ((int64_t)e * K + k) * 3;
      o3[0] = lx; o3[1] = ly; o3[2] = lz;
      if (out_idx) out_idx[(int64_t)e * K + k] = imin;
    }
  }
  if (rank == 0 && out_idx) {
    __syncthreads();
    const int32_t* o = om + (int64_t)e * (P + 4);
    for (int k = tid; k < K; k += FPS_CL_THREADS) out_idx[(int64_t)e * K + k] = o[out_idx[(int64_t)e * K + k]];
  }
  cluster.sync();
}
#undef FPS_PR_LOAD
#undef FPS_PR_GG

// ---- sparse voxels (utils/depth2tsdf.py:88-120, TSDFVolume.sparse_voxel): the voxels whose fused TSDF lies strictly inside
//      (lo, hi), in row-major (x, y, z) order (torch.where), farthest-point-sampled on their integer coordinates, returned as
//      (x, y, z, tsdf).  The band test is an ordered compaction that writes straight into the samplers' compacted-cloud layout
//      (coordinates as exact fp32 integers, voxel number as the original index), so the picks come from fps_compacted_kernel.
__global__ void __launch_bounds__(FPS_THREADS, 1)
tsdf_band_compact_kernel(const float* __restrict__ tsdf /* (E, R3) */, int R, int R3, int Pp /* R3 rounded up to 4 */, float lo, float hi,
                         float* __restrict__ cp /* (E, Pp+4, 3) */, int32_t* __restrict__ om /* (E, Pp+4) */, int32_t* __restrict__ counts) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int e = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* t = tsdf + (int64_t)e * R3;
  float* c = cp + (int64_t)e * (Pp + 4) * 3;
  int32_t* o = om + (int64_t)e * (Pp + 4);
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < R3; base += FPS_THREADS) {
    const int v = base + tid;
    bool keep = false;
    if (v < R3) { const float x = t[v]; keep = x < hi && x > lo; }
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[warp] = __popc(mask);
    __syncthreads();
    int woff = 0, total = 0;
    for (int w = 0; w < 32; ++w) { const int n = s_warp[w]; if (w < warp) woff += n; total += n; }
    const int carry = s_carry;
    if (keep) {
      const int pos = carry + woff + __popc(mask & ((1u << lane) - 1u));
      c[(int64_t)pos * 3] = (float)(v / (R * R)); c[(int64_t)pos * 3 + 1] = (float)((v / R) % R); c[(int64_t)pos * 3 + 2] = (float)(v % R);
      o[pos] = v;
    }
    __syncthreads();
    if (tid == 0) s_carry = carry + total;
    __syncthreads();
  }
  if (tid == 0) {
    int n = s_carry;
    if (n == 0) { c[0] = 0.f; c[1] = 0.f; c[2] = 0.f; o[0] = 0; n = 1; }    // empty band (the reference would fail): voxel 0 stands in
    counts[e] = n;
    while (n & 3) { c[(int64_t)n * 3] = c[0]; c[(int64_t)n * 3 + 1] = c[1]; c[(int64_t)n * 3 + 2] = c[2]; o[n] = o[0]; ++n; }
  }
}

__global__ void __launch_bounds__(256)
tsdf_sparse_gather_kernel(const float* __restrict__ tsdf, int R3, const float* __restrict__ pts /* (E,K,3) */, const int64_t* __restrict__ idx,
                          int E, int K, float* __restrict__ out /* (E,K,4) */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= E * K) return;
  const int e = i / K;
  reinterpret_cast<float4*>(out)[i] = make_float4(pts[(int64_t)i * 3], pts[(int64_t)i * 3 + 1], pts[(int64_t)i * 3 + 2],
                                                  tsdf[(int64_t)e * R3 + idx[i]]);
}

}  // namespace

extern "C" {

int pm_depth2pc_backproject(const float* depth, int E, int M, int H, int W, const float* cam_intr /* host, 3x3 row-major */,
                            const float* cam_pose_dev /* device, (M,4,4) */, const float* vol_origin /* host, 3 */, float size,
                            float* out, pm_stream_t s) {
  PM_REQUIRE(depth && cam_intr && cam_pose_dev && vol_origin && out, PM_ERR_ARG, "pm_depth2pc_backproject: null pointer");
  PM_REQUIRE(E > 0 && M > 0 && H > 0 && W > 0, PM_ERR_SHAPE, "pm_depth2pc_backproject: E=%d M=%d H=%d W=%d", E, M, H, W);
  const int64_t total = (int64_t)E * M * H * W;
  int blocks = (int)((total + 255) / 256 < (int64_t)PM_NUM_SMS * 16 ? (total + 255) / 256 : (int64_t)PM_NUM_SMS * 16);
  if ((H * W) % 4 == 0 && pm_aligned(depth, 16) && pm_aligned(out, 16)) {
    const int64_t t4 = total / 4;
    blocks = (int)((t4 + 255) / 256 < (int64_t)PM_NUM_SMS * 16 ? (t4 + 255) / 256 : (int64_t)PM_NUM_SMS * 16);
    backproject4_kernel<<<blocks, 256, 0, pm_st(s)>>>(depth, E, M, H * W, W, cam_intr[2], cam_intr[5], cam_intr[0], cam_intr[4],
                                                      cam_pose_dev, vol_origin[0], vol_origin[1], vol_origin[2], size, out);
  } else {
    backproject_kernel<<<blocks, 256, 0, pm_st(s)>>>(depth, E, M, H * W, W, cam_intr[2], cam_intr[5], cam_intr[0], cam_intr[4],
                                                     cam_pose_dev, vol_origin[0], vol_origin[1], vol_origin[2], size, out);
  }
  PM_CHECK_LAUNCH("pm_depth2pc_backproject");
  return PM_OK;
}

int pm_depth2pc_backproject_views(const float* const* depth_views_dev, int E, int M, int H, int W, int aligned16, int negate,
                                  float inf_value, const float* cam_intr, const float* cam_pose_dev, const float* vol_origin, float size,
                                  float* out, pm_stream_t s) {
  PM_REQUIRE(depth_views_dev && cam_intr && cam_pose_dev && vol_origin && out, PM_ERR_ARG, "pm_depth2pc_backproject_views: null pointer");
  PM_REQUIRE(E > 0 && M > 0 && H > 0 && W > 0, PM_ERR_SHAPE, "pm_depth2pc_backproject_views: E=%d M=%d H=%d W=%d", E, M, H, W);
  const int64_t total = (int64_t)E * M * H * W;
  const bool vec = aligned16 && (H * W) % 4 == 0 && pm_aligned(out, 16);
  const int64_t work = vec ? total / 4 : total;
  const int blocks = (int)((work + 255) / 256 < (int64_t)PM_NUM_SMS * 16 ? (work + 255) / 256 : (int64_t)PM_NUM_SMS * 16);
  if (vec)
    backproject_views_kernel<true><<<blocks, 256, 0, pm_st(s)>>>(depth_views_dev, E, M, H * W, W, negate, inf_value, cam_intr[2], cam_intr[5],
                                                                 cam_intr[0], cam_intr[4], cam_pose_dev, vol_origin[0], vol_origin[1],
                                                                 vol_origin[2], size, out);
  else
    backproject_views_kernel<false><<<blocks, 256, 0, pm_st(s)>>>(depth_views_dev, E, M, H * W, W, negate, inf_value, cam_intr[2], cam_intr[5],
                                                                  cam_intr[0], cam_intr[4], cam_pose_dev, vol_origin[0], vol_origin[1],
                                                                  vol_origin[2], size, out);
  PM_CHECK_LAUNCH("pm_depth2pc_backproject_views");
  return PM_OK;
}

// workspace: [compacted points (E,P+4,3) | original indices (E,P+4) | counts (E) | min distances (E,P+4)]
size_t pm_fps_ws_bytes(int E, int P) {
  return pm_align_up((size_t)E * (P + 4) * 12, 256) + pm_align_up((size_t)E * (P + 4) * 4, 256) + pm_align_up((size_t)E * 4, 256) +
         pm_align_up((size_t)E * (P + 4) * 4, 256);
}

int pm_fps_cluster_max_active(void) {
  cudaError_t e1 = cudaFuncSetAttribute(fps_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FPS_CL_NS4 * 48);
  if (e1 != cudaSuccess) PM_FAIL(PM_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(FPS_CL * 148, 1, 1);
  cfg.blockDim = dim3(FPS_CL_THREADS, 1, 1);
  cfg.dynamicSmemBytes = FPS_CL_NS4 * 48;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = FPS_CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  e1 = cudaOccupancyMaxActiveClusters(&n, fps_cluster_kernel, &cfg);
  if (e1 != cudaSuccess) PM_FAIL(PM_ERR_CUDA, "cudaOccupancyMaxActiveClusters: %s", cudaGetErrorString(e1));
  return n;
}

int pm_farthest_point_sample(const float* points, int E, int P, int K, int compact, float* out, int64_t* out_idx, void* ws,
                             size_t ws_bytes, pm_stream_t s) {
  PM_REQUIRE(points && out, PM_ERR_ARG, "pm_farthest_point_sample: null pointer");
  PM_REQUIRE(E > 0 && P > 0 && K > 0 && K <= P, PM_ERR_SHAPE, "pm_farthest_point_sample: E=%d P=%d K=%d (need K <= P)", E, P, K);
  if (compact) {
    PM_REQUIRE(ws && ws_bytes >= pm_fps_ws_bytes(E, P) && pm_aligned(ws, 256), PM_ERR_ARG, "pm_farthest_point_sample: workspace too small / unaligned");
    char* base = reinterpret_cast<char*>(ws);
    float* cp = reinterpret_cast<float*>(base);
    int32_t* om = reinterpret_cast<int32_t*>(base + pm_align_up((size_t)E * (P + 4) * 12, 256));
    int32_t* counts = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(om) + pm_align_up((size_t)E * (P + 4) * 4, 256));
    float* mind = reinterpret_cast<float*>(reinterpret_cast<char*>(counts) + pm_align_up((size_t)E * 4, 256));
    PM_REQUIRE(((size_t)(P + 4) * 12) % 16 == 0, PM_ERR_SHAPE, "pm_farthest_point_sample: compact path needs P %% 4 == 0 (P=%d)", P);
    PM_REQUIRE(compact >= 1 && compact <= 4, PM_ERR_ARG, "pm_farthest_point_sample: compact=%d (0 none, 1 auto, 2 cluster, 3 one CTA per cloud, 4 pruned cluster)", compact);
    static bool attr_set_c = false;
    if (!attr_set_c) {
      cudaError_t e1 = cudaFuncSetAttribute(fps_compacted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FPS_SMEM_POINTS * 4);
      cudaError_t e2 = cudaFuncSetAttribute(fps_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FPS_CL_NS4 * 48);
      cudaError_t e3 = cudaFuncSetAttribute(fps_cluster_pruned_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FPS_CL_NS4 * 48);
      if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
        PM_FAIL(PM_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3));
      attr_set_c = true;
    }
    fps_compact_kernel<<<E, FPS_THREADS, 0, pm_st(s)>>>(points, P, cp, om, counts);
    // clouds too large for one CTA's shared memory go to the 8-CTA cluster kernel (registers + DSMEM exchange)
    if (compact == 4 || (compact == 1 && P > FPS_SMEM_POINTS))
      fps_cluster_pruned_kernel<<<E * FPS_CL, FPS_CL_THREADS, FPS_CL_NS4 * 48, pm_st(s)>>>(cp, om, counts, P, K, mind, out, out_idx);
    else if (compact == 2)
      fps_cluster_kernel<<<E * FPS_CL, FPS_CL_THREADS, FPS_CL_NS4 * 48, pm_st(s)>>>(cp, om, counts, P, K, mind, out, out_idx);
    else
      fps_compacted_kernel<<<E, FPS_THREADS, FPS_SMEM_POINTS * 4, pm_st(s)>>>(cp, om, counts, P, K, mind, out, out_idx);
    PM_CHECK_LAUNCH("pm_farthest_point_sample(compact)");
    return PM_OK;
  }
  const bool vec = (P % 4 == 0) && pm_aligned(points, 16);
  if (P <= FPS_SMEM_POINTS) {
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e1 = cudaFuncSetAttribute(fps_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FPS_SMEM_POINTS * 4);
      cudaError_t e2 = cudaFuncSetAttribute(fps_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FPS_SMEM_POINTS * 4);
      if (e1 != cudaSuccess || e2 != cudaSuccess) PM_FAIL(PM_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
      attr_set = true;
    }
    if (vec) fps_kernel<true, true><<<E, FPS_THREADS, (size_t)P * 4, pm_st(s)>>>(points, P, K, nullptr, out, out_idx);
    else fps_kernel<true, false><<<E, FPS_THREADS, (size_t)P * 4, pm_st(s)>>>(points, P, K, nullptr, out, out_idx);
  } else {
    PM_REQUIRE(ws && ws_bytes >= pm_fps_ws_bytes(E, P), PM_ERR_ARG, "pm_farthest_point_sample: workspace too small");
    PM_REQUIRE(pm_aligned(ws, 16), PM_ERR_ALIGN, "pm_farthest_point_sample: workspace must be 16-byte aligned");
    if (vec) fps_kernel<false, true><<<E, FPS_THREADS, 0, pm_st(s)>>>(points, P, K, reinterpret_cast<float*>(ws), out, out_idx);
    else fps_kernel<false, false><<<E, FPS_THREADS, 0, pm_st(s)>>>(points, P, K, reinterpret_cast<float*>(ws), out, out_idx);
  }
  PM_CHECK_LAUNCH("pm_farthest_point_sample");
  return PM_OK;
}

// workspace: [pm_fps_ws_bytes(E, R^3 rounded up to 4) | picked points (E,K,3) | picked voxel numbers (E,K) int64]
size_t pm_tsdf_sparse_voxel_ws_bytes(int E, int resolution, int K) {
  const int Pp = (resolution * resolution * resolution + 3) & ~3;
  return pm_align_up(pm_fps_ws_bytes(E, Pp), 256) + pm_align_up((size_t)E * K * 12, 256) + pm_align_up((size_t)E * K * 8, 256);
}

int pm_tsdf_sparse_voxel(const float* tsdf, int E, int resolution, float lo, float hi, int K, float* out, void* ws, size_t ws_bytes,
                         pm_stream_t s) {
  PM_REQUIRE(tsdf && out && ws, PM_ERR_ARG, "pm_tsdf_sparse_voxel: null pointer");
  PM_REQUIRE(E > 0 && resolution > 0 && resolution <= 512 && K > 0, PM_ERR_SHAPE, "pm_tsdf_sparse_voxel: E=%d resolution=%d K=%d", E, resolution, K);
  PM_REQUIRE(ws_bytes >= pm_tsdf_sparse_voxel_ws_bytes(E, resolution, K) && pm_aligned(ws, 256) && pm_aligned(out, 16), PM_ERR_ARG,
             "pm_tsdf_sparse_voxel: workspace too small / unaligned");
  const int R3 = resolution * resolution * resolution, Pp = (R3 + 3) & ~3;
  char* base = reinterpret_cast<char*>(ws);
  float* cp = reinterpret_cast<float*>(base);
  int32_t* om = reinterpret_cast<int32_t*>(base + pm_align_up((size_t)E * (Pp + 4) * 12, 256));
  int32_t* counts = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(om) + pm_align_up((size_t)E * (Pp + 4) * 4, 256));
  float* mind = reinterpret_cast<float*>(reinterpret_cast<char*>(counts) + pm_align_up((size_t)E * 4, 256));
  float* pts = reinterpret_cast<float*>(base + pm_align_up(pm_fps_ws_bytes(E, Pp), 256));
  int64_t* idx = reinterpret_cast<int64_t*>(reinterpret_cast<char*>(pts) + pm_align_up((size_t)E * K * 12, 256));
  cudaError_t e1 = cudaFuncSetAttribute(fps_compacted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FPS_SMEM_POINTS * 4);
  if (e1 != cudaSuccess) PM_FAIL(PM_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e1));
  tsdf_band_compact_kernel<<<E, FPS_THREADS, 0, pm_st(s)>>>(tsdf, resolution, R3, Pp, lo, hi, cp, om, counts);
  fps_compacted_kernel<<<E, FPS_THREADS, FPS_SMEM_POINTS * 4, pm_st(s)>>>(cp, om, counts, Pp, K, mind, pts, idx);
  tsdf_sparse_gather_kernel<<<(E * K + 255) / 256, 256, 0, pm_st(s)>>>(tsdf, R3, pts, idx, E, K, out);
  PM_CHECK_LAUNCH("pm_tsdf_sparse_voxel");
  return PM_OK;
}

}  // extern "C"
