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

// workspace: [compacted points (E,P+4,3) | original indices (E,P+4) | counts (E) | min distances (E,P+4)]
size_t pm_fps_ws_bytes(int E, int P) {
  return pm_align_up((size_t)E * (P + 4) * 12, 256) + pm_align_up((size_t)E * (P + 4) * 4, 256) + pm_align_up((size_t)E * 4, 256) +
         pm_align_up((size_t)E * (P + 4) * 4, 256);
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
    static bool attr_set_c = false;
    if (!attr_set_c) {
      cudaError_t e1 = cudaFuncSetAttribute(fps_compacted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FPS_SMEM_POINTS * 4);
      if (e1 != cudaSuccess) PM_FAIL(PM_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e1));
      attr_set_c = true;
    }
    fps_compact_kernel<<<E, FPS_THREADS, 0, pm_st(s)>>>(points, P, cp, om, counts);
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

}  // extern "C"
