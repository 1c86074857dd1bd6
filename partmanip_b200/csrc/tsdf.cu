// NEXT ROW (first half) — depth images -> fused TSDF volume, the observation `depth_tsdf` feeds to the Conv3D
// students.  reference: utils/depth2tsdf.py:14-62 (voxel -> pixel tables of __init__ / register_camera) and :68-86 (integrate).
//
//   tsdf_tables_kernel     one thread per (view, voxel): world = origin + voxel_size * (x, y, z), camera = (world - t) . R,
//                          pixel = round-half-even(c * f / z + c0); packs the pixel offset (row * W + col, -1 = outside the
//                          image or behind the camera) and keeps the camera-space depth.  Explicit round-to-nearest mul / add
//                          in the oracle's order, so the tables are the reference's bit for bit.
//   tsdf_integrate_kernel  one thread per (env, voxel): per view gather one depth pixel, d = (depth - z) / trunc clamped at 1,
//                          valid when the pixel exists, depth > 0 and d >= -1; equal-weight mean over the valid views, `default`
//                          where there is none.  Streaming: tables are shared by all envs (L2-resident), 4 B written per voxel.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
tsdf_tables_kernel(const float* __restrict__ cam_pose /* (M,4,4) */, int M, int R, int H, int W, float fx, float fy, float cx, float cy,
                   float ox, float oy, float oz, float voxel_size, int32_t* __restrict__ pix_off, float* __restrict__ pix_z) {
  const int R3 = R * R * R;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * R3) return;
  const int m = i / R3, v = i - m * R3;
  const int x = v / (R * R), y = (v / R) % R, z = v % R;                 // torch.meshgrid 'ij': x slowest (depth2tsdf.py:22-23)
  const float* T = cam_pose + m * 16;
  const float d0 = __fsub_rn(__fadd_rn(ox, __fmul_rn(voxel_size, (float)x)), T[3]);
  const float d1 = __fsub_rn(__fadd_rn(oy, __fmul_rn(voxel_size, (float)y)), T[7]);
  const float d2 = __fsub_rn(__fadd_rn(oz, __fmul_rn(voxel_size, (float)z)), T[11]);
  float c[3];
#pragma unroll
  for (int j = 0; j < 3; ++j)                                            // (world - t) . R[:, j]
    c[j] = __fadd_rn(__fadd_rn(__fmul_rn(d0, T[j]), __fmul_rn(d1, T[4 + j])), __fmul_rn(d2, T[8 + j]));
  const float px = rintf(__fadd_rn(__fdiv_rn(__fmul_rn(c[0], fx), c[2]), cx));
  const float py = rintf(__fadd_rn(__fdiv_rn(__fmul_rn(c[1], fy), c[2]), cy));
  const bool valid = px >= 0.f && px < (float)W && py >= 0.f && py < (float)H && c[2] > 0.f;
  pix_off[i] = valid ? (int)py * W + (int)px : -1;
  pix_z[i] = c[2];
}

__global__ void __launch_bounds__(256)
tsdf_integrate_kernel(const float* __restrict__ depth /* (E,M,HW) */, const int32_t* __restrict__ pix_off, const float* __restrict__ pix_z,
                      int E, int M, int HW, int R3, float trunc, float default_tsdf, float* __restrict__ out /* (E,R3) */) {
  const int64_t total = (int64_t)E * R3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(i / R3), v = (int)(i - (int64_t)e * R3);
    const float* de = depth + (int64_t)e * M * HW;
    int cnt = 0;
    for (int m = 0; m < M; ++m) {
      const int off = __ldg(pix_off + m * R3 + v);
      const float dv = __ldg(de + (int64_t)m * HW + (off < 0 ? 0 : off));           // invalid pixels read pixel (0,0), as the reference
      const float diff = __fsub_rn(dv, __ldg(pix_z + m * R3 + v));
      cnt += (off >= 0 && dv > 0.f && diff >= -trunc) ? 1 : 0;
    }
    const float w = __fdiv_rn(1.f, (float)cnt);                                      // inf when no view is valid: never multiplied in
    float acc = 0.f;
    for (int m = 0; m < M; ++m) {
      const int off = __ldg(pix_off + m * R3 + v);
      const float dv = __ldg(de + (int64_t)m * HW + (off < 0 ? 0 : off));
      const float diff = __fsub_rn(dv, __ldg(pix_z + m * R3 + v));
      const float t = fminf(__fdiv_rn(diff, trunc), 1.f);
      const bool vp = off >= 0 && dv > 0.f && diff >= -trunc;
      const float prod = __fmul_rn(t, vp ? w : 0.f);
      acc = m == 0 ? prod : __fadd_rn(acc, prod);
    }
    out[i] = __fadd_rn(acc, cnt == 0 ? default_tsdf : 0.f);
  }
}

// M <= 4 views (the reference uses 3): one pass, blocked for gather locality.  A CTA owns a 4 x 4 bundle of z-columns (800 voxels
// at R = 50) for a group of 8 envs: the bundle projects onto a narrow band of each view, so the 32-byte sectors its gathers touch
// are shared by many voxels (the z-fastest grid-stride walk of the generic kernel touched about one sector per gather and did it
// twice); the voxel -> pixel table entries are loaded once per voxel and reused for the 8 envs.  Per-view (tsdf, valid) pairs stay
// in registers.  Same arithmetic and order as tsdf_integrate_kernel: bit-identical.  (Measured alternative: env-outer / voxel-inner
// with the table entries of 4 voxels per thread in registers — 1.75 ms instead of 1.14 ms at E = 1024.)
constexpr int TB_COLS = 4, TB_ENVS = 8, TB_THREADS = 256;
__global__ void __launch_bounds__(TB_THREADS)
tsdf_integrate_blocked_kernel(const float* __restrict__ depth, const int32_t* __restrict__ pix_off, const float* __restrict__ pix_z,
                              int E, int M, int HW, int R, float trunc, float default_tsdf, float* __restrict__ out) {
  const int R3 = R * R * R;
  const int nb = (R + TB_COLS - 1) / TB_COLS;                           // column bundles per axis
  const int bx = blockIdx.x / nb, by = blockIdx.x % nb;
  const int e0 = blockIdx.y * TB_ENVS, e1 = min(E, e0 + TB_ENVS);
  const int nx = min(TB_COLS, R - bx * TB_COLS), ny = min(TB_COLS, R - by * TB_COLS);
  const int n_vox = nx * ny * R;
  for (int idx = threadIdx.x; idx < n_vox; idx += TB_THREADS) {
    const int col = idx / R, z = idx - col * R;
    const int x = bx * TB_COLS + col / ny, y = by * TB_COLS + col % ny;
    const int v = (x * R + y) * R + z;
    int off[4];
    float pz[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      off[m] = -1; pz[m] = 0.f;
      if (m < M) { off[m] = __ldg(pix_off + m * R3 + v); pz[m] = __ldg(pix_z + m * R3 + v); }
    }
    for (int e = e0; e < e1; ++e) {
      const float* de = depth + (int64_t)e * M * HW;
      float t[4];
      bool vp[4];
      int cnt = 0;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        t[m] = 0.f;
        vp[m] = false;
        if (m < M) {
          const float dv = __ldg(de + (int64_t)m * HW + (off[m] < 0 ? 0 : off[m]));     // invalid pixels read pixel (0,0), as the reference
          const float diff = __fsub_rn(dv, pz[m]);
          t[m] = fminf(__fdiv_rn(diff, trunc), 1.f);
          vp[m] = off[m] >= 0 && dv > 0.f && diff >= -trunc;
          cnt += vp[m] ? 1 : 0;
        }
      }
      const float w = __fdiv_rn(1.f, (float)cnt);
      float acc = 0.f;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        if (m < M) {
          const float prod = __fmul_rn(t[m], vp[m] ? w : 0.f);
          acc = m == 0 ? prod : __fadd_rn(acc, prod);
        }
      }
      out[(int64_t)e * R3 + v] = __fadd_rn(acc, cnt == 0 ? default_tsdf : 0.f);
    }
  }
}

}  // namespace

extern "C" {

int pm_tsdf_voxel_tables(const float* cam_pose_dev, int M, const float* cam_intr, int H, int W, float size, int resolution,
                         const float* vol_origin, int32_t* pix_off, float* pix_z, pm_stream_t s) {
  PM_REQUIRE(cam_pose_dev && cam_intr && vol_origin && pix_off && pix_z, PM_ERR_ARG, "pm_tsdf_voxel_tables: null pointer");
  PM_REQUIRE(M > 0 && H > 0 && W > 0 && resolution > 0 && resolution <= 512 && size > 0.f, PM_ERR_SHAPE,
             "pm_tsdf_voxel_tables: M=%d H=%d W=%d resolution=%d size=%g", M, H, W, resolution, (double)size);
  const int64_t n = (int64_t)M * resolution * resolution * resolution;
  PM_REQUIRE(n < (1ll << 31), PM_ERR_SHAPE, "pm_tsdf_voxel_tables: M * resolution^3 too large");
  const float voxel_size = (float)((double)size / resolution);
  tsdf_tables_kernel<<<(int)((n + 255) / 256), 256, 0, pm_st(s)>>>(cam_pose_dev, M, resolution, H, W, cam_intr[0], cam_intr[4], cam_intr[2],
                                                                   cam_intr[5], vol_origin[0], vol_origin[1], vol_origin[2], voxel_size,
                                                                   pix_off, pix_z);
  PM_CHECK_LAUNCH("pm_tsdf_voxel_tables");
  return PM_OK;
}

int pm_tsdf_integrate(const float* depth, int E, int M, int H, int W, const int32_t* pix_off, const float* pix_z, float size,
                      int resolution, float default_tsdf, float* out, pm_stream_t s) {
  PM_REQUIRE(depth && pix_off && pix_z && out, PM_ERR_ARG, "pm_tsdf_integrate: null pointer");
  PM_REQUIRE(E > 0 && M > 0 && H > 0 && W > 0 && resolution > 0 && resolution <= 512 && size > 0.f, PM_ERR_SHAPE,
             "pm_tsdf_integrate: E=%d M=%d H=%d W=%d resolution=%d", E, M, H, W, resolution);
  const int R3 = resolution * resolution * resolution;
  const int64_t total = (int64_t)E * R3;
  const float trunc = (float)(4.0 * ((double)size / resolution));                    // depth2tsdf.py:16-17
  if (M <= 4) {
    const int nb = pm_cdiv(resolution, TB_COLS);
    PM_REQUIRE(pm_cdiv(E, TB_ENVS) <= 65535, PM_ERR_SHAPE, "pm_tsdf_integrate: E=%d too large for one launch", E);
    tsdf_integrate_blocked_kernel<<<dim3(nb * nb, pm_cdiv(E, TB_ENVS)), TB_THREADS, 0, pm_st(s)>>>(depth, pix_off, pix_z, E, M, H * W,
                                                                                                  resolution, trunc, default_tsdf, out);
  } else {
    const int blocks = (int)((total + 255) / 256 < (int64_t)PM_NUM_SMS * 16 ? (total + 255) / 256 : (int64_t)PM_NUM_SMS * 16);
    tsdf_integrate_kernel<<<blocks, 256, 0, pm_st(s)>>>(depth, pix_off, pix_z, E, M, H * W, R3, trunc, default_tsdf, out);
  }
  PM_CHECK_LAUNCH("pm_tsdf_integrate");
  return PM_OK;
}

}  // extern "C"
