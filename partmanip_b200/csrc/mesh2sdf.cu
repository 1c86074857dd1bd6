// NEXT ROW (third part) — TSDF of the scene from per-part signed-distance grids (the `mesh_tsdf` observation of
// dagger_tsdf.yaml).  reference: utils/mesh2sdf.py:119-139 (query_tsdf_parallel) + :239-272 (triplet_interpolation_query_parallel).
//
// One thread per (env, voxel): for every part m the voxel centre goes into the part's frame, q = (c - T[e,m]) . R[e,m], the part's
// grid is sampled trilinearly (8 gathers; queries outside [1, res - 2] of the part's OWN resolution give +1), the running minimum
// over the parts and the initial volume is divided by the truncation distance and clamped to [-1, 1].  The reference materialises
// (b, m, n, 3) coordinates, eight (b, m, n) gathers and a dozen temporaries in HBM (~60 B per (env, part, voxel)); here the only
// HBM traffic is the initial volume in and the result out (8 B per voxel) — the part grids (a few MB) and poses are L2 / L1 hits.
#include "common.cuh"

namespace {

struct M2sP {
  const float* field;          // (M, Xm*Ym*Zm) padded grids
  const int32_t* res;          // (M, 3) each part's own resolution
  const float* voxel;          // (M)
  const float* bbox_min;       // (M, 3)
  const float* pose_R;         // (E, M, 3, 3)
  const float* pose_T;         // (E, M, 3)
  const float* init_tsdf;      // (E, R^3)
  float* out;                  // (E, R^3)
  int E, M, R, ry, rz;
  int64_t field_stride;
  float ox, oy, oz, vox, trunc;
};

__global__ void __launch_bounds__(256)
mesh2sdf_query_kernel(const M2sP p) {
  extern __shared__ float sm[];                               // per part: R (9) T (3) bbox_min (3) voxel (1) res (3 as float) = 19 floats
  const int R3 = p.R * p.R * p.R;
  const int e = blockIdx.y;
  for (int i = threadIdx.x; i < p.M * 19; i += blockDim.x) {
    const int m = i / 19, k = i - m * 19;
    float v;
    if (k < 9) v = p.pose_R[((int64_t)e * p.M + m) * 9 + k];
    else if (k < 12) v = p.pose_T[((int64_t)e * p.M + m) * 3 + (k - 9)];
    else if (k < 15) v = p.bbox_min[m * 3 + (k - 12)];
    else if (k == 15) v = p.voxel[m];
    else v = (float)p.res[m * 3 + (k - 16)];
    sm[i] = v;
  }
  __syncthreads();
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= R3) return;
  const int x = v / (p.R * p.R), y = (v / p.R) % p.R, z = v % p.R;
  // centre = index * vox + origin, float32 ops in the reference's order (mesh2sdf.py:32)
  const float cx = __fadd_rn(__fmul_rn((float)x, p.vox), p.ox), cy = __fadd_rn(__fmul_rn((float)y, p.vox), p.oy),
              cz = __fadd_rn(__fmul_rn((float)z, p.vox), p.oz);
  float best = p.init_tsdf[(int64_t)e * R3 + v];
  for (int m = 0; m < p.M; ++m) {
    const float* s = sm + m * 19;
    const float dx = __fsub_rn(cx, s[9]), dy = __fsub_rn(cy, s[10]), dz = __fsub_rn(cz, s[11]);
    float q[3];
#pragma unroll
    for (int j = 0; j < 3; ++j)                               // (c - T) . R[:, j]
      q[j] = __fadd_rn(__fadd_rn(__fmul_rn(dx, s[j]), __fmul_rn(dy, s[3 + j])), __fmul_rn(dz, s[6 + j]));
    float qi[3];
    bool valid = true;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      qi[j] = __fdiv_rn(__fsub_rn(q[j], s[12 + j]), s[15]);
      valid = valid && (qi[j] >= 1.f) && (__fsub_rn(qi[j], s[16 + j]) <= -2.f);
    }
    float val = 1.f;
    if (valid) {
      const int lx = (int)qi[0], ly = (int)qi[1], lz = (int)qi[2];
      const float fx = __fsub_rn(qi[0], (float)lx), fy = __fsub_rn(qi[1], (float)ly), fz = __fsub_rn(qi[2], (float)lz);
      const float* f = p.field + (int64_t)m * p.field_stride + ((int64_t)lx * p.ry + ly) * p.rz + lz;
      const int sy = p.rz, sx = p.rz * p.ry;
      const float gx = __fsub_rn(1.f, fx), gy = __fsub_rn(1.f, fy), gz = __fsub_rn(1.f, fz);
      const float a00 = __fadd_rn(__fmul_rn(__ldg(f), gz), __fmul_rn(__ldg(f + 1), fz));
      const float a01 = __fadd_rn(__fmul_rn(__ldg(f + sy), gz), __fmul_rn(__ldg(f + sy + 1), fz));
      const float a10 = __fadd_rn(__fmul_rn(__ldg(f + sx), gz), __fmul_rn(__ldg(f + sx + 1), fz));
      const float a11 = __fadd_rn(__fmul_rn(__ldg(f + sx + sy), gz), __fmul_rn(__ldg(f + sx + sy + 1), fz));
      const float b0 = __fadd_rn(__fmul_rn(a00, gy), __fmul_rn(a01, fy));
      const float b1 = __fadd_rn(__fmul_rn(a10, gy), __fmul_rn(a11, fy));
      val = __fadd_rn(__fmul_rn(b0, gx), __fmul_rn(b1, fx));
    }
    best = fminf(best, val);
  }
  p.out[(int64_t)e * R3 + v] = fminf(fmaxf(__fdiv_rn(best, p.trunc), -1.f), 1.f);
}

}  // namespace

extern "C" {

int pm_mesh2sdf_query(const float* sdf_field, int64_t field_stride, const int32_t* sdf_res, const float* sdf_voxel, const float* sdf_bbox_min,
                      int M, int bbox_res_y, int bbox_res_z, const float* pose_R, const float* pose_T, const float* init_tsdf, int E,
                      int resolution, const float* vox_origin /* host, 3 */, float size, float* out, pm_stream_t st) {
  PM_REQUIRE(sdf_field && sdf_res && sdf_voxel && sdf_bbox_min && pose_R && pose_T && init_tsdf && vox_origin && out, PM_ERR_ARG,
             "pm_mesh2sdf_query: null pointer");
  PM_REQUIRE(E > 0 && E <= 65535 && M > 0 && M <= 256 && resolution > 0 && resolution <= 512 && size > 0.f, PM_ERR_SHAPE,
             "pm_mesh2sdf_query: E=%d M=%d resolution=%d", E, M, resolution);
  M2sP p{};
  p.field = sdf_field; p.res = sdf_res; p.voxel = sdf_voxel; p.bbox_min = sdf_bbox_min; p.pose_R = pose_R; p.pose_T = pose_T;
  p.init_tsdf = init_tsdf; p.out = out; p.E = E; p.M = M; p.R = resolution; p.ry = bbox_res_y; p.rz = bbox_res_z; p.field_stride = field_stride;
  p.ox = vox_origin[0]; p.oy = vox_origin[1]; p.oz = vox_origin[2];
  p.vox = (float)((double)size / resolution);                          // mesh2sdf.py:24 (python float division, then float32 tensors)
  p.trunc = (float)(4.0 * ((double)size / resolution));
  const int R3 = resolution * resolution * resolution;
  mesh2sdf_query_kernel<<<dim3(pm_cdiv(R3, 256), E), 256, (size_t)M * 19 * sizeof(float), pm_st(st)>>>(p);
  PM_CHECK_LAUNCH("pm_mesh2sdf_query");
  return PM_OK;
}

}  // extern "C"
