// NEXT ROW (second half) — the Conv3D student that consumes the fused TSDF volume.
// reference: algorithms/algo_utils/network.py:56-63 (conv_stride = nn.Conv3d(stride, padding = k // 2)), :67-97 (Conv3DNet),
// :119-135 (Encoder: Conv3d(1,16,k5,s3) - act - Conv3d(16,32,k3,s3) - act - Conv3d(32,32,k3,s2) - act).
//
// Each convolution runs as  patch gather -> dense layer on the tensor cores (dense_tc.cu: bias + activation fused, three-term bf16
// split for the 1e-4 gate)  with activations kept CHANNELS-LAST ((sample, voxel) rows x channels), so that a layer's output is
// directly the next layer's gather source and the dense kernels see plain row-major matrices:
//   im2col   cols[(b, od, oh, ow), (c, kd, kh, kw)] = in[b, (od*s - p + kd, oh*s - p + kh, ow*s - p + kw), c]   (0 outside)
//            the column order (c, kd, kh, kw) is nn.Conv3d's weight.view(Cout, -1) order: the weight tensor IS the dense layer's W.
//   col2im   the adjoint, in gather form (each input voxel sums the <= ceil(k/s)^3 patch entries that cover it, in a fixed order:
//            deterministic, no atomics), fused with the previous activation's derivative.
//   flatten  (B, 27, 32) channels-last <-> (B, 32*27 [+ proprio]) channel-major rows, the order x.reshape(batch, -1) gives the
//            reference's final_mlp (network.py:92-96).
#include "common.cuh"

namespace {

// one thread per element of cols (row-major, ld = Kpad >= C*k^3; padding columns are written as 0)
__global__ void __launch_bounds__(256)
im2col3d_kernel(const float* __restrict__ in, int64_t ld_in /* floats between consecutive voxels' channel vectors */, int64_t sample_stride,
                int C, int Din, int k, int s, int pad, int Dout, int Kpad, int64_t n_rows, float* __restrict__ cols) {
  const int64_t total = n_rows * Kpad;
  const int K = C * k * k * k, P = Dout * Dout * Dout;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / Kpad;
    const int col = (int)(i - row * Kpad);
    float v = 0.f;
    if (col < K) {
      const int b = (int)(row / P), pos = (int)(row - (int64_t)b * P);
      const int od = pos / (Dout * Dout), oh = (pos / Dout) % Dout, ow = pos % Dout;
      const int c = col / (k * k * k), t = col - c * k * k * k;
      const int kd = t / (k * k), kh = (t / k) % k, kw = t % k;
      const int id = od * s - pad + kd, ih = oh * s - pad + kh, iw = ow * s - pad + kw;
      if (id >= 0 && id < Din && ih >= 0 && ih < Din && iw >= 0 && iw < Din)
        v = __ldg(in + (int64_t)b * sample_stride + ((int64_t)(id * Din + ih) * Din + iw) * ld_in + c);
    }
    cols[i] = v;
  }
}

// din[(b, id, ih, iw), c] = (sum over patches covering the voxel of dcols[(b, od, oh, ow), (c, kd, kh, kw)]) * act'(y[(b, voxel), c])
__global__ void __launch_bounds__(256)
col2im3d_kernel(const float* __restrict__ dcols, int Kpad, int C, int Din, int k, int s, int pad, int Dout, int64_t n_in_rows,
                const float* __restrict__ y, int act, float* __restrict__ din) {
  const int64_t total = n_in_rows * C;
  const int Pin = Din * Din * Din, P = Dout * Dout * Dout;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / C;
    const int c = (int)(i - row * C);
    const int b = (int)(row / Pin), pos = (int)(row - (int64_t)b * Pin);
    const int id = pos / (Din * Din), ih = (pos / Din) % Din, iw = pos % Din;
    float a = 0.f;
    // od with 0 <= id + pad - od*s < k  <=>  od in [ceil((id + pad - k + 1) / s), floor((id + pad) / s)]
    const int d0 = max(0, (id + pad - k + s) / s), d1 = min(Dout - 1, (id + pad) / s);
    const int h0 = max(0, (ih + pad - k + s) / s), h1 = min(Dout - 1, (ih + pad) / s);
    const int w0 = max(0, (iw + pad - k + s) / s), w1 = min(Dout - 1, (iw + pad) / s);
    for (int od = d0; od <= d1; ++od)
      for (int oh = h0; oh <= h1; ++oh)
        for (int ow = w0; ow <= w1; ++ow) {
          const int kd = id + pad - od * s, kh = ih + pad - oh * s, kw = iw + pad - ow * s;
          const int64_t r = (int64_t)b * P + (od * Dout + oh) * Dout + ow;
          a += __ldg(dcols + r * Kpad + ((c * k + kd) * k + kh) * k + kw);
        }
    din[i] = a * pm_act_bwd(act, y[i]);
  }
}

// to_rows != 0: out[b, c*P + pos] = in[(b*P + pos), c]   (channels-last -> the reference's flatten order; out row stride ld_out)
// to_rows == 0: out[(b*P + pos), c] = in[b, c*P + pos]   (gradient of the flattened row back to channels-last; in row stride ld_out)
__global__ void __launch_bounds__(256)
flatten3d_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int P, int C, int64_t ld_row, int to_rows) {
  const int64_t total = (int64_t)B * P * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / (P * C)), rem = (int)(i - (int64_t)b * P * C);
    if (to_rows) {
      const int c = rem / P, pos = rem - c * P;                       // consecutive threads write consecutive row elements
      out[(int64_t)b * ld_row + rem] = in[((int64_t)b * P + pos) * C + c];
    } else {
      const int pos = rem / C, c = rem - pos * C;
      out[i] = in[(int64_t)b * ld_row + c * P + pos];
    }
  }
}

inline int grid_for(int64_t total) {
  int64_t g = (total + 255) / 256;
  const int64_t cap = (int64_t)PM_NUM_SMS * 32;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace

extern "C" {

int pm_conv3d_out_dim(int Din, int k, int s) { return (Din + 2 * (k / 2) - k) / s + 1; }

int pm_conv3d_im2col(const float* in, int64_t ld_in, int64_t sample_stride, int B, int C, int Din, int k, int s, float* cols, int Kpad,
                     pm_stream_t st) {
  PM_REQUIRE(in && cols && B > 0 && C > 0 && Din > 0 && k > 0 && s > 0, PM_ERR_ARG, "pm_conv3d_im2col: bad arguments");
  PM_REQUIRE(Kpad >= C * k * k * k && ld_in >= C, PM_ERR_SHAPE, "pm_conv3d_im2col: Kpad=%d < C*k^3=%d or ld_in < C", Kpad, C * k * k * k);
  const int Dout = pm_conv3d_out_dim(Din, k, s);
  const int64_t rows = (int64_t)B * Dout * Dout * Dout;
  im2col3d_kernel<<<grid_for(rows * Kpad), 256, 0, pm_st(st)>>>(in, ld_in, sample_stride, C, Din, k, s, k / 2, Dout, Kpad, rows, cols);
  PM_CHECK_LAUNCH("pm_conv3d_im2col");
  return PM_OK;
}

int pm_conv3d_col2im(const float* dcols, int Kpad, int B, int C, int Din, int k, int s, const float* y, int act, float* din,
                     pm_stream_t st) {
  PM_REQUIRE(dcols && y && din && B > 0 && C > 0 && Din > 0 && k > 0 && s > 0, PM_ERR_ARG, "pm_conv3d_col2im: bad arguments");
  PM_REQUIRE(Kpad >= C * k * k * k, PM_ERR_SHAPE, "pm_conv3d_col2im: Kpad=%d < C*k^3", Kpad);
  PM_REQUIRE(act >= PM_ACT_NONE && act <= PM_ACT_SIGMOID, PM_ERR_ARG, "pm_conv3d_col2im: activation %d", act);
  const int Dout = pm_conv3d_out_dim(Din, k, s);
  const int64_t rows = (int64_t)B * Din * Din * Din;
  col2im3d_kernel<<<grid_for(rows * C), 256, 0, pm_st(st)>>>(dcols, Kpad, C, Din, k, s, k / 2, Dout, rows, y, act, din);
  PM_CHECK_LAUNCH("pm_conv3d_col2im");
  return PM_OK;
}

int pm_conv3d_flatten(const float* in, float* out, int B, int P, int C, int64_t ld_row, int to_rows, pm_stream_t st) {
  PM_REQUIRE(in && out && B > 0 && P > 0 && C > 0 && ld_row >= (int64_t)P * C, PM_ERR_ARG, "pm_conv3d_flatten: bad arguments");
  flatten3d_kernel<<<grid_for((int64_t)B * P * C), 256, 0, pm_st(st)>>>(in, out, B, P, C, ld_row, to_rows);
  PM_CHECK_LAUNCH("pm_conv3d_flatten");
  return PM_OK;
}

}  // extern "C"
