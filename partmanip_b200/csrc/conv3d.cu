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
// Measured alternatives for the first layer (round 2, not kept; per 2048 volumes, conv1_fwd_kernel = 2.3 ms, conv1_dw_kernel = 2.7 ms):
//  * an implicit-im2col GEMM on the tcgen05 dense kernel (A tiles gathered straight from the volume, three-term split, MMA 16 columns
//    wide): correct, 5.9 ms forward / 4.4 ms weight gradient — with one input channel every gathered element feeds only 16
//    multiply-adds, which does not pay for its index arithmetic, split into bf16 terms and shared-memory image store;
//  * the weights in constant memory (FFMA with uniform-register operands, one shared load per tap): correct, 3.3 ms — the uniform
//    datapath has to deliver a fresh operand for every FFMA and becomes the limiter;
//  * two output positions per thread (each broadcast weight vector feeds 8 FFMAs, 160-thread CTAs): 2.6 ms; eight instead of four
//    position groups in conv1_dw_kernel (256 threads): unchanged.  ncu on the kept kernels: forward LSU data pipe 84 % of peak
//    (shared-memory wavefronts), weight gradient 18 % warp occupancy at 43 % l1tex throughput.
#include "common.cuh"

namespace {

// Patch columns are TAP-major: cols[row, tap * C + c] with tap = (kd * k + kh) * k + kw — a tap's C channels are contiguous on both
// sides (channels-last activations), so for C % 4 == 0 the gather moves 16-byte vectors with one index decomposition per vector
// (the channel-major order of nn.Conv3d's weight.view(Cout, -1) would make every element a 4-byte scatter: 1.9 ms instead of 0.5 ms
// for the second layer at 2048 volumes).  The weights are permuted to the same order by conv3d_weight_permute_kernel (they are tiny).
template <int VEC>
__global__ void __launch_bounds__(256)
im2col3d_kernel(const float* __restrict__ in, int64_t ld_in /* floats between consecutive voxels' channel vectors */, int64_t sample_stride,
                int C, int Din, int k, int s, int pad, int Dout, int Kpad, int64_t n_rows, float* __restrict__ cols) {
  const int CV = C / VEC, K3 = k * k * k, per_row = K3 * CV, P = Dout * Dout * Dout, k2 = k * k, D2 = Dout * Dout;
  const int64_t total = n_rows * per_row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / per_row;
    const int r = (int)(i - row * per_row);
    const int tap = r / CV, q = r - tap * CV;
    const int b = (int)(row / P), pos = (int)(row - (int64_t)b * P);
    const int od = pos / D2, oh = (pos - od * D2) / Dout, ow = pos - od * D2 - oh * Dout;
    const int kd = tap / k2, kh = (tap - kd * k2) / k, kw = tap - kd * k2 - kh * k;
    const int id = od * s - pad + kd, ih = oh * s - pad + kh, iw = ow * s - pad + kw;
    const bool ok = (unsigned)id < (unsigned)Din && (unsigned)ih < (unsigned)Din && (unsigned)iw < (unsigned)Din;
    const float* src = in + (int64_t)b * sample_stride + ((int64_t)(id * Din + ih) * Din + iw) * ld_in + q * VEC;
    float* dst = cols + row * Kpad + tap * C + q * VEC;
    if (VEC == 4) *reinterpret_cast<float4*>(dst) = ok ? __ldg(reinterpret_cast<const float4*>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
    else *dst = ok ? __ldg(src) : 0.f;
  }
  const int padc = Kpad - K3 * C;                                  // padding columns (a multiple-of-4 row stride for the dense layer)
  if (padc > 0)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows * padc; i += (int64_t)gridDim.x * blockDim.x)
      cols[(i / padc) * Kpad + K3 * C + (int)(i % padc)] = 0.f;
}

// din[(b, id, ih, iw), c] = (sum over patches covering the voxel of dcols[(b, od, oh, ow), tap * C + c]) * act'(y[(b, voxel), c])
template <int VEC>
__global__ void __launch_bounds__(256)
col2im3d_kernel(const float* __restrict__ dcols, int Kpad, int C, int Din, int k, int s, int pad, int Dout, int64_t n_in_rows,
                const float* __restrict__ y, int act, float* __restrict__ din) {
  const int CV = C / VEC;
  const int64_t total = n_in_rows * CV;
  const int Pin = Din * Din * Din, P = Dout * Dout * Dout, Din2 = Din * Din;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / CV;
    const int q = (int)(i - row * CV);
    const int b = (int)(row / Pin), pos = (int)(row - (int64_t)b * Pin);
    const int id = pos / Din2, ih = (pos - id * Din2) / Din, iw = pos - id * Din2 - ih * Din;
    float a[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) a[e] = 0.f;
    // od with 0 <= id + pad - od*s < k  <=>  od in [ceil((id + pad - k + 1) / s), floor((id + pad) / s)]
    const int d0 = max(0, (id + pad - k + s) / s), d1 = min(Dout - 1, (id + pad) / s);
    const int h0 = max(0, (ih + pad - k + s) / s), h1 = min(Dout - 1, (ih + pad) / s);
    const int w0 = max(0, (iw + pad - k + s) / s), w1 = min(Dout - 1, (iw + pad) / s);
    for (int od = d0; od <= d1; ++od)
      for (int oh = h0; oh <= h1; ++oh)
        for (int ow = w0; ow <= w1; ++ow) {
          const int kd = id + pad - od * s, kh = ih + pad - oh * s, kw = iw + pad - ow * s;
          const int64_t r = (int64_t)b * P + (od * Dout + oh) * Dout + ow;
          const float* src = dcols + r * Kpad + ((kd * k + kh) * k + kw) * C + q * VEC;
          if (VEC == 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src));
            a[0] += v.x; a[1 % VEC] += v.y; a[2 % VEC] += v.z; a[3 % VEC] += v.w;
          } else {
            a[0] += __ldg(src);
          }
        }
    const int64_t o = row * C + q * VEC;
#pragma unroll
    for (int e = 0; e < VEC; ++e) din[o + e] = a[e] * pm_act_bwd(act, y[o + e]);
  }
}

// to_tap != 0: dst[co][tap][c] = src[co][c][tap] (nn.Conv3d's weight.view(Cout, C, k^3) -> the patch order above); to_tap == 0: the inverse
__global__ void conv3d_weight_permute_kernel(const float* __restrict__ src, int Cout, int C, int K3, int to_tap, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * C * K3) return;
  const int co = i / (C * K3), r = i - co * C * K3;
  if (to_tap) { const int tap = r / C, c = r - tap * C; dst[i] = src[(co * C + c) * K3 + tap]; }
  else        { const int c = r / K3, tap = r - c * K3; dst[i] = src[(co * K3 + tap) * C + c]; }
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

// ---------------------------------------------------------------------------------------------------------------- first layer, direct
// Conv3d(1, 16, k = 5, s = 3, p = 2) on the raw volume (network.py:70): one input channel — as a GEMM it would be a 10^7 x 125 patch
// matrix (5 GB at 2048 samples) against a 16-column weight; done directly instead, in exact fp32 like the reference.
// A CTA owns one output z-plane of one sample: the five input planes it needs sit in shared memory (zero-padded), the 125 x 16
// weights too ([tap][channel], read as broadcast float4s); a thread owns output positions and all 16 channels.
constexpr int C1_OUT = 16, C1_K = 5, C1_TAPS = 125, C1_THREADS = 320;     // stride S: a kernel argument (3 in Conv3DNet, 2 in PoolConv3DNet)

// slab[kd][y][x] = vol[S*od - 2 + kd][y - 2][x - 2] (0 outside), y, x in [0, W): asynchronous 4-byte copies (cp.async with zero fill) —
// every element's load is in flight at once (a plain load loop serialises on the ~700-cycle miss latency: 40 us per slab measured);
// one integer division per row, none per element.  The caller waits with conv1_slab_wait().
__device__ __forceinline__ void conv1_load_slab(float* slab, const float* __restrict__ vol, int Din, int W, int od, int S) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int r = warp; r < C1_K * W; r += nwarps) {
    const int kd = r / W, y = r - kd * W;
    const int id = od * S - 2 + kd, ih = y - 2;
    const bool row_ok = id >= 0 && id < Din && ih >= 0 && ih < Din;
    const float* src = vol + ((int64_t)(row_ok ? id : 0) * Din + (row_ok ? ih : 0)) * Din;
    for (int x = lane; x < W; x += 32) {
      const int iw = x - 2;
      const bool ok = row_ok && iw >= 0 && iw < Din;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(slab + r * W + x);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src + (ok ? iw : 0)), "r"(ok ? 4 : 0) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void conv1_slab_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(C1_THREADS)
conv1_fwd_kernel(const float* __restrict__ x, int64_t ldx, int Din, int Dout, int S, const float* __restrict__ w /* (16,125) */,
                 const float* __restrict__ bias, int act, float* __restrict__ y /* ((b, od, oh, ow), 16) */) {
  extern __shared__ __align__(16) float sm1[];
  const int W = (Dout - 1) * S + C1_K;
  float* slab = sm1;
  float4* wt = reinterpret_cast<float4*>(sm1 + ((C1_K * W * W + 3) / 4 * 4));    // [tap][4 x float4]
  const int od = blockIdx.x, b = blockIdx.y;
  conv1_load_slab(slab, x + (int64_t)b * ldx, Din, W, od, S);
  for (int i = threadIdx.x; i < C1_TAPS * C1_OUT; i += blockDim.x) {
    const int tap = i / C1_OUT, co = i - tap * C1_OUT;
    reinterpret_cast<float*>(wt)[i] = w[co * C1_TAPS + tap];
  }
  conv1_slab_wait();
  __syncthreads();
  for (int pos = threadIdx.x; pos < Dout * Dout; pos += blockDim.x) {
    const int oh = pos / Dout, ow = pos - oh * Dout;
    float acc[C1_OUT];
#pragma unroll
    for (int c = 0; c < C1_OUT; ++c) acc[c] = bias[c];
    const float* base = slab + (oh * S) * W + ow * S;
#pragma unroll 1
    for (int kd = 0; kd < C1_K; ++kd)
#pragma unroll
      for (int kh = 0; kh < C1_K; ++kh)
#pragma unroll
        for (int kw = 0; kw < C1_K; ++kw) {
          const float v = base[(kd * W + kh) * W + kw];
          const float4* wr = wt + ((kd * C1_K + kh) * C1_K + kw) * 4;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 ww = wr[q];
            acc[4 * q] = fmaf(v, ww.x, acc[4 * q]); acc[4 * q + 1] = fmaf(v, ww.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(v, ww.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(v, ww.w, acc[4 * q + 3]);
          }
        }
    float4* dst = reinterpret_cast<float4*>(y + (((int64_t)b * Dout + od) * Dout * Dout + pos) * C1_OUT);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      dst[q] = make_float4(pm_act_fwd(act, acc[4 * q]), pm_act_fwd(act, acc[4 * q + 1]), pm_act_fwd(act, acc[4 * q + 2]),
                           pm_act_fwd(act, acc[4 * q + 3]));
  }
}

// dW1[co][tap] partials: a persistent CTA walks (sample, output plane) pairs; thread = ((kd, kh) tap row, position group g of 4) keeps the
// 5 x 16 accumulators of its five kw taps in registers and visits the positions pos = g (mod 4): per position 5 input loads + 4
// broadcast float4 loads of dpre feed 80 FFMAs.  part[cta][g][tap][16]; the reduce kernel sums ctas and groups in a fixed order.
constexpr int C1_DW_THREADS = 128, C1_DW_GROUPS = 4;
__global__ void __launch_bounds__(C1_DW_THREADS)
conv1_dw_kernel(const float* __restrict__ x, int64_t ldx, int B, int Din, int Dout, int S, const float* __restrict__ dpre /* ((b,od,oh,ow),16) */,
                float* __restrict__ part) {
  extern __shared__ __align__(16) float sm1[];
  const int W = (Dout - 1) * S + C1_K, P2 = Dout * Dout;
  float* slab = sm1;
  float4* dp = reinterpret_cast<float4*>(sm1 + ((C1_K * W * W + 3) / 4 * 4));    // [pos][4 x float4]
  int* off = reinterpret_cast<int*>(dp + P2 * 4);                                // [pos] = (3 oh) W + 3 ow
  const int t = threadIdx.x, kdh = t % 25, g = t / 25;                           // threads >= 100 only help with the loads
  const int row_off = ((kdh / 5) * W + (kdh % 5)) * W;
  float acc[C1_K][C1_OUT];
#pragma unroll
  for (int kw = 0; kw < C1_K; ++kw)
#pragma unroll
    for (int c = 0; c < C1_OUT; ++c) acc[kw][c] = 0.f;
  for (int pos = t; pos < P2; pos += blockDim.x) off[pos] = (pos / Dout) * S * W + (pos % Dout) * S;
  const int64_t pairs = (int64_t)B * Dout;
  for (int64_t pr = blockIdx.x; pr < pairs; pr += gridDim.x) {
    const int b = (int)(pr / Dout), od = (int)(pr - (int64_t)b * Dout);
    __syncthreads();                                                             // previous pair fully consumed
    conv1_load_slab(slab, x + (int64_t)b * ldx, Din, W, od, S);
    const float4* src = reinterpret_cast<const float4*>(dpre + ((int64_t)b * Dout + od) * P2 * C1_OUT);
    for (int i = t; i < P2 * 4; i += blockDim.x) {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(dp + i);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    conv1_slab_wait();
    __syncthreads();
    if (g < C1_DW_GROUPS) {
      for (int pos = g; pos < P2; pos += C1_DW_GROUPS) {
        const float* in = slab + row_off + off[pos];
        float v[C1_K];
#pragma unroll
        for (int kw = 0; kw < C1_K; ++kw) v[kw] = in[kw];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 d = dp[pos * 4 + q];
#pragma unroll
          for (int kw = 0; kw < C1_K; ++kw) {
            acc[kw][4 * q] = fmaf(v[kw], d.x, acc[kw][4 * q]); acc[kw][4 * q + 1] = fmaf(v[kw], d.y, acc[kw][4 * q + 1]);
            acc[kw][4 * q + 2] = fmaf(v[kw], d.z, acc[kw][4 * q + 2]); acc[kw][4 * q + 3] = fmaf(v[kw], d.w, acc[kw][4 * q + 3]);
          }
        }
      }
    }
  }
  if (g < C1_DW_GROUPS) {
#pragma unroll
    for (int kw = 0; kw < C1_K; ++kw) {
      float4* o = reinterpret_cast<float4*>(part + (((int64_t)blockIdx.x * C1_DW_GROUPS + g) * C1_TAPS + kdh * C1_K + kw) * C1_OUT);
#pragma unroll
      for (int q = 0; q < 4; ++q) o[q] = make_float4(acc[kw][4 * q], acc[kw][4 * q + 1], acc[kw][4 * q + 2], acc[kw][4 * q + 3]);
    }
  }
}
// dW1[co][tap] = sum over (cta, group) partials (one warp per output, lanes stride the partials, shuffle tree: fixed order)
__global__ void __launch_bounds__(256)
conv1_dw_reduce_kernel(const float* __restrict__ part, int n_part, float* __restrict__ dW) {
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= C1_TAPS * C1_OUT) return;
  const int tap = i / C1_OUT, co = i - tap * C1_OUT;
  float tsum = 0.f;
  for (int c = lane; c < n_part; c += 32) tsum += part[((int64_t)c * C1_TAPS + tap) * C1_OUT + co];
  tsum = pm_warp_sum(tsum);
  if (lane == 0) dW[co * C1_TAPS + tap] = tsum;
}
// ---------------------------------------------------------------------------------------------------------------- max pooling
// nn.MaxPool3d(kernel_size = k) (stride k, no padding; network.py:103) on channels-last rows: out[(b, cell), c] = max over the cell's
// k^3 voxels, scanned in (d, h, w) order with a strict comparison (the first maximum wins, as in torch); arg = that voxel's index in
// the sample.  Voxels beyond Dp * k (Din not a multiple of k) belong to no cell, exactly as in torch.
__global__ void __launch_bounds__(256)
maxpool3d_fwd_kernel(const float* __restrict__ y, int B, int C, int Din, int k, int Dp, float* __restrict__ out, int32_t* __restrict__ arg) {
  const int64_t total = (int64_t)B * Dp * Dp * Dp * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t cell = i / C;
    const int pw = (int)(cell % Dp); cell /= Dp;
    const int ph = (int)(cell % Dp); cell /= Dp;
    const int pd = (int)(cell % Dp);
    const int b = (int)(cell / Dp);
    const float* src = y + (int64_t)b * Din * Din * Din * C + c;
    float best = -INFINITY;
    int bi = ((pd * k) * Din + ph * k) * Din + pw * k;
    for (int d = 0; d < k; ++d)
      for (int h = 0; h < k; ++h)
        for (int w = 0; w < k; ++w) {
          const int v = ((pd * k + d) * Din + ph * k + h) * Din + pw * k + w;
          const float val = __ldg(src + (int64_t)v * C);
          if (val > best || val != val) { best = val; bi = v; }
        }
    out[i] = best;
    arg[i] = bi;
  }
}

// dpre[(b, voxel), c] = (voxel == arg[(b, cell(voxel)), c] ? dout[(b, cell), c] : 0) * act'(y[(b, voxel), c])
__global__ void __launch_bounds__(256)
maxpool3d_bwd_kernel(const float* __restrict__ dout, const int32_t* __restrict__ arg, const float* __restrict__ y, int act, int B, int C, int Din,
                     int k, int Dp, float* __restrict__ dpre) {
  const int64_t total = (int64_t)B * Din * Din * Din * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t vox = i / C;
    const int w = (int)(vox % Din); vox /= Din;
    const int h = (int)(vox % Din); vox /= Din;
    const int d = (int)(vox % Din);
    const int b = (int)(vox / Din);
    const int pd = d / k, ph = h / k, pw = w / k;
    float g = 0.f;
    if (pd < Dp && ph < Dp && pw < Dp) {
      const int64_t cell = ((((int64_t)b * Dp + pd) * Dp + ph) * Dp + pw) * C + c;
      if (arg[cell] == (d * Din + h) * Din + w) g = dout[cell] * pm_act_bwd(act, y[i]);
    }
    dpre[i] = g;
  }
}

inline size_t conv1_smem_fwd(int Dout, int S) {
  const int W = (Dout - 1) * S + C1_K;
  return ((size_t)(C1_K * W * W + 3) / 4 * 4 + C1_TAPS * C1_OUT) * sizeof(float);
}
inline size_t conv1_smem_dw(int Dout, int S) {
  const int W = (Dout - 1) * S + C1_K;
  return ((size_t)(C1_K * W * W + 3) / 4 * 4 + Dout * Dout * C1_OUT + Dout * Dout) * sizeof(float);
}
constexpr int C1_DW_CTAS = 3 * PM_NUM_SMS;

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
  const bool v4 = (C % 4 == 0) && (ld_in % 4 == 0) && (sample_stride % 4 == 0) && (Kpad % 4 == 0) && pm_aligned(in, 16) && pm_aligned(cols, 16);
  if (v4) im2col3d_kernel<4><<<grid_for(rows * (C / 4) * k * k * k), 256, 0, pm_st(st)>>>(in, ld_in, sample_stride, C, Din, k, s, k / 2, Dout, Kpad, rows, cols);
  else im2col3d_kernel<1><<<grid_for(rows * C * k * k * k), 256, 0, pm_st(st)>>>(in, ld_in, sample_stride, C, Din, k, s, k / 2, Dout, Kpad, rows, cols);
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
  const bool v4 = (C % 4 == 0) && (Kpad % 4 == 0) && pm_aligned(dcols, 16);
  if (v4) col2im3d_kernel<4><<<grid_for(rows * (C / 4)), 256, 0, pm_st(st)>>>(dcols, Kpad, C, Din, k, s, k / 2, Dout, rows, y, act, din);
  else col2im3d_kernel<1><<<grid_for(rows * C), 256, 0, pm_st(st)>>>(dcols, Kpad, C, Din, k, s, k / 2, Dout, rows, y, act, din);
  PM_CHECK_LAUNCH("pm_conv3d_col2im");
  return PM_OK;
}

int pm_conv3d_weight_permute(const float* src, int Cout, int C, int k, int to_tap_major, float* dst, pm_stream_t st) {
  PM_REQUIRE(src && dst && src != dst && Cout > 0 && C > 0 && k > 0, PM_ERR_ARG, "pm_conv3d_weight_permute: bad arguments");
  const int n = Cout * C * k * k * k;
  conv3d_weight_permute_kernel<<<pm_cdiv(n, 256), 256, 0, pm_st(st)>>>(src, Cout, C, k * k * k, to_tap_major, dst);
  PM_CHECK_LAUNCH("pm_conv3d_weight_permute");
  return PM_OK;
}

// first layer of the student, Conv3d(1, 16, 5, stride 3, padding 2) + activation, directly on the volume rows x (B, >= Din^3)
int pm_conv3d_first_forward(const float* x, int64_t ldx, int B, int Din, int stride, const float* w, const float* bias, int act, float* y,
                            pm_stream_t st) {
  PM_REQUIRE(x && w && bias && y && B > 0 && B <= 65535 && Din >= 3 && stride >= 1 && ldx >= (int64_t)Din * Din * Din, PM_ERR_ARG,
             "pm_conv3d_first_forward: bad arguments (B=%d Din=%d stride=%d)", B, Din, stride);
  PM_REQUIRE(act >= PM_ACT_NONE && act <= PM_ACT_SIGMOID, PM_ERR_ARG, "pm_conv3d_first_forward: activation %d", act);
  const int Dout = pm_conv3d_out_dim(Din, C1_K, stride);
  const size_t smem = conv1_smem_fwd(Dout, stride);
  PM_REQUIRE(smem <= 227 * 1024, PM_ERR_UNSUPPORTED, "pm_conv3d_first_forward: volume resolution %d needs %zu B of shared memory", Din, smem);
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(conv1_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) PM_FAIL(PM_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr = smem;
  }
  conv1_fwd_kernel<<<dim3(Dout, B), C1_THREADS, smem, pm_st(st)>>>(x, ldx, Din, Dout, stride, w, bias, act, y);
  PM_CHECK_LAUNCH("pm_conv3d_first_forward");
  return PM_OK;
}

size_t pm_conv3d_first_backward_ws_bytes(void) { return (size_t)C1_DW_CTAS * C1_DW_GROUPS * C1_TAPS * C1_OUT * sizeof(float); }

// dW (16,1,5,5,5) of that layer from dpre ((b, voxel), 16) — no patch matrix; db is a column sum of dpre (pm_rms_colsum et al.)
int pm_conv3d_first_backward(const float* x, int64_t ldx, int B, int Din, int stride, const float* dpre, float* dW, void* ws, pm_stream_t st) {
  PM_REQUIRE(x && dpre && dW && ws && B > 0 && Din >= 3 && stride >= 1, PM_ERR_ARG, "pm_conv3d_first_backward: bad arguments");
  const int Dout = pm_conv3d_out_dim(Din, C1_K, stride);
  const size_t smem = conv1_smem_dw(Dout, stride);
  PM_REQUIRE(smem <= 227 * 1024, PM_ERR_UNSUPPORTED, "pm_conv3d_first_backward: volume resolution %d needs %zu B of shared memory", Din, smem);
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(conv1_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) PM_FAIL(PM_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr = smem;
  }
  float* part = reinterpret_cast<float*>(ws);
  const int64_t pairs = (int64_t)B * Dout;
  const int n_cta = (int)(pairs < C1_DW_CTAS ? pairs : C1_DW_CTAS);
  conv1_dw_kernel<<<n_cta, C1_DW_THREADS, smem, pm_st(st)>>>(x, ldx, B, Din, Dout, stride, dpre, part);
  conv1_dw_reduce_kernel<<<pm_cdiv(C1_TAPS * C1_OUT, 8), 256, 0, pm_st(st)>>>(part, n_cta * C1_DW_GROUPS, dW);
  PM_CHECK_LAUNCH("pm_conv3d_first_backward");
  return PM_OK;
}

int pm_conv3d_flatten(const float* in, float* out, int B, int P, int C, int64_t ld_row, int to_rows, pm_stream_t st) {
  PM_REQUIRE(in && out && B > 0 && P > 0 && C > 0 && ld_row >= (int64_t)P * C, PM_ERR_ARG, "pm_conv3d_flatten: bad arguments");
  flatten3d_kernel<<<grid_for((int64_t)B * P * C), 256, 0, pm_st(st)>>>(in, out, B, P, C, ld_row, to_rows);
  PM_CHECK_LAUNCH("pm_conv3d_flatten");
  return PM_OK;
}

int pm_maxpool3d_forward(const float* y, int B, int C, int Din, int k, float* out, int32_t* argmax, pm_stream_t st) {
  PM_REQUIRE(y && out && argmax && B > 0 && C > 0 && k > 0 && Din >= k, PM_ERR_ARG, "pm_maxpool3d_forward: bad arguments (Din=%d k=%d)", Din, k);
  const int Dp = (Din - k) / k + 1;
  maxpool3d_fwd_kernel<<<grid_for((int64_t)B * Dp * Dp * Dp * C), 256, 0, pm_st(st)>>>(y, B, C, Din, k, Dp, out, argmax);
  PM_CHECK_LAUNCH("pm_maxpool3d_forward");
  return PM_OK;
}

int pm_maxpool3d_backward(const float* dout, const int32_t* argmax, const float* y, int act, int B, int C, int Din, int k, float* dpre,
                          pm_stream_t st) {
  PM_REQUIRE(dout && argmax && y && dpre && B > 0 && C > 0 && k > 0 && Din >= k, PM_ERR_ARG, "pm_maxpool3d_backward: bad arguments");
  PM_REQUIRE(act >= PM_ACT_NONE && act <= PM_ACT_SIGMOID, PM_ERR_ARG, "pm_maxpool3d_backward: activation %d", act);
  const int Dp = (Din - k) / k + 1;
  maxpool3d_bwd_kernel<<<grid_for((int64_t)B * Din * Din * Din * C), 256, 0, pm_st(st)>>>(dout, argmax, y, act, B, C, Din, k, Dp, dpre);
  PM_CHECK_LAUNCH("pm_maxpool3d_backward");
  return PM_OK;
}

}  // extern "C"
