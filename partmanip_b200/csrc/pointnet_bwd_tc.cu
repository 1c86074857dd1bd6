// K2 (bf16) — PointNet encoder backward on the 5th-gen tensor cores, one fused persistent kernel.
// reference: autograd of algorithms/algo_utils/network.py:148-150 (per-point Linear-act-Linear-act-Linear) + :182 (max).
//
// The max-pool routes dfeat[b,c] to the single point n = argmax[b,c].  Every (cloud b, channel c) pair is treated as an
// independent ROW: its point's activations are recomputed from the 4C-byte input, and because everything downstream of
// dPre2 is linear, rows that share a point need not be merged (sum of per-row contributions == autograd's result up to
// fp32 summation order).  No compaction, no sort, no gather kernels.
//
// A CTA owns a block of 64 output channels (c0 = 64*(blockIdx & 7)) and walks pairs of clouds; a tile = 2 clouds x 64
// channels = 128 rows (row r: cloud 2p + (r>>6), channel c0 + (r&63)).  Per tile:
//   S1  CUDA cores : gather x[b, argmax], H1 = act(W1 x + b1)                     -> smem bf16 (SW128, K-major)
//   M1  tcgen05    : D2[128 x 256] = H1 . W2^T                     (M=128 N=256 K=128)       TMEM cols [0,256)
//   S2  8 warps    : H2 = act(D2 + b2);  dW3[c,:] += g*H2 (REGISTER accumulators: thread = row = fixed channel);
//                    dPre2 = g * W3[c,:] * act'(H2)                               -> smem bf16 (SW128)
//   M3  tcgen05    : dH1[128 x 128] = dPre2 . W2       (A K-major, B = the same W2 image read MN-major)   cols [0,128)
//   M2  tcgen05    : dW2[256 x 128] += dPre2^T . H1    (both operands MN-major views of the smem tiles above;
//                    accumulated in TMEM cols [256,512) across ALL tiles of the CTA)
//   S3  8 warps    : dPre1 = dH1 * act'(H1) -> smem (in place over H1)
//   M5/M4 tcgen05  : the column sums as N=16 MMAs against Xe = [x | 1 | 0..] (bf16, MN-major, no swizzle):
//                    dPre2^T . Xe -> db2 (the ones column), dPre1^T . Xe -> dW1 | db1; read back per tile (1 tcgen05.ld each)
// Per-CTA partial gradients go to the workspace; a fixed-order reduce kernel produces the six gradient tensors.
// HBM traffic per minibatch: argmax + dfeat (4 KB/cloud) + 12 B per row gathered + ~40 MB of partials.
#include "tc_common.cuh"

namespace {
using namespace pmtc;

constexpr int BT_THREADS = 256;          // 8 warps, thread = (row, column half); thread 0 also issues the MMAs (256 threads => 255 regs).
// Measured alternatives (scripts/bwd_timing.py): 16 compute warps + an MMA warp (96-register cap) spills the dW3
// accumulators to local memory and runs 30 % slower; a dedicated issuer warp does not help either, because the M3/M2 operand
// streams (8 KB of SMEM per 64-cycle MMA) saturate shared memory and the column sums that run beside them slow down equally;
// M2 as 8 N=256 MMAs (dW2 transposed, both operands MN-major) is correct but slower too (0.468 vs 0.427 ms); round 2: 16 warps with
// thread = (row, column QUARTER) at the 128-register cap (88 B of spills) is correct and slower as well (0.463 ms) — the epilogues are
// bound by MUFU.TANH (16/clk) and the shared-memory reads of W3 / H1, not by issue slots.
constexpr int NCB = 8;                   // channel blocks of 64
constexpr uint32_t KB16 = 16384;         // one 64-wide k-block of a 128-row operand
constexpr uint32_t KB32 = 32768;         // one 64-wide k-block of a 256-row operand

// ---- shared-memory map (bytes); operand bases 1024-aligned
constexpr uint32_t SB_W2 = 0;            // bf16 W2 [256 h2 x 128 h1], K-major SW128: 2 k-blocks x 32 KB          = 65536
constexpr uint32_t SB_H1 = 65536;        // bf16 H1 / dPre1 [128 rows x 128]: 2 k-blocks x 16 KB                   = 32768
constexpr uint32_t SB_DP2 = 98304;       // bf16 dPre2 [128 rows x 256]: 4 k-blocks x 16 KB                         = 65536
constexpr uint32_t SB_W3 = 163840;       // bf16x2 W3 block [128 k-pairs][64 ch]                                    = 32768
constexpr uint32_t SB_XE = 196608;       // bf16 Xe [128 rows x 16]: cols 0..C-1 = x, col C = 1, rest 0; MN-major, no swizzle:
                                         // 8-row x 8-col core matrices (128 B), N blocks 128 B apart, K groups 256 B apart =  4096
constexpr uint32_t SB_W1 = 200704;       // float4 W1 rows [128]                                                    =  2048
constexpr uint32_t SB_B1 = 202752;       // fp32 b1 [128]                                                           =   512
constexpr uint32_t SB_B2 = 203264;       // fp32 b2 [256]                                                           =  1024
constexpr uint32_t SB_BAR = 204288;      // 4 mbarriers + tmem slot
constexpr uint32_t SB_TOTAL = 204288 + 64;
// transient N=16 accumulators inside the (free after S2 / S3) upper half of the D2 region
constexpr uint32_t TC_DB2 = 128;         // + hh*16: dPre2^T . Xe for h2-channel half hh
constexpr uint32_t TC_DW1 = 160;         // dPre1^T . Xe

enum { BAR_ACC2_FULL = 0, BAR_ACC1_FULL, BAR_M4_FULL, BAR_W_LOADED, NUM_BARS };

// ---- per-CTA partial-gradient record (floats)
constexpr int PW2 = 0;                   // [256][128]
constexpr int PW3 = 32768;               // [2 cloud parities][64 ch][256]
constexpr int PB3 = 65536;               // [2][64]
constexpr int PB2 = 65664;               // [256]
constexpr int PW1 = 65920;               // [128][4]
constexpr int PB1 = 66432;               // [128]
constexpr int PART_FLOATS = 66560;

constexpr size_t W2IMG_BYTES = 65536;
constexpr size_t W3PACK_BYTES = (size_t)NCB * 32768;

template <int ACT>
__device__ __forceinline__ float act_bwd_c(float y) { return pm_act_bwd(ACT, y); }

__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// ------------------------------------------------------------------------------------------------ weight packing
// W2 fp32 (256,128) -> bf16 K-major SW128 image; W3 fp32 (512,256) -> per channel block [k-pair][64 ch] bf16x2 words
__global__ void pack_bwd_weights_kernel(const float* __restrict__ W2, const float* __restrict__ W3, uint8_t* __restrict__ w2img,
                                        uint32_t* __restrict__ w3pack) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthr = gridDim.x * blockDim.x;
  for (int i = tid; i < 2 * 256 * 8; i += nthr) {          // k-block kb, row n, 16-byte chunk c8
    const int c8 = i & 7, n = (i >> 3) & 255, kb = i >> 11;
    const float* src = W2 + (size_t)n * 128 + kb * 64 + c8 * 8;
    uint4 v;
    v.x = pack_bf16(src[0], src[1]); v.y = pack_bf16(src[2], src[3]);
    v.z = pack_bf16(src[4], src[5]); v.w = pack_bf16(src[6], src[7]);
    *reinterpret_cast<uint4*>(w2img + (size_t)kb * KB32 + n * 128 + ((c8 ^ (n & 7)) << 4)) = v;
  }
  for (int i = tid; i < NCB * 128 * 64; i += nthr) {       // block cb, k-pair k2, channel ch
    const int ch = i & 63, k2 = (i >> 6) & 127, cb = i >> 13;
    const float* src = W3 + (size_t)(cb * 64 + ch) * 256 + 2 * k2;
    w3pack[i] = pack_bf16(src[0], src[1]);
  }
}

// ------------------------------------------------------------------------------------------------ the backward kernel
template <int ACT>
__global__ void __launch_bounds__(BT_THREADS, 1)
encoder_bwd_tc(const float* __restrict__ x, int64_t ldx, int B, int N, int C, const int32_t* __restrict__ argmax,
               const float* __restrict__ dfeat, int64_t lddf, const uint8_t* __restrict__ w2img,
               const uint32_t* __restrict__ w3pack, const float* __restrict__ W1, const float* __restrict__ b1,
               const float* __restrict__ b2, float* __restrict__ part_all, ErrSink err) {
#ifdef PM_TC_TIMING
  long long* dbg = reinterpret_cast<long long*>(err.last + 16);
#define BSTAMP(slot) do { if (blockIdx.x == 0 && tid == 0 && it == 3) dbg[(slot)] = clock64(); } while (0)
#else
#define BSTAMP(slot) do { } while (0)
#endif
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x, j = blockIdx.x;
  const int cb = j & (NCB - 1), slab = j >> 3, n_slabs = (G - cb + NCB - 1) >> 3;
  const int c0 = cb * 64;
  const int P = (B + 1) >> 1;                                     // cloud pairs
  const int n_tiles = slab < P ? (P - slab + n_slabs - 1) / n_slabs : 0;
  float* part = part_all + (size_t)j * PART_FLOATS;
  float4* sW1 = reinterpret_cast<float4*>(smem + SB_W1);
  float* sB1 = reinterpret_cast<float*>(smem + SB_B1);
  float* sB2 = reinterpret_cast<float*>(smem + SB_B2);
  const uint32_t* sW3 = reinterpret_cast<const uint32_t*>(smem + SB_W3);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SB_BAR + 32);
  auto bar = [&](int i) { return sbase + SB_BAR + 8u * i; };

  // ---------------- prologue
  if ((sbase & 1023u) != 0 && tid == 0) err_report(err, 900);
  {
    for (int i = tid; i < 128; i += BT_THREADS) {
      float w[4] = {0.f, 0.f, 0.f, 0.f};
      for (int c = 0; c < C; ++c) w[c] = W1[i * C + c];
      sW1[i] = make_float4(w[0], w[1], w[2], w[3]);
      sB1[i] = b1[i];
    }
    for (int i = tid; i < 256; i += BT_THREADS) sB2[i] = b2[i];
    for (int i = tid; i < 4096 / 16; i += BT_THREADS) reinterpret_cast<uint4*>(smem + SB_XE)[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (tid == 0) {
    for (int i = 0; i < NUM_BARS; ++i) mbar_init(bar(i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // resident weight images through the TMA engine: W2 (64 KB) + this CTA's W3 channel block (32 KB), 16 KB bulk copies
    mbar_expect_tx(bar(BAR_W_LOADED), (uint32_t)(W2IMG_BYTES + 32768));
    for (uint32_t off = 0; off < (uint32_t)W2IMG_BYTES; off += 16384u) bulk_g2s(sbase + SB_W2 + off, w2img + off, 16384u, bar(BAR_W_LOADED));
    const uint8_t* src3 = reinterpret_cast<const uint8_t*>(w3pack + (size_t)cb * 8192);
    for (uint32_t off = 0; off < 32768u; off += 16384u) bulk_g2s(sbase + SB_W3 + off, src3 + off, 16384u, bar(BAR_W_LOADED));
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  bool ok = mbar_wait(bar(BAR_W_LOADED), 0, err, 210);

  const uint32_t idesc_m1 = umma_idesc_ex(128, 256, 0, 0);
  const uint32_t idesc_m3 = umma_idesc_ex(128, 128, 0, 1);
  const uint32_t idesc_m2 = umma_idesc_ex(128, 128, 1, 1);
  const uint32_t idesc_m45 = umma_idesc_ex(128, 16, 1, 1);
  {
    // =========================================================== thread = (row r, column half hsel)
    const int q = warp & 3, hsel = warp >> 2;
    const int r = q * 32 + lane;                                   // tile row == TMEM lane
    const int par = r >> 6, ch = r & 63;
    const int c = c0 + ch;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float acc3[128];                                               // dW3[c, hsel*128 + i] over this thread's clouds
#pragma unroll
    for (int i = 0; i < 128; ++i) acc3[i] = 0.f;
    float db3 = 0.f, db2 = 0.f, dw1[4] = {0.f, 0.f, 0.f, 0.f}, db1 = 0.f;

    // Two-deep prefetch so neither global latency is exposed: (argmax index, dfeat) of tile it+2 and — with the index
    // fetched one tile earlier — the point of tile it+1.  Statically indexed (C <= 4 predicated loads) so the values stay in
    // registers and in flight.
    auto load_idx = [&](int t, int& n, float& g) {
      const int b = 2 * (slab + t * n_slabs) + par;
      n = 0; g = 0.f;
      if (t < n_tiles && b < B) {
        n = __ldg(argmax + (int64_t)b * 512 + c);
        g = __ldg(dfeat + (int64_t)b * lddf + c);
      }
    };
    auto load_x = [&](int t, int n, float& x0, float& x1, float& x2, float& x3) {
      const int b = 2 * (slab + t * n_slabs) + par;
      x0 = x1 = x2 = x3 = 0.f;
      if (t < n_tiles && b < B) {
        n = min(max(n, 0), N - 1);
        const float* xp = x + (int64_t)b * ldx + (int64_t)n * C;
        x0 = __ldg(xp);
        if (C > 1) x1 = __ldg(xp + 1);
        if (C > 2) x2 = __ldg(xp + 2);
        if (C > 3) x3 = __ldg(xp + 3);
      }
    };
    uint32_t h1[32];                          // H1 row of the current (or, once precomputed, the next) tile, packed bf16x2
    auto layer1 = [&](const float (&xv)[4], uint32_t (&h)[32]) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int k = hsel * 64 + 2 * i;
        const float4 w0 = sW1[k], w1 = sW1[k + 1];
        const float a0 = fmaf(xv[3], w0.w, fmaf(xv[2], w0.z, fmaf(xv[1], w0.y, fmaf(xv[0], w0.x, sB1[k]))));
        const float a1 = fmaf(xv[3], w1.w, fmaf(xv[2], w1.z, fmaf(xv[1], w1.y, fmaf(xv[0], w1.x, sB1[k + 1]))));
        h[i] = pack_bf16(act_fast<ACT>(a0), act_fast<ACT>(a1));
      }
    };
    int n1 = 0, n2 = 0;                       // argmax index of tile it+1 / it+2
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;       // dfeat of tile it / it+1 / it+2
    float xn0 = 0.f, xn1 = 0.f, xn2 = 0.f, xn3 = 0.f;
    {
      int n0;
      load_idx(0, n0, g0);
      load_idx(1, n1, g1);
      load_x(0, n0, xn0, xn1, xn2, xn3);
    }

    for (int it = 0; it < n_tiles && ok; ++it) {
      const float g = g0;
      const float xv[4] = {xn0, xn1, xn2, xn3};
      load_idx(it + 2, n2, g2);                                   // prefetch: index two tiles ahead ...
      load_x(it + 1, n1, xn0, xn1, xn2, xn3);                     // ... and the next tile's point through last tile's index
      g0 = g1; g1 = g2; n1 = n2;
      BSTAMP(0);
      // ---- S1: layer 1 for this thread's 64 channels (= one k-block of H1).  Warps 1-7 computed it one tile early, while
      //      the tensor core ran M3/M5/M2 (see below); warp 0 was busy issuing those MMAs and does it here.
      {
        if (it == 0 || warp == 0) layer1(xv, h1);
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8)
          *reinterpret_cast<uint4*>(smem + SB_H1 + hsel * KB16 + sw128(r, c8)) =
              make_uint4(h1[4 * c8], h1[4 * c8 + 1], h1[4 * c8 + 2], h1[4 * c8 + 3]);
        if (hsel == 0) {                                           // Xe row r: [x_0..x_{C-1}, 1, 0, ...] in bf16 (first 8 columns)
          const float e1 = C == 1 ? 1.f : xv[1], e2 = C == 2 ? 1.f : xv[2], e3 = C == 3 ? 1.f : xv[3], e4 = C == 4 ? 1.f : 0.f;
          *reinterpret_cast<uint4*>(smem + SB_XE + (r >> 3) * 256 + (r & 7) * 16) =
              make_uint4(pack_bf16(xv[0], e1), pack_bf16(e2, e3), pack_bf16(e4, 0.f), 0u);
          db3 += g;
        }
      }
      fence_proxy_async();
      BSTAMP(1);
      __syncthreads();
      BSTAMP(2);
      if (tid == 0) {                                              // M1: D2 = H1 . W2^T, K = 128 h1 channels
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t ko = (k & 3) * 32;
          umma_bf16_1cta(tmem_base, umma_desc(sbase + SB_H1 + (k >> 2) * KB16 + ko),
                         umma_desc(sbase + SB_W2 + (k >> 2) * KB32 + ko), idesc_m1, k > 0);
        }
        umma_commit_1cta(bar(BAR_ACC2_FULL));
      }
      __syncwarp();
      // ---- S2: H2, dW3 accumulation, dPre2
      BSTAMP(3);
      ok = mbar_wait(bar(BAR_ACC2_FULL), it & 1, err, 201);
      if (!ok) break;
      tc_fence_after();
      BSTAMP(4);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int col0 = hsel * 128 + cc * 32;
        uint32_t v[32];
        tmem_ld32(lane_taddr + col0, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int k = col0 + 2 * i;
          const float h0 = act_fast<ACT>(__uint_as_float(v[2 * i]) + sB2[k]);
          const float h1v = act_fast<ACT>(__uint_as_float(v[2 * i + 1]) + sB2[k + 1]);
          acc3[cc * 32 + 2 * i] = fmaf(g, h0, acc3[cc * 32 + 2 * i]);
          acc3[cc * 32 + 2 * i + 1] = fmaf(g, h1v, acc3[cc * 32 + 2 * i + 1]);
          const uint32_t w = sW3[(k >> 1) * 64 + ch];
          const float d0 = g * bf16lo(w) * act_bwd_c<ACT>(h0);
          const float d1 = g * bf16hi(w) * act_bwd_c<ACT>(h1v);
          pk[i] = pack_bf16(d0, d1);
        }
#pragma unroll
        for (int qq = 0; qq < 4; ++qq)
          *reinterpret_cast<uint4*>(smem + SB_DP2 + (col0 >> 6) * KB16 + sw128(r, ((col0 & 63) >> 3) + qq)) =
              make_uint4(pk[4 * qq], pk[4 * qq + 1], pk[4 * qq + 2], pk[4 * qq + 3]);
      }
      tc_fence_before();
      fence_proxy_async();
      BSTAMP(5);
      __syncthreads();
      BSTAMP(6);
      if (tid == 0) {
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 16; ++k) {                             // M3: K = 256 h2 channels; B = W2 image, MN-major
          umma_bf16_1cta(tmem_base, umma_desc(sbase + SB_DP2 + (k >> 2) * KB16 + (k & 3) * 32),
                         umma_desc_mn(sbase + SB_W2 + k * 2048, KB32, 1024), idesc_m3, k > 0);
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {                         // M5: dPre2^T . Xe (N = 16): column C is db2 of this tile
            umma_bf16_1cta(tmem_base + TC_DB2 + hh * 16, umma_desc_mn(sbase + SB_DP2 + hh * 2 * KB16 + ks * 2048, KB16, 1024),
                           umma_desc_mn_noswz(sbase + SB_XE + ks * 512, 256, 128), idesc_m45, ks > 0);
          }
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {                         // M2: K = 128 rows; A = dPre2^T, B = H1, both MN-major
            umma_bf16_1cta(tmem_base + 256 + hh * 128, umma_desc_mn(sbase + SB_DP2 + hh * 2 * KB16 + ks * 2048, KB16, 1024),
                           umma_desc_mn(sbase + SB_H1 + ks * 2048, KB16, 1024), idesc_m2, (it > 0 || ks > 0) ? 1u : 0u);
          }
        }
        umma_commit_1cta(bar(BAR_ACC1_FULL));
      }
      __syncwarp();
      BSTAMP(7);
      if (warp != 0 && it + 1 < n_tiles) {                         // next tile's layer 1 under the MMAs (its point is prefetched)
        const float xnext[4] = {xn0, xn1, xn2, xn3};
        layer1(xnext, h1);
      }
      BSTAMP(8);
      // ---- S3: dPre1 = dH1 * act'(H1), in place over H1
      ok = mbar_wait(bar(BAR_ACC1_FULL), it & 1, err, 202);
      if (!ok) break;
      tc_fence_after();
      BSTAMP(9);
      {                                                            // db2 of this tile: lane = h2 channel hsel*128 + r, column C
        uint32_t v[16];
        tmem_ld16(lane_taddr + TC_DB2 + hsel * 16, v);
        tmem_ld_wait();
        db2 += __uint_as_float(C == 1 ? v[1] : (C == 2 ? v[2] : (C == 3 ? v[3] : v[4])));
      }
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t v[32];
        tmem_ld32(lane_taddr + hsel * 64 + cc * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          uint4* p = reinterpret_cast<uint4*>(smem + SB_H1 + hsel * KB16 + sw128(r, cc * 4 + qq));
          const uint4 hq = *p;
          const uint32_t hw[4] = {hq.x, hq.y, hq.z, hq.w};
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float d0 = __uint_as_float(v[qq * 8 + 2 * e]) * act_bwd_c<ACT>(bf16lo(hw[e]));
            const float d1 = __uint_as_float(v[qq * 8 + 2 * e + 1]) * act_bwd_c<ACT>(bf16hi(hw[e]));
            o[e] = pack_bf16(d0, d1);
          }
          *p = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      tc_fence_before();
      fence_proxy_async();
      BSTAMP(10);
      __syncthreads();
      BSTAMP(11);
      // ---- M4: dPre1^T . Xe (M = 128 h1 channels, N = 16, K = 128 rows): columns 0..C-1 = dW1, column C = db1 of this tile
      fence_proxy_async();
      if (tid == 0) {
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_bf16_1cta(tmem_base + TC_DW1, umma_desc_mn(sbase + SB_H1 + ks * 2048, KB16, 1024),
                         umma_desc_mn_noswz(sbase + SB_XE + ks * 512, 256, 128), idesc_m45, ks > 0);
        umma_commit_1cta(bar(BAR_M4_FULL));
      }
      __syncwarp();
      ok = mbar_wait(bar(BAR_M4_FULL), it & 1, err, 206);
      if (!ok) break;
      tc_fence_after();
      if (hsel == 0) {                                             // lane = h1 channel r
        uint32_t v[16];
        tmem_ld16(lane_taddr + TC_DW1, v);
        tmem_ld_wait();
        dw1[0] += __uint_as_float(v[0]); dw1[1] += __uint_as_float(v[1]);
        dw1[2] += __uint_as_float(v[2]); dw1[3] += __uint_as_float(v[3]);
        db1 += __uint_as_float(C == 1 ? v[1] : (C == 2 ? v[2] : (C == 3 ? v[3] : v[4])));
      }
      tc_fence_before();
      BSTAMP(12);
      __syncthreads();                                             // H1 / Xs / dPre2 free for the next tile
      BSTAMP(13);
    }

    // ---------------- epilogue: partial gradients of this CTA
    if (ok) {
      // dW2 from TMEM (lane = h2 channel within the half, column = h1 channel)
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t v[32];
          if (n_tiles > 0) {
            tmem_ld32(lane_taddr + 256 + hh * 128 + hsel * 64 + cc * 32, v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0u;
          }
          float4* dst = reinterpret_cast<float4*>(part + PW2 + (size_t)(hh * 128 + r) * 128 + hsel * 64 + cc * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                 __uint_as_float(v[4 * i + 3]));
        }
      }
      float4* d3 = reinterpret_cast<float4*>(part + PW3 + (size_t)(par * 64 + ch) * 256 + hsel * 128);
#pragma unroll
      for (int i = 0; i < 32; ++i) d3[i] = make_float4(acc3[4 * i], acc3[4 * i + 1], acc3[4 * i + 2], acc3[4 * i + 3]);
      if (hsel == 0) part[PB3 + par * 64 + ch] = db3;
      part[PB2 + hsel * 128 + r] = db2;
      if (hsel == 0) {
        reinterpret_cast<float4*>(part + PW1)[r] = make_float4(dw1[0], dw1[1], dw1[2], dw1[3]);
        part[PB1 + r] = db1;
      }
    }
  }

  // ---------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// fixed-order sum of the per-CTA partials -> the six gradient tensors (overwritten)
__global__ void __launch_bounds__(256)
reduce_bwd_partials_kernel(const float* __restrict__ part, int G, int C, pm_encoder_grads g) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int nW3 = 512 * 256, nW2 = 256 * 128;
  if (idx < nW3) {
    const int c = idx >> 8, k = idx & 255, cb = c >> 6, ch = c & 63;
    float t = 0.f;
    for (int j = cb; j < G; j += NCB) {
      const float* p = part + (size_t)j * PART_FLOATS + PW3 + (size_t)ch * 256 + k;
      t += p[0] + p[64 * 256];
    }
    g.W3[idx] = t;
    return;
  }
  int i = idx - nW3;
  if (i < nW2) {
    float t = 0.f;
    for (int j = 0; j < G; ++j) t += part[(size_t)j * PART_FLOATS + PW2 + i];
    g.W2[i] = t;
    return;
  }
  i -= nW2;
  if (i < 512) {
    const int cb = i >> 6, ch = i & 63;
    float t = 0.f;
    for (int j = cb; j < G; j += NCB) {
      const float* p = part + (size_t)j * PART_FLOATS + PB3 + ch;
      t += p[0] + p[64];
    }
    g.b3[i] = t;
    return;
  }
  i -= 512;
  if (i < 256) {
    float t = 0.f;
    for (int j = 0; j < G; ++j) t += part[(size_t)j * PART_FLOATS + PB2 + i];
    g.b2[i] = t;
    return;
  }
  i -= 256;
  if (i < 128 * 4) {
    const int ch = i >> 2, cc = i & 3;
    if (cc >= C) return;
    float t = 0.f;
    for (int j = 0; j < G; ++j) {
      t += part[(size_t)j * PART_FLOATS + PW1 + ch * 4 + cc];
    }
    g.W1[ch * C + cc] = t;
    return;
  }
  i -= 128 * 4;
  if (i < 128) {
    float t = 0.f;
    for (int j = 0; j < G; ++j) {
      t += part[(size_t)j * PART_FLOATS + PB1 + i];
    }
    g.b1[i] = t;
  }
}

inline int bwd_grid(int B) {
  const int P = (B + 1) / 2;
  if (P >= 19) return PM_NUM_SMS;
  return NCB * P;
}

}  // namespace

extern "C" {

size_t pm_pointnet_encode_backward_tc_ws_bytes(int B, int, int) {
  return W2IMG_BYTES + W3PACK_BYTES + 4096 + (size_t)bwd_grid(B) * PART_FLOATS * sizeof(float);
}

int pm_pointnet_encode_backward_tc(const float* x, int64_t ldx, int B, int N, int C, const pm_encoder_params* p, int act,
                                   const float* dfeat, int64_t lddf, const int32_t* argmax, const pm_encoder_grads* g,
                                   void* ws, size_t ws_bytes, pm_stream_t s) {
  PM_REQUIRE(C >= 1 && C <= 4, PM_ERR_UNSUPPORTED, "bf16 encoder backward: C=%d channels per point (supports 1..4; use PM_PREC_FP32)", C);
  PM_REQUIRE(ws && ws_bytes >= pm_pointnet_encode_backward_tc_ws_bytes(B, N, C), PM_ERR_ARG, "bf16 encoder backward: workspace too small");
  PM_REQUIRE(pm_aligned(ws, 256), PM_ERR_ALIGN, "bf16 encoder backward: workspace must be 256-byte aligned");
  cudaStream_t st = pm_st(s);
  uint8_t* w2img = reinterpret_cast<uint8_t*>(ws);
  uint32_t* w3pack = reinterpret_cast<uint32_t*>(w2img + W2IMG_BYTES);
  int32_t* err = reinterpret_cast<int32_t*>(w2img + W2IMG_BYTES + W3PACK_BYTES);
  float* part = reinterpret_cast<float*>(w2img + W2IMG_BYTES + W3PACK_BYTES + 4096);
  cudaMemsetAsync(err, 0, sizeof(int32_t), st);
  const ErrSink sink{err, pm_tc_sticky_word()};
  pack_bwd_weights_kernel<<<64, 256, 0, st>>>(p->W2, p->W3, w2img, w3pack);
  const int G = bwd_grid(B);
#define PM_BT_LAUNCH(ACTV)                                                                                                  \
  case ACTV: {                                                                                                              \
    static bool attr_set = false;                                                                                           \
    if (!attr_set) {                                                                                                        \
      cudaError_t e1 = cudaFuncSetAttribute(encoder_bwd_tc<ACTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SB_TOTAL); \
      if (e1 != cudaSuccess) PM_FAIL(PM_ERR_CUDA, "cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e1));               \
      attr_set = true;                                                                                                      \
    }                                                                                                                       \
    encoder_bwd_tc<ACTV><<<G, BT_THREADS, SB_TOTAL, st>>>(x, ldx, B, N, C, argmax, dfeat, lddf, w2img, w3pack, p->W1, p->b1, \
                                                          p->b2, part, sink);                                               \
  } break;
  switch (act) {
    PM_BT_LAUNCH(PM_ACT_TANH)
    PM_BT_LAUNCH(PM_ACT_RELU)
    PM_BT_LAUNCH(PM_ACT_ELU)
    PM_BT_LAUNCH(PM_ACT_SELU)
    PM_BT_LAUNCH(PM_ACT_LRELU)
    PM_BT_LAUNCH(PM_ACT_SIGMOID)
    PM_BT_LAUNCH(PM_ACT_NONE)
    default: PM_FAIL(PM_ERR_ARG, "bf16 encoder backward: activation %d", act);
  }
#undef PM_BT_LAUNCH
  const int n_out = 512 * 256 + 256 * 128 + 512 + 256 + 128 * 4 + 128;
  reduce_bwd_partials_kernel<<<pm_cdiv(n_out, 256), 256, 0, st>>>(part, G, C, *g);
  PM_CHECK_LAUNCH("pm_pointnet_encode_backward_tc");
  return PM_OK;
}

// diagnostic: the protocol error word of the last launch (0 = clean); synchronises the stream
int pm_pointnet_bwd_tc_last_error(const void* ws, pm_stream_t s) {
  int32_t h = -1;
  cudaMemcpyAsync(&h, reinterpret_cast<const uint8_t*>(ws) + W2IMG_BYTES + W3PACK_BYTES, sizeof(int32_t), cudaMemcpyDeviceToHost, pm_st(s));
  cudaStreamSynchronize(pm_st(s));
  return h;
}

}  // extern "C"
