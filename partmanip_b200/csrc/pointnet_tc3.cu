// K1 (fp32 parity mode on the tensor cores) — PointNet encoder forward with fp16 hi/lo SPLIT operands: every product
// a.b of layers 2 and 3 is formed as a_hi.b_hi + a_lo.b_hi + a_hi.b_lo (three tcgen05 MMAs, fp32 accumulation in TMEM), which
// carries 22 mantissa bits per operand — the reference's fp32 arithmetic (network.py:148-150, TF32 off) within the 1e-4 gate
// — at a third of the bf16 tensor rate instead of the FFMA rate.
// reference: algorithms/algo_utils/network.py:148-150 (per-point Linear-act-Linear-act-Linear) + :182 (max over points).
//
// Same CTA-pair tile schedule, TMEM plan and epilogue orientation as pointnet_tc.cu (read its header first).  What differs:
//  * operands are fp16 (11-bit mantissa): x = hi + lo with hi = fp16(x), lo = fp16(x - hi).  Activations are bounded
//    (tanh / sigmoid) or clamped to the fp16 range; weights are scaled by 2^6 before the split so that their lo terms stay
//    normal numbers, and the accumulators are scaled back by 2^-6 in the epilogues (exact).
//  * hi AND lo images of H1 / H2 live in shared memory (128 KB per CTA), so the weights no longer fit next to them: the pair
//    STREAMS its weight images from L2 — 20 stages of 16 KB per CTA per 256-point tile, in exactly the order the MMAs consume
//    them — through a 5-deep ring filled by the TMA engine (cp.async.bulk + mbarrier complete_tx; a producer warp per CTA).  The
//    point tile itself (128 points x C floats per CTA) arrives the same way, one tile ahead.
//  * activations are evaluated to fp32 accuracy (tanh.approx's 2^-11 error would eat the whole 1e-4 budget).
// Stage order per tile and CTA rank r (each a [128 rows x 64 k] fp16 SWIZZLE_128B image):
//    0..3   W2 (rows = output channels r*128 + n):  hi kb0, lo kb0, hi kb1, lo kb1
//    4..11  W3 chunk 0 (rows = channels r*256 + m): hi kb0, lo kb0, ... hi kb3, lo kb3
//   12..19  W3 chunk 1 (rows = channels r*256 + 128 + m)
// A "hi" stage feeds 8 MMAs (against the hi and the lo image of the activation k-block), a "lo" stage 4.
#include "tc_common.cuh"
#include <cuda_fp16.h>

namespace {
using namespace pmtc;

constexpr int T3_THREADS = 576;          // warps 0-15: epilogue groups A0 A1 B0 B1 | 16: MMA issue (rank 0) / stage relay (rank 1) | 17: TMA producer
constexpr int PTS_PER_CTA = 128;
constexpr int PTS_PER_TILE = 256;
constexpr uint32_t KB = 16384;           // one 64-wide k-block of a 128-row fp16 operand
constexpr int NSTAGE = 5;
constexpr int STAGES_PER_TILE = 20;
constexpr float WSCALE = 64.f, INV_WSCALE = 1.f / 64.f;

// ---- shared-memory map (bytes); operand bases 1024-aligned
constexpr uint32_t S_HHI = 0;            // H2 hi: 4 k-blocks x 16 KB; H1 hi overlays k-blocks 0-1
constexpr uint32_t S_HLO = 65536;        // H2 lo: 4 k-blocks x 16 KB; H1 lo overlays k-blocks 0-1
constexpr uint32_t S_RING = 131072;      // NSTAGE x 16 KB weight stages
constexpr uint32_t S_X = S_RING + NSTAGE * KB;     // 2 x 2048: the CTA's 128 points x C floats of a tile
constexpr uint32_t S_W1 = S_X + 4096;    // 128 x 4 fp32
constexpr uint32_t S_B1 = S_W1 + 2048;   // 128 fp32
constexpr uint32_t S_BAR = S_B1 + 512;
enum { B_H1_FULL = 0, B_ACC2_FULL, B_H2_KB0, B_H2_KB1, B_H2_KB2, B_H2_KB3, B_G_FULL0, B_G_FULL1, B_G_EMPTY0, B_G_EMPTY1, B_L3_DONE,
       B_W_FULL0, B_W_EMPTY0 = B_W_FULL0 + NSTAGE, B_PEER_FULL0 = B_W_EMPTY0 + NSTAGE, B_X_FULL0 = B_PEER_FULL0 + NSTAGE,
       B_X_EMPTY0 = B_X_FULL0 + 2, NUM_BARS = B_X_EMPTY0 + 2 };
constexpr uint32_t S_TMEM_SLOT = S_BAR + 8 * NUM_BARS;
constexpr uint32_t S_TOTAL = S_BAR + 8 * NUM_BARS + 16;
static_assert(S_TOTAL <= 232448, "shared memory budget");
__host__ __device__ constexpr uint32_t region_col(int p) { return p ? 256u : 0u; }

constexpr size_t WIMG_PER_RANK = (size_t)STAGES_PER_TILE * KB;      // 320 KB

// instruction descriptor: D = f32, A = B = f16, both K-major, dense
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// x = hi + lo in fp16 pairs (low half = first element)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// activations at fp32 accuracy.  tanh(x) = 1 - 2 / (exp(2x) + 1): absolute error <= 4e-7 (ex2.approx / rcp.approx are 1-2 ulp)
template <int ACT>
__device__ __forceinline__ float act_acc(float x) {
  if (ACT == PM_ACT_TANH) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
    return fmaf(-2.f, r, 1.f);
  }
  float y = pm_act_fwd(ACT, x);
  if (ACT != PM_ACT_SIGMOID) y = fminf(fmaxf(y, -60000.f), 60000.f);      // unbounded activations: stay inside fp16's range
  return y;
}

// ------------------------------------------------------------------------------------------------ weight packing
// fp32 W2 (256,128) / W3 (512,256) -> per CTA rank the 20 stage images, scaled by 2^6, hi / lo fp16
__global__ void pack_weights3_kernel(const float* __restrict__ W2, const float* __restrict__ W3, uint8_t* __restrict__ out) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthr = gridDim.x * blockDim.x;
  // one item = one 16-byte chunk (8 k) of one row of one (rank, layer, chunk, k-block): hi and lo images
  for (int i = tid; i < 2 * 10 * 128 * 8; i += nthr) {
    const int c8 = i & 7, row = (i >> 3) & 127;
    const int blk = (i >> 10) % 10, r = (i >> 10) / 10;      // blk 0,1: W2 kb0,kb1 | 2..5: W3 chunk 0 kb0..3 | 6..9: W3 chunk 1
    const float* src;
    int stage_hi;
    if (blk < 2) { src = W2 + (size_t)(r * 128 + row) * 128 + blk * 64 + c8 * 8; stage_hi = blk * 2; }
    else { const int c = (blk - 2) >> 2, kb = (blk - 2) & 3; src = W3 + (size_t)(r * 256 + c * 128 + row) * 256 + kb * 64 + c8 * 8; stage_hi = 4 + c * 8 + kb * 2; }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float a = fminf(fmaxf(src[2 * q] * WSCALE, -60000.f), 60000.f), b = fminf(fmaxf(src[2 * q + 1] * WSCALE, -60000.f), 60000.f);
      split2(a, b, hi[q], lo[q]);
    }
    uint8_t* dst = out + (size_t)r * WIMG_PER_RANK + (size_t)stage_hi * KB + (row * 128 + ((c8 ^ (row & 7)) << 4));
    *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(dst + KB) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float max16(const float (&k)[16]) {
  const float m0 = fmax3(k[0], k[1], k[2]), m1 = fmax3(k[3], k[4], k[5]), m2 = fmax3(k[6], k[7], k[8]);
  const float m3 = fmax3(k[9], k[10], k[11]), m4 = fmax3(k[12], k[13], k[14]);
  return fmaxf(fmax3(m0, m1, m2), fmax3(m3, m4, k[15]));
}

template <int ACT, bool WANT_ARGMAX>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T3_THREADS, 1)
encoder_fwd_tc3(const float* __restrict__ x, int64_t ldx, int B, int N, int C, int x_tma, const uint8_t* __restrict__ wimg,
                const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ b2,
                const float* __restrict__ b3, float* __restrict__ feat, int64_t ldf, int32_t* __restrict__ argmax, ErrSink err,
                uint8_t* __restrict__ scratch) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int tpc = N / PTS_PER_TILE;
  const int n_clouds = cluster_id < B ? (B - cluster_id + n_clusters - 1) / n_clusters : 0;
  const int n_tiles = n_clouds * tpc;
  float* sW1 = reinterpret_cast<float*>(smem + S_W1);
  float* sB1 = reinterpret_cast<float*>(smem + S_B1);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S_TMEM_SLOT);
  auto bar = [&](int i) { return sbase + S_BAR + 8u * i; };
  auto tile_cloud = [&](int it) { return cluster_id + (it / tpc) * n_clusters; };
  // this CTA's 128 points of tile `it`
  auto tile_x = [&](int it) { return x + (int64_t)tile_cloud(it) * ldx + (int64_t)((it % tpc) * PTS_PER_TILE + rank * PTS_PER_CTA) * C; };

  // ---------------- prologue
  if ((sbase & 1023u) != 0 && tid == 0) err_report(err, 920);
  for (int i = tid; i < 128 * 4; i += T3_THREADS) sW1[i] = ((i & 3) < C) ? W1[(i >> 2) * C + (i & 3)] : 0.f;
  if (tid < 128) sB1[tid] = b1[tid];
  if (tid == 0) {
    mbar_init(bar(B_H1_FULL), 4);
    mbar_init(bar(B_ACC2_FULL), 1);
    for (int kb = 0; kb < 4; ++kb) mbar_init(bar(B_H2_KB0 + kb), 8);
    for (int i = 0; i < 2; ++i) { mbar_init(bar(B_G_FULL0 + i), 1); mbar_init(bar(B_G_EMPTY0 + i), 4); }
    mbar_init(bar(B_L3_DONE), 1);
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(bar(B_W_FULL0 + i), 1); mbar_init(bar(B_W_EMPTY0 + i), 1); mbar_init(bar(B_PEER_FULL0 + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(B_X_FULL0 + i), 1); mbar_init(bar(B_X_EMPTY0 + i), 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  bool ok = true;

  auto load_b2 = [&](float4 (&bb)[4], int col) {
#pragma unroll
    for (int q = 0; q < 4; ++q) bb[q] = __ldg(reinterpret_cast<const float4*>(b2 + col) + q);
  };
  // ---- layer-2 epilogue: acc -> * 2^-6 + b2 -> act -> fp16 hi / lo -> smem.  Group g takes the g-th 16-column slice of every
  //      64-channel k-block of H2 (k-blocks complete in order; layer-3 chunk 0 follows k-block by k-block).
  auto e2_slices = [&](uint32_t taddr, int row, int sc, const float4 (&bfirst)[4], int bar_id, bool elect) {
    auto e2_sub = [&](const uint32_t (&v)[16], const float4 (&bb)[4], int kb) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float a0 = fmaf(__uint_as_float(v[4 * q]), INV_WSCALE, bb[q].x), a1 = fmaf(__uint_as_float(v[4 * q + 1]), INV_WSCALE, bb[q].y);
        const float a2 = fmaf(__uint_as_float(v[4 * q + 2]), INV_WSCALE, bb[q].z), a3 = fmaf(__uint_as_float(v[4 * q + 3]), INV_WSCALE, bb[q].w);
        split2(act_acc<ACT>(a0), act_acc<ACT>(a1), hi[2 * q], lo[2 * q]);
        split2(act_acc<ACT>(a2), act_acc<ACT>(a3), hi[2 * q + 1], lo[2 * q + 1]);
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint32_t off = kb * KB + sw128(row, 2 * sc + q);
        *reinterpret_cast<uint4*>(smem + S_HHI + off) = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
        *reinterpret_cast<uint4*>(smem + S_HLO + off) = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
      }
    };
    auto release = [&](int kb, bool last) {
      if (last) tc_fence_before();
      fence_proxy_async();
      named_bar_sync(bar_id, 128);
      if (elect) mbar_arrive_cluster(bar(B_H2_KB0 + kb), 0);
    };
    const uint32_t t0 = taddr + sc * 16;
    uint32_t v0[16], v1[16];
    float4 b0[4], b1v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) b0[q] = bfirst[q];
    tmem_ld16(t0, v0);
    tmem_ld_wait();
    tmem_ld16(t0 + 64, v1);  load_b2(b1v, 64 + sc * 16); e2_sub(v0, b0, 0);  tmem_ld_wait(); release(0, false);
    tmem_ld16(t0 + 128, v0); load_b2(b0, 128 + sc * 16); e2_sub(v1, b1v, 1); tmem_ld_wait(); release(1, false);
    tmem_ld16(t0 + 192, v1); load_b2(b1v, 192 + sc * 16); e2_sub(v0, b0, 2); tmem_ld_wait(); release(2, false);
    e2_sub(v1, b1v, 3);
    release(3, true);
  };

  const int grp = warp >> 2;                                  // 0,1: A0,A1 | 2,3: B0,B1 | 4: warps 16-17
  const int row = (warp & 3) * 32 + lane;
  const bool elect = (tid & 127) == 0;
  const uint32_t lane_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);

  if (grp < 2) {
    // =========================================================== groups A0/A1: layer 1 (k-block grp of H1: 64 channels) + their
    //                                                               slices of the layer-2 epilogue
    float xn[4] = {0.f, 0.f, 0.f, 0.f};
    auto ldg_point = [&](int it, float (&xv)[4]) {
      const float* xp = tile_x(it) + (int64_t)row * C;
      xv[0] = __ldg(xp);
      xv[1] = C > 1 ? __ldg(xp + 1) : 0.f;
      xv[2] = C > 2 ? __ldg(xp + 2) : 0.f;
      xv[3] = C > 3 ? __ldg(xp + 3) : 0.f;
    };
    if (!x_tma && n_tiles > 0) ldg_point(0, xn);
    for (int it = 0; it < n_tiles && ok; ++it) {
      float xv[4];
      if (x_tma) {
        ok = mbar_wait(bar(B_X_FULL0 + (it & 1)), (it >> 1) & 1, err, 921);
        if (!ok) break;
        const float* xs = reinterpret_cast<const float*>(smem + S_X + (it & 1) * 2048) + row * C;
        xv[0] = xs[0];
        xv[1] = C > 1 ? xs[1] : 0.f;
        xv[2] = C > 2 ? xs[2] : 0.f;
        xv[3] = C > 3 ? xs[3] : 0.f;
      } else {
        xv[0] = xn[0]; xv[1] = xn[1]; xv[2] = xn[2]; xv[3] = xn[3];
        if (it + 1 < n_tiles) ldg_point(it + 1, xn);
      }
      // layer 1 into registers (hi / lo) while the previous tile's last layer-3 MMAs run
      uint32_t hhi[32], hlo[32];
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const int ch = grp * 64 + 2 * q;
        const float4 w0 = *reinterpret_cast<const float4*>(sW1 + ch * 4);
        const float4 w1 = *reinterpret_cast<const float4*>(sW1 + (ch + 1) * 4);
        const float a0 = fmaf(xv[3], w0.w, fmaf(xv[2], w0.z, fmaf(xv[1], w0.y, fmaf(xv[0], w0.x, sB1[ch]))));
        const float a1 = fmaf(xv[3], w1.w, fmaf(xv[2], w1.z, fmaf(xv[1], w1.y, fmaf(xv[0], w1.x, sB1[ch + 1]))));
        split2(act_acc<ACT>(a0), act_acc<ACT>(a1), hhi[q], hlo[q]);
      }
      // k-blocks 0-1 of H (hi and lo) are still read by the first half of the previous tile's last layer-3 group
      if (it > 0) ok = mbar_wait(bar(B_L3_DONE), (it - 1) & 1, err, 922);
      if (!ok) break;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        const uint32_t off = grp * KB + sw128(row, c8);
        *reinterpret_cast<uint4*>(smem + S_HHI + off) = make_uint4(hhi[4 * c8], hhi[4 * c8 + 1], hhi[4 * c8 + 2], hhi[4 * c8 + 3]);
        *reinterpret_cast<uint4*>(smem + S_HLO + off) = make_uint4(hlo[4 * c8], hlo[4 * c8 + 1], hlo[4 * c8 + 2], hlo[4 * c8 + 3]);
      }
      fence_proxy_async();
      named_bar_sync(1 + grp, 128);
      if (elect) {
        mbar_arrive_cluster(bar(B_H1_FULL), 0);
        if (x_tma) mbar_arrive(bar(B_X_EMPTY0 + (it & 1)));     // this group has read its x rows of the slot
      }
      float4 bf[4];
      load_b2(bf, grp * 16);
      ok = mbar_wait(bar(B_ACC2_FULL), it & 1, err, 923);
      if (!ok) break;
      tc_fence_after();
      e2_slices(lane_taddr + region_col(it & 1), row, grp, bf, 1 + grp, elect);
    }
  } else if (grp < 4) {
    // =========================================================== groups B0/B1: their slices of the layer-2 epilogue, then the
    //          layer-3 epilogue (thread = channel; B0 columns [0,128), B1 [128,256) of both chunk accumulators)
    const int g = grp - 2;
    float best[2] = {-INFINITY, -INFINITY};
    int bestp[2] = {0, 0};
    uint4* merge = reinterpret_cast<uint4*>(scratch) + (size_t)blockIdx.x * 128;
    for (int it = 0; it < n_tiles && ok; ++it) {
      const int b = tile_cloud(it), j = it % tpc;
      {
        float4 bf[4];
        load_b2(bf, grp * 16);
        ok = mbar_wait(bar(B_ACC2_FULL), it & 1, err, 924);
        if (!ok) break;
        tc_fence_after();
        e2_slices(lane_taddr + region_col(it & 1), row, grp, bf, 1 + grp, elect);
      }
#pragma unroll
      for (int G = 0; G < 2; ++G) {
        ok = mbar_wait(bar(B_G_FULL0 + G), it & 1, err, 925);
        if (!ok) break;
        tc_fence_after();
        float bv = best[G];
        int bp = bestp[G];
        {
          auto e3_sub = [&](uint32_t (&v)[16], int sc) {
            float k[16];
            if (WANT_ARGMAX) {
              // key = value with its low 4 mantissa bits replaced by (15 - column): one FMNMX tree yields max AND position
#pragma unroll
              for (int i = 0; i < 16; ++i) k[i] = __uint_as_float((v[i] & 0xFFFFFFF0u) | (uint32_t)(15 - i));
              const float m = max16(k);
              if (m > bv) { bv = m; bp = j * PTS_PER_TILE + g * 128 + sc * 16; }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) k[i] = __uint_as_float(v[i]);
              bv = fmaxf(bv, max16(k));
            }
          };
          const uint32_t t0 = lane_taddr + region_col((it & 1) ^ (G ^ 1)) + g * 128;
          uint32_t v0[16], v1[16];
          tmem_ld16(t0, v0);
          tmem_ld_wait();
#pragma unroll
          for (int c2 = 0; c2 < 4; ++c2) {
            tmem_ld16(t0 + (2 * c2 + 1) * 16, v1); e3_sub(v0, 2 * c2); tmem_ld_wait();
            if (c2 < 3) tmem_ld16(t0 + (2 * c2 + 2) * 16, v0);
            e3_sub(v1, 2 * c2 + 1);
            if (c2 < 3) tmem_ld_wait();
          }
        }
        best[G] = bv;
        bestp[G] = bp;
        tc_fence_before();
        named_bar_sync(1 + grp, 128);
        if (elect) mbar_arrive_cluster(bar(B_G_EMPTY0 + G), 0);
      }
      if (!ok) break;
      if (j == tpc - 1) {
        if (g == 1) {
          merge[row] = make_uint4(__float_as_uint(best[0]), (uint32_t)bestp[0], __float_as_uint(best[1]), (uint32_t)bestp[1]);
          __threadfence_block();
        }
        named_bar_sync(5, 256);
        if (g == 0) {
          const uint4 o = merge[row];
          const float ob[2] = {__uint_as_float(o.x), __uint_as_float(o.z)};
          const int op[2] = {(int)o.y, (int)o.w};
#pragma unroll
          for (int G = 0; G < 2; ++G) {
            bool take = ob[G] > best[G];
            if (WANT_ARGMAX) {
              // equal VALUES (bit-identical duplicate points, e.g. the env's (0,0,0) padding): the smaller point index wins, like torch.max
              const uint32_t ka = __float_as_uint(ob[G]), kb2 = __float_as_uint(best[G]);
              const float va = __uint_as_float(ka & 0xFFFFFFF0u), vb = __uint_as_float(kb2 & 0xFFFFFFF0u);
              const int ia = op[G] + 15 - (int)(ka & 15u), ib = bestp[G] + 15 - (int)(kb2 & 15u);
              take = va > vb || (va == vb && ia < ib);
            }
            const float bv = take ? ob[G] : best[G];
            const int bp = take ? op[G] : bestp[G];
            const int ch = rank * 256 + G * 128 + row;
            const uint32_t kbits = __float_as_uint(bv);
            if (WANT_ARGMAX) {
              feat[(int64_t)b * ldf + ch] = fmaf(__uint_as_float(kbits & 0xFFFFFFF0u), INV_WSCALE, __ldg(b3 + ch));
              argmax[(int64_t)b * 512 + ch] = bp + 15 - (int)(kbits & 15u);
            } else {
              feat[(int64_t)b * ldf + ch] = fmaf(bv, INV_WSCALE, __ldg(b3 + ch));
            }
          }
        }
        named_bar_sync(5, 256);          // merge[] is reused by the next cloud
#pragma unroll
        for (int G = 0; G < 2; ++G) { best[G] = -INFINITY; bestp[G] = 0; }
      }
    }
  } else if (warp == 17) {
    // =========================================================== TMA producer (one lane per CTA): x tiles one tile ahead, then the
    //          20 weight stages of every tile in consumption order through the NSTAGE-deep ring
    if (lane == 0) {
      const uint8_t* wsrc = wimg + (size_t)rank * WIMG_PER_RANK;
      const uint32_t xbytes = (uint32_t)(PTS_PER_CTA * C * sizeof(float));
      auto load_x = [&](int t) -> bool {
        const int xs = t & 1;
        if (t >= 2 && !mbar_wait(bar(B_X_EMPTY0 + xs), ((t >> 1) - 1) & 1, err, 926)) return false;
        mbar_expect_tx(bar(B_X_FULL0 + xs), xbytes);
        bulk_g2s(sbase + S_X + xs * 2048, tile_x(t), xbytes, bar(B_X_FULL0 + xs));
        return true;
      };
      if (x_tma && n_tiles > 0) ok = load_x(0);
      for (int it = 0; it < n_tiles && ok; ++it) {
        if (x_tma && it + 1 < n_tiles) ok = load_x(it + 1);
#pragma unroll 1
        for (int j = 0; j < STAGES_PER_TILE && ok; ++j) {
          const int slot = j % NSTAGE;
          // n-th fill of the slot overall = it*4 + j/5; it waits for the consumption of the previous fill
          if (it > 0 || j >= NSTAGE) ok = mbar_wait(bar(B_W_EMPTY0 + slot), ((j / NSTAGE) + 1) & 1, err, 927);
          if (!ok) break;
          mbar_expect_tx(bar(B_W_FULL0 + slot), KB);
          bulk_g2s(sbase + S_RING + slot * KB, wsrc + (size_t)j * KB, KB, bar(B_W_FULL0 + slot));
        }
      }
    }
  } else if (rank == 1) {
    // =========================================================== warp 16 of CTA 1: tells the leader when CTA 1's copy of a stage has
    //          landed (a bulk copy can only signal a barrier of its own CTA)
    if (lane == 0) {
      for (int it = 0; it < n_tiles && ok; ++it) {
#pragma unroll 1
        for (int j = 0; j < STAGES_PER_TILE && ok; ++j) {
          const int slot = j % NSTAGE;
          ok = mbar_wait(bar(B_W_FULL0 + slot), (j / NSTAGE) & 1, err, 928);
          if (ok) mbar_arrive_cluster(bar(B_PEER_FULL0 + slot), 0);
        }
      }
    }
  } else {
    // =========================================================== warp 16 of the leader CTA: MMA issue
    const uint32_t idesc = umma_idesc_f16(256, 256);
    // four K=16 steps of one k-block: D[dcol] (+)= A[a_addr] . B[b_addr]^T
    auto mma4 = [&](uint32_t dcol, uint32_t a_addr, uint32_t b_addr, bool fresh) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16_2cta(tmem_base + dcol, umma_desc(a_addr + k * 32), umma_desc(b_addr + k * 32), idesc, (fresh && k == 0) ? 0u : 1u);
    };
    auto wait_stage = [&](int j) -> bool {
      const int slot = j % NSTAGE;
      const uint32_t par = (j / NSTAGE) & 1;
      const bool r = mbar_wait(bar(B_W_FULL0 + slot), par, err, 929) && mbar_wait(bar(B_PEER_FULL0 + slot), par, err, 930);
      tc_fence_after();
      return r;
    };
    auto ring = [&](int j) { return sbase + S_RING + (uint32_t)(j % NSTAGE) * KB; };
    for (int it = 0; it < n_tiles && ok; ++it) {
      const uint32_t colp = region_col(it & 1), colq = region_col((it & 1) ^ 1);
      // region p held layer-3 group 0 of the previous tile: E3 must have drained it before layer 2 overwrites it
      ok = mbar_wait(bar(B_H1_FULL), it & 1, err, 931) && mbar_wait(bar(B_G_EMPTY0), (it & 1) ^ 1, err, 932);
      if (!ok) break;
      tc_fence_after();
      // ---- layer 2: D2 = (H1hi + H1lo) . (W2hi + W2lo)^T without the lo.lo term
#pragma unroll 1
      for (int kb = 0; kb < 2 && ok; ++kb) {
        int j = 2 * kb;
        ok = wait_stage(j);
        if (!ok) break;
        if (lane == 0) {
          mma4(colp, sbase + S_HHI + kb * KB, ring(j), kb == 0);
          mma4(colp, sbase + S_HLO + kb * KB, ring(j), false);
          umma_commit_mc(bar(B_W_EMPTY0 + j % NSTAGE));
        }
        __syncwarp();
        ++j;
        ok = wait_stage(j);
        if (!ok) break;
        if (lane == 0) {
          mma4(colp, sbase + S_HHI + kb * KB, ring(j), false);
          umma_commit_mc(bar(B_W_EMPTY0 + j % NSTAGE));
        }
        __syncwarp();
      }
      if (!ok) break;
      if (lane == 0) umma_commit_mc(bar(B_ACC2_FULL));
      __syncwarp();
      // ---- layer 3, channel chunk 0 into region p^1 (held chunk 1 of the previous tile), k-block by k-block behind E2;
      //      then chunk 1 into region p, which E2 has drained by then (all four H2 k-blocks arrived)
      ok = mbar_wait(bar(B_G_EMPTY1), (it & 1) ^ 1, err, 933);
      if (!ok) break;
#pragma unroll 1
      for (int c = 0; c < 2 && ok; ++c) {
        const uint32_t dcol = c == 0 ? colq : colp;
#pragma unroll 1
        for (int kb = 0; kb < 4 && ok; ++kb) {
          if (c == 0) {
            ok = mbar_wait(bar(B_H2_KB0 + kb), it & 1, err, 934);
            if (!ok) break;
            tc_fence_after();
          }
          int j = 4 + c * 8 + kb * 2;
          ok = wait_stage(j);
          if (!ok) break;
          if (lane == 0) {
            mma4(dcol, ring(j), sbase + S_HHI + kb * KB, kb == 0);
            mma4(dcol, ring(j), sbase + S_HLO + kb * KB, false);
            umma_commit_mc(bar(B_W_EMPTY0 + j % NSTAGE));
          }
          __syncwarp();
          ++j;
          ok = wait_stage(j);
          if (!ok) break;
          if (lane == 0) {
            mma4(dcol, ring(j), sbase + S_HHI + kb * KB, false);
            umma_commit_mc(bar(B_W_EMPTY0 + j % NSTAGE));
            // after k-block 1 of the LAST chunk nothing reads k-blocks 0-1 of H any more — where the next tile's H1 goes
            if (c == 1 && kb == 1) umma_commit_mc(bar(B_L3_DONE));
          }
          __syncwarp();
        }
        if (!ok) break;
        if (lane == 0) umma_commit_mc(bar(c == 0 ? B_G_FULL0 : B_G_FULL1));
        __syncwarp();
      }
    }
  }

  // ---------------- teardown
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 16) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

extern "C" {

// workspace: the two stage-ordered weight images (rank 0 / rank 1) + an error word + the per-CTA merge scratch
size_t pm_pointnet_encode_forward_tc3_ws_bytes(int, int, int) { return 2 * WIMG_PER_RANK + 4096 + (size_t)PM_NUM_SMS * 2048; }

int pm_pointnet_encode_forward_tc3_supported(int N, int C) { return C >= 1 && C <= 4 && N % PTS_PER_TILE == 0; }

int pm_pointnet_encode_forward_tc3(const float* x, int64_t ldx, int B, int N, int C, const pm_encoder_params* p, int act,
                                   float* feat, int64_t ldf, int32_t* argmax, void* ws, size_t ws_bytes, pm_stream_t s) {
  PM_REQUIRE(pm_pointnet_encode_forward_tc3_supported(N, C), PM_ERR_UNSUPPORTED, "split-fp16 encoder: N=%d C=%d (N %% 256 == 0, C <= 4)", N, C);
  PM_REQUIRE(ws && ws_bytes >= pm_pointnet_encode_forward_tc3_ws_bytes(B, N, C), PM_ERR_ARG, "split-fp16 encoder: workspace too small");
  PM_REQUIRE(pm_aligned(ws, 256), PM_ERR_ALIGN, "split-fp16 encoder: workspace must be 256-byte aligned");
  PM_REQUIRE(pm_aligned(p->b2, 16), PM_ERR_ALIGN, "split-fp16 encoder: b2 must be 16-byte aligned");
  cudaStream_t st = pm_st(s);
  uint8_t* wimg = reinterpret_cast<uint8_t*>(ws);
  int32_t* errw = reinterpret_cast<int32_t*>(wimg + 2 * WIMG_PER_RANK);
  uint8_t* scratch = wimg + 2 * WIMG_PER_RANK + 4096;
  cudaMemsetAsync(errw, 0, sizeof(int32_t), st);
  const ErrSink sink{errw, pm_tc_sticky_word()};
  pack_weights3_kernel<<<80, 256, 0, st>>>(p->W2, p->W3, wimg);
  // the point tiles travel by TMA when every tile start is 16-byte aligned (a cloud row may carry a proprio tail: ldx = N*C + p)
  const int x_tma = pm_aligned(x, 16) && ((ldx * (int64_t)sizeof(float)) % 16 == 0) ? 1 : 0;
  int n_clusters = B < PM_NUM_SMS / 2 ? B : PM_NUM_SMS / 2;
  dim3 grid(2 * n_clusters);
#define PM_T3_LAUNCH(ACTV)                                                                                                  \
  case ACTV: {                                                                                                              \
    static bool attr_set = false;                                                                                           \
    if (!attr_set) {                                                                                                        \
      cudaError_t e1 = cudaFuncSetAttribute(encoder_fwd_tc3<ACTV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S_TOTAL);  \
      cudaError_t e2 = cudaFuncSetAttribute(encoder_fwd_tc3<ACTV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S_TOTAL); \
      if (e1 != cudaSuccess || e2 != cudaSuccess) PM_FAIL(PM_ERR_CUDA, "cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2)); \
      attr_set = true;                                                                                                      \
    }                                                                                                                       \
    if (argmax)                                                                                                             \
      encoder_fwd_tc3<ACTV, true><<<grid, T3_THREADS, S_TOTAL, st>>>(x, ldx, B, N, C, x_tma, wimg, p->W1, p->b1, p->b2, p->b3, feat, ldf, argmax, sink, scratch); \
    else                                                                                                                    \
      encoder_fwd_tc3<ACTV, false><<<grid, T3_THREADS, S_TOTAL, st>>>(x, ldx, B, N, C, x_tma, wimg, p->W1, p->b1, p->b2, p->b3, feat, ldf, nullptr, sink, scratch); \
  } break;
  switch (act) {
    PM_T3_LAUNCH(PM_ACT_TANH)
    PM_T3_LAUNCH(PM_ACT_RELU)
    PM_T3_LAUNCH(PM_ACT_ELU)
    PM_T3_LAUNCH(PM_ACT_SELU)
    PM_T3_LAUNCH(PM_ACT_LRELU)
    PM_T3_LAUNCH(PM_ACT_SIGMOID)
    PM_T3_LAUNCH(PM_ACT_NONE)
    default: PM_FAIL(PM_ERR_ARG, "split-fp16 encoder: activation %d", act);
  }
#undef PM_T3_LAUNCH
  PM_CHECK_LAUNCH("pm_pointnet_encode_forward_tc3");
  return PM_OK;
}

// diagnostic: protocol error word of the last launch in `ws` (0 = clean); synchronises the stream
int pm_pointnet_tc3_last_error(const void* ws, pm_stream_t s) {
  int32_t h = -1;
  cudaMemcpyAsync(&h, reinterpret_cast<const uint8_t*>(ws) + 2 * WIMG_PER_RANK, sizeof(int32_t), cudaMemcpyDeviceToHost, pm_st(s));
  cudaStreamSynchronize(pm_st(s));
  return h;
}

}  // extern "C"
