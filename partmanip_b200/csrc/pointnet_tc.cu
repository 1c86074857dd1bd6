// K1 (bf16) — PointNet encoder forward on the 5th-gen tensor cores: tcgen05.mma, accumulators in TMEM,
// CTA pairs (cta_group::2) so that the two big weight matrices stay RESIDENT in shared memory.
// reference: algorithms/algo_utils/network.py:148-150 (per-point Linear-act-Linear-act-Linear) + :182 (max over points).
//
// Why a CTA pair: W2 (256x128) + W3 (512x256) are 320 KB in bf16 — more than one SM's 227 KB — and streaming W3
// from L2 once per 128-point tile would need ~2x the L2 bandwidth the chip has.  With cta_group::2 each CTA holds
// HALF of W3 (its 256 output channels, 128 KB) and half of W2 (128 channels, 32 KB); the hardware shares the
// per-CTA operand halves, so nothing is replicated and nothing is re-read from L2/HBM after the prologue.
//
// Per tile of 256 points (128 per CTA); 17 warps = 4 epilogue groups of 4 warps (one warp of each group per scheduler) + 1
// MMA-issue warp; every MMA is M=256 (pair), N=256, K=16:
//   L1  groups A0,A1 : h1 = act(W1 x + b1)  (K = C <= 4: not a GEMM), 64 channels per group, computed into registers while the
//                      previous tile's last MMAs run                          -> smem, bf16, 128B-swizzled K-major (A operand)
//   L2  tcgen05      : D2[256 pts x 256 ch] = H1 . W2^T (8 MMAs)               -> TMEM region p (p = tile parity)
//   E2  all 4 groups : tcgen05.ld -> +b2 -> act -> bf16 -> smem H2 [128 pts x 256] (B operand of L3; overlays H1).  Group g takes
//                      the g-th 16-column slice of EVERY 64-channel k-block, so k-blocks complete in order and ...
//   L3  tcgen05      : ... D3^T[256 ch x 256 pts] = W3 . H2^T of channel chunk 0 is issued k-block by k-block BEHIND E2 into
//                      region p^1; chunk 1 follows into region p, which E2 has drained by then (16 MMAs each)
//   E3  groups B0,B1 : lane = channel, columns = points (B0 columns [0,128), B1 [128,256)): the transposed orientation makes the
//                      symmetric max-pool a per-thread FMNMX3 tree over TMEM columns; the argmax rides in the low 4 mantissa
//                      bits of the compared value; the two column halves are merged once per cloud.
// TMEM/b2/x loads are software-pipelined one step ahead.  Per-role clock64 stamps: make EXTRA=-DPM_TC_TIMING + scripts/tc_timing.py.
// Round-2 experiments that did NOT help (both parity-green): groups A alone running the layer-2 epilogue (0.517 ms, same: the epilogue is
// MUFU-bound whoever runs it), and half of the next tile's layer 1 computed inside the accumulator wait (0.552 ms, slower).
// The (points x 128/256/512) activations never leave the SM; HBM sees 4C bytes per point in and 4 KB per cloud out.
#include "tc_common.cuh"

namespace {
using namespace pmtc;

constexpr int TC_THREADS = 544;          // warps 0-3 A0 | 4-7 A1 | 8-11 B0 | 12-15 B1 (epilogue groups, 4 per scheduler) | warp 16: TMEM alloc + MMA issue
constexpr int PTS_PER_CTA = 128;
constexpr int PTS_PER_TILE = 256;

// ---- shared-memory map (bytes); every operand base is 1024-aligned (SWIZZLE_128B atoms)
constexpr uint32_t SM_W3 = 0;            // 2 chunks x 4 k-blocks x (128 rows x 128 B)      = 131072
constexpr uint32_t SM_W2 = 131072;       // 2 k-blocks x (128 rows x 128 B)                 =  32768
constexpr uint32_t SM_H = 163840;        // H2: 4 k-blocks x (128 rows x 128 B) = 65536; H1 overlays the first 32768
constexpr uint32_t SM_W1 = 229376;       // 128 x 4 fp32                                    =   2048
constexpr uint32_t SM_B1 = 231424;       // 128 fp32                                        =    512
constexpr uint32_t SM_BAR = 231936;      // 13 mbarriers (104 B) + tmem base (4 B, at +112)
constexpr uint32_t SM_TOTAL = 232064;
constexpr uint32_t KBLOCK_BYTES = 128 * 128;   // one 64-wide k-block of a 128-row operand

enum { BAR_H1_FULL = 0, BAR_ACC2_FULL, BAR_H2_KB0, BAR_H2_KB1, BAR_H2_KB2, BAR_H2_KB3, BAR_G_FULL0, BAR_G_FULL1,
       BAR_G_EMPTY0, BAR_G_EMPTY1, BAR_L3_DONE, BAR_W_LOADED, NUM_BARS };
// TMEM plan: two 256-column regions that swap roles every tile (p = tile parity):
//   region p   : layer-2 accumulator of this tile, then (once E2 has read it) layer-3 group 1 (channel chunk 1)
//   region p^1 : layer-3 group 0 (channel chunk 0) of this tile; it held group 1 of the previous tile
// Both layer-3 groups are N=256 MMAs (all 256 points of the pair tile): half the shared-memory operand traffic per
// flop of the N=128 shape, which ran at ~70 % of the tensor rate because the W3 operand stream saturated SMEM.
__host__ __device__ constexpr uint32_t region_col(int p) { return p ? 256u : 0u; }

// packed weight image per network, per CTA rank: [W3 half (131072) | W2 half (32768)] ready to memcpy into smem
constexpr size_t WPACK_PER_RANK = 131072 + 32768;

// ------------------------------------------------------------------------------------------------ weight packing
// fp32 W2 (256,128) / W3 (512,256) -> bf16 smem images for CTA rank 0 and 1
__global__ void pack_weights_kernel(const float* __restrict__ W2, const float* __restrict__ W3, uint8_t* __restrict__ out) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthr = gridDim.x * blockDim.x;
  // W3: rank r, chunk j, k-block kb, row m, chunk c8 -> channel r*256 + j*128 + m, k = kb*64 + c8*8 .. +8
  for (int i = tid; i < 2 * 2 * 4 * 128 * 8; i += nthr) {
    const int c8 = i & 7, m = (i >> 3) & 127, kb = (i >> 10) & 3, j = (i >> 12) & 1, r = i >> 13;
    const float* src = W3 + (size_t)(r * 256 + j * 128 + m) * 256 + kb * 64 + c8 * 8;
    uint4 v;
    v.x = pack_bf16(src[0], src[1]); v.y = pack_bf16(src[2], src[3]);
    v.z = pack_bf16(src[4], src[5]); v.w = pack_bf16(src[6], src[7]);
    uint8_t* dst = out + (size_t)r * WPACK_PER_RANK + SM_W3 + j * 65536 + kb * KBLOCK_BYTES + (m * 128 + ((c8 ^ (m & 7)) << 4));
    *reinterpret_cast<uint4*>(dst) = v;
  }
  // W2: rank r holds output channels r*128 + n (the B operand's N half): k-block kb, row n, chunk c8
  for (int i = tid; i < 2 * 2 * 128 * 8; i += nthr) {
    const int c8 = i & 7, n = (i >> 3) & 127, kb = (i >> 10) & 1, r = i >> 11;
    const float* src = W2 + (size_t)(r * 128 + n) * 128 + kb * 64 + c8 * 8;
    uint4 v;
    v.x = pack_bf16(src[0], src[1]); v.y = pack_bf16(src[2], src[3]);
    v.z = pack_bf16(src[4], src[5]); v.w = pack_bf16(src[6], src[7]);
    uint8_t* dst = out + (size_t)r * WPACK_PER_RANK + 131072 + kb * KBLOCK_BYTES + (n * 128 + ((c8 ^ (n & 7)) << 4));
    *reinterpret_cast<uint4*>(dst) = v;
  }
}

// ------------------------------------------------------------------------------------------------ the encoder

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
// max of 32 values: 16 three-input FMNMX in 4 independent chains
__device__ __forceinline__ float max32(const float (&k)[32]) {
  float m0 = fmax3(k[0], k[1], k[2]), m1 = fmax3(k[3], k[4], k[5]), m2 = fmax3(k[6], k[7], k[8]), m3 = fmax3(k[9], k[10], k[11]);
  m0 = fmax3(m0, k[12], k[13]); m1 = fmax3(m1, k[14], k[15]); m2 = fmax3(m2, k[16], k[17]); m3 = fmax3(m3, k[18], k[19]);
  m0 = fmax3(m0, k[20], k[21]); m1 = fmax3(m1, k[22], k[23]); m2 = fmax3(m2, k[24], k[25]); m3 = fmax3(m3, k[26], k[27]);
  m0 = fmax3(m0, k[28], k[29]); m1 = fmax3(m1, k[30], k[31]);
  return fmax3(m0, m1, fmaxf(m2, m3));
}

// max of 16 values: 8 FMNMX(3)
__device__ __forceinline__ float max16(const float (&k)[16]) {
  const float m0 = fmax3(k[0], k[1], k[2]), m1 = fmax3(k[3], k[4], k[5]), m2 = fmax3(k[6], k[7], k[8]);
  const float m3 = fmax3(k[9], k[10], k[11]), m4 = fmax3(k[12], k[13], k[14]);
  return fmaxf(fmax3(m0, m1, m2), fmax3(m3, m4, k[15]));
}

// layer 1 for the 16-byte chunks [C0, C1) of this thread's point row: h[(c-C0)*4 + q] = bf16x2 of channels 8c+2q, 8c+2q+1
template <int ACT, int C0, int C1>
__device__ __forceinline__ void layer1_part(const float (&xv)[4], const float* sW1, const float* sB1, uint32_t (&h)[(C1 - C0) * 4]) {
#pragma unroll
  for (int q = C0 * 4; q < C1 * 4; ++q) {
    const float4 w0 = *reinterpret_cast<const float4*>(sW1 + (2 * q) * 4);
    const float4 w1 = *reinterpret_cast<const float4*>(sW1 + (2 * q + 1) * 4);
    const float a0 = fmaf(xv[3], w0.w, fmaf(xv[2], w0.z, fmaf(xv[1], w0.y, fmaf(xv[0], w0.x, sB1[2 * q]))));
    const float a1 = fmaf(xv[3], w1.w, fmaf(xv[2], w1.z, fmaf(xv[1], w1.y, fmaf(xv[0], w1.x, sB1[2 * q + 1]))));
    h[q - C0 * 4] = pack_bf16(act_fast<ACT>(a0), act_fast<ACT>(a1));
  }
}
template <int C0, int C1>
__device__ __forceinline__ void store_h1_part(uint8_t* smem, int row, const uint32_t (&h)[(C1 - C0) * 4]) {
#pragma unroll
  for (int c16 = C0; c16 < C1; ++c16) {
    const uint32_t off = SM_H + (c16 >> 3) * KBLOCK_BYTES + sw128(row, c16 & 7);
    *reinterpret_cast<uint4*>(smem + off) =
        make_uint4(h[4 * (c16 - C0)], h[4 * (c16 - C0) + 1], h[4 * (c16 - C0) + 2], h[4 * (c16 - C0) + 3]);
  }
}
__device__ __forceinline__ void load_point(const float* __restrict__ x, int64_t ldx, int C, int b, int pt, float (&xv)[4]) {
  const float* xp = x + (int64_t)b * ldx + (int64_t)pt * C;
  xv[0] = __ldg(xp);
  xv[1] = C > 1 ? __ldg(xp + 1) : 0.f;
  xv[2] = C > 2 ? __ldg(xp + 2) : 0.f;
  xv[3] = C > 3 ? __ldg(xp + 3) : 0.f;
}

// Roles (288 threads): warps 0-3 "A": layer 1 (80 of 128 channels) + layer-2 epilogue | warps 4-7 "B": layer-3 epilogue
// (max-pool) + the other 48 layer-1 channels of the NEXT tile | warp 8: TMEM alloc + MMA issue (leader CTA only).
// Pipeline per 256-point tile (t):   H1(t) -> L2 MMA -> E2 writes H2 one 64-channel k-block at a time; the layer-3 MMAs of
// the first channel chunk (both point halves, two TMEM buffers) are issued k-block by k-block BEHIND E2, so 2 of the 4
// MMA groups run under the epilogue; the second chunk's groups reuse the buffers as soon as E3 has drained them.
template <int ACT, bool WANT_ARGMAX>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
encoder_fwd_tc(const float* __restrict__ x, int64_t ldx, int B, int N, int C, const uint8_t* __restrict__ wpack,
               const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ b2,
               const float* __restrict__ b3, float* __restrict__ feat, int64_t ldf,
               int32_t* __restrict__ argmax, ErrSink err, uint8_t* __restrict__ scratch) {
#ifdef PM_TC_TIMING
  long long* dbg = reinterpret_cast<long long*>(err.last + 16);
#define TSTAMP(slot) do { if (blockIdx.x < 2 && it == 3) dbg[blockIdx.x * 64 + (slot)] = clock64(); } while (0)
#else
#define TSTAMP(slot) do { } while (0)
#endif
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int tpc = N / PTS_PER_TILE;                                   // tiles per cloud
  const int n_clouds = cluster_id < B ? (B - cluster_id + n_clusters - 1) / n_clusters : 0;
  const int n_tiles = n_clouds * tpc;
  float* sW1 = reinterpret_cast<float*>(smem + SM_W1);
  float* sB1 = reinterpret_cast<float*>(smem + SM_B1);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 112);
  auto bar = [&](int i) { return sbase + SM_BAR + 8u * i; };
  auto tile_cloud = [&](int it) { return cluster_id + (it / tpc) * n_clusters; };

  // ---------------- prologue: resident weights, barriers, TMEM
  if ((sbase & 1023u) != 0 && tid == 0) err_report(err, 900);
  {
    for (int i = tid; i < 128 * 4; i += TC_THREADS) sW1[i] = ((i & 3) < C) ? W1[(i >> 2) * C + (i & 3)] : 0.f;
    if (tid < 128) sB1[tid] = b1[tid];
  }
  if (tid == 0) {
    // the CTA's resident weight image (160 KB) comes in through the TMA engine: ten 16 KB bulk copies on one mbarrier
    mbar_init(bar(BAR_W_LOADED), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar(BAR_W_LOADED), (uint32_t)WPACK_PER_RANK);
    const uint8_t* src = wpack + (size_t)rank * WPACK_PER_RANK;
    for (uint32_t off = 0; off < (uint32_t)WPACK_PER_RANK; off += 16384u) bulk_g2s(sbase + off, src + off, 16384u, bar(BAR_W_LOADED));
  }
  if (tid == 0) {
    mbar_init(bar(BAR_H1_FULL), 4);
    mbar_init(bar(BAR_ACC2_FULL), 1);
    for (int kb = 0; kb < 4; ++kb) mbar_init(bar(BAR_H2_KB0 + kb), 8);    // 4 epilogue groups x 2 CTAs
    for (int i = 0; i < 2; ++i) { mbar_init(bar(BAR_G_FULL0 + i), 1); mbar_init(bar(BAR_G_EMPTY0 + i), 4); }
    mbar_init(bar(BAR_L3_DONE), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 16) {   // one warp per CTA allocates all 512 TMEM columns for the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  bool ok = mbar_wait(bar(BAR_W_LOADED), 0, err, 110);          // weight images landed (async proxy wrote them; the MMA reads them via the same proxy)
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- layer-2 epilogue: acc row -> +b2 -> act -> bf16 -> smem.  Each of the four epilogue groups takes the SAME 16-column
  //      slice (slice = group index) of every 64-channel k-block of H2, so the k-blocks complete one after the other (not all
  //      at the end) and the layer-3 MMAs of group 0 follow the epilogue k-block by k-block; four warps per scheduler
  //      interleave their MUFU.TANH / FADD / pack / store streams.  The tcgen05.ld and the b2 loads (from L2: the 227 KB
  //      carve-out leaves almost no L1) of k-block kb+1 are in flight while k-block kb is processed.
  auto load_b2 = [&](float4 (&bb)[4], int col) {
#pragma unroll
    for (int q = 0; q < 4; ++q) bb[q] = __ldg(reinterpret_cast<const float4*>(b2 + col) + q);
  };
  auto e2_slices = [&](uint32_t taddr, int row, int sc, const float4 (&bfirst)[4], int bar_id, bool elect) {
    auto e2_sub = [&](const uint32_t (&v)[16], const float4 (&bb)[4], int kb) {
      uint32_t pk[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float a0 = __uint_as_float(v[4 * q]) + bb[q].x, a1 = __uint_as_float(v[4 * q + 1]) + bb[q].y;
        const float a2 = __uint_as_float(v[4 * q + 2]) + bb[q].z, a3 = __uint_as_float(v[4 * q + 3]) + bb[q].w;
        pk[2 * q] = pack_bf16(act_fast<ACT>(a0), act_fast<ACT>(a1));
        pk[2 * q + 1] = pack_bf16(act_fast<ACT>(a2), act_fast<ACT>(a3));
      }
#pragma unroll
      for (int q = 0; q < 2; ++q)
        *reinterpret_cast<uint4*>(smem + SM_H + kb * KBLOCK_BYTES + sw128(row, 2 * sc + q)) =
            make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
    };
    auto release = [&](int kb, bool last) {                    // this group's slice of k-block kb is in smem
      if (last) tc_fence_before();
      fence_proxy_async();
      named_bar_sync(bar_id, 128);
      if (elect) mbar_arrive_cluster(bar(BAR_H2_KB0 + kb), 0);
    };
    const uint32_t t0 = taddr + sc * 16;
    uint32_t v0[16], v1[16];
    float4 b0[4], b1[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) b0[q] = bfirst[q];
    tmem_ld16(t0, v0);
    tmem_ld_wait();
    tmem_ld16(t0 + 64, v1);  load_b2(b1, 64 + sc * 16);  e2_sub(v0, b0, 0); tmem_ld_wait(); release(0, false);
    tmem_ld16(t0 + 128, v0); load_b2(b0, 128 + sc * 16); e2_sub(v1, b1, 1); tmem_ld_wait(); release(1, false);
    tmem_ld16(t0 + 192, v1); load_b2(b1, 192 + sc * 16); e2_sub(v0, b0, 2); tmem_ld_wait(); release(2, false);
    e2_sub(v1, b1, 3);
    release(3, true);
  };

  const int grp = warp >> 2;                                  // 0,1: A0,A1 | 2,3: B0,B1 | 4: MMA warp
  const int row = (warp & 3) * 32 + lane;                     // TMEM lane == point row (layers 1/2) == channel in chunk (layer 3)
  const bool elect = (tid & 127) == 0;
  const uint32_t lane_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);

  if (grp < 2) {
    // =========================================================== groups A0/A1: layer 1 (64 channels each) + layer-2 epilogue of
    //                                                               k-block grp
    float xn[4] = {0.f, 0.f, 0.f, 0.f};                       // next tile's point, fetched one tile ahead (HBM latency)
    if (n_tiles > 0) load_point(x, ldx, C, tile_cloud(0), rank * PTS_PER_CTA + row, xn);
    for (int it = 0; it < n_tiles && ok; ++it) {
      if (tid == 0) TSTAMP(0);
      const float xv[4] = {xn[0], xn[1], xn[2], xn[3]};
      if (it + 1 < n_tiles) load_point(x, ldx, C, tile_cloud(it + 1), ((it + 1) % tpc) * PTS_PER_TILE + rank * PTS_PER_CTA + row, xn);
      // layer 1 into registers while the previous tile's last layer-3 MMAs run
      uint32_t h1[32];
      if (grp == 0) layer1_part<ACT, 0, 8>(xv, sW1, sB1, h1); else layer1_part<ACT, 8, 16>(xv, sW1, sB1, h1);
      if (tid == 0) TSTAMP(1);
      // k-blocks 0-1 of H (= H1) are still being read by the first half of the previous tile's last layer-3 group
      if (it > 0) ok = mbar_wait(bar(BAR_L3_DONE), (it - 1) & 1, err, 101);
      if (!ok) break;
      if (tid == 0) TSTAMP(2);
      if (grp == 0) store_h1_part<0, 8>(smem, row, h1); else store_h1_part<8, 16>(smem, row, h1);
      fence_proxy_async();
      named_bar_sync(1 + grp, 128);
      if (elect) mbar_arrive_cluster(bar(BAR_H1_FULL), 0);
      if (tid == 0) TSTAMP(3);
      float4 bf[4];                                           // b2 of the first step, in flight across the ACC2_FULL wait
      load_b2(bf, grp * 16);
      ok = mbar_wait(bar(BAR_ACC2_FULL), it & 1, err, 102);
      if (!ok) break;
      tc_fence_after();
      if (tid == 0) TSTAMP(4);
      e2_slices(lane_taddr + region_col(it & 1), row, grp, bf, 1 + grp, elect);
      if (tid == 0) TSTAMP(5);
    }
  } else if (grp < 4) {
    // =========================================================== groups B0/B1: their slices of the layer-2 epilogue (thread =
    //          point row), then the layer-3 epilogue (thread = channel): B0 reduces columns [0,128) of BOTH channel-chunk
    //          accumulators, B1 columns [128,256); the two partial (max, argmax) pairs are merged once per cloud
    const int g = grp - 2;
    float best[2] = {-INFINITY, -INFINITY};                   // running max (key when WANT_ARGMAX: low 4 bits = 15 - column)
    int bestp[2] = {0, 0};                                    // point index of column 0 of the winning 16-column group
    uint4* merge = reinterpret_cast<uint4*>(scratch) + (size_t)blockIdx.x * 128;
    for (int it = 0; it < n_tiles && ok; ++it) {
      const int b = tile_cloud(it), j = it % tpc;
      {
        float4 bf[4];
        load_b2(bf, grp * 16);
        if (tid == 256) TSTAMP(20);
        ok = mbar_wait(bar(BAR_ACC2_FULL), it & 1, err, 109);
        if (!ok) break;
        tc_fence_after();
        e2_slices(lane_taddr + region_col(it & 1), row, grp, bf, 1 + grp, elect);
        if (tid == 256) TSTAMP(21);
      }
#pragma unroll
      for (int G = 0; G < 2; ++G) {                           // layer-3 group G = channel chunk G
        if (tid == 256) TSTAMP(8 + 3 * G);
        ok = mbar_wait(bar(BAR_G_FULL0 + G), it & 1, err, 103);
        if (!ok) break;
        tc_fence_after();
        if (tid == 256) TSTAMP(9 + 3 * G);
        float bv = best[G];
        int bp = bestp[G];
        {
          auto e3_sub = [&](uint32_t (&v)[16], int sc) {        // 16 columns (points) of this thread's channel
            float k[16];
            if (WANT_ARGMAX) {
              // key = value with its low 4 mantissa bits replaced by (15 - column): one FMNMX tree yields max AND position
              // (values closer than 2^-19 relative may swap order — far below the bf16 operand rounding)
#pragma unroll
              for (int i = 0; i < 16; ++i) k[i] = __uint_as_float((v[i] & 0xFFFFFFF0u) | (uint32_t)(15 - i));
              const float m = max16(k);
              // column n -> point: columns [0,128) are CTA 0's rows, [128,256) CTA 1's (the B operand's N halves)
              if (m > bv) { bv = m; bp = j * PTS_PER_TILE + g * 128 + sc * 16; }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) k[i] = __uint_as_float(v[i]);
              bv = fmaxf(bv, max16(k));
            }
          };
          // group 0 lives in region p^1, group 1 in region p; this epilogue group takes columns [g*128, g*128+128)
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
        if (elect) mbar_arrive_cluster(bar(BAR_G_EMPTY0 + G), 0);
        if (tid == 256) TSTAMP(10 + 3 * G);
      }
      if (!ok) break;
      if (j == tpc - 1) {                                     // cloud complete: merge the two column halves, pooled outputs
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
            const bool take = ob[G] > best[G];
            const float bv = take ? ob[G] : best[G];
            const int bp = take ? op[G] : bestp[G];
            const int ch = rank * 256 + G * 128 + row;
            const uint32_t kb = __float_as_uint(bv);
            if (WANT_ARGMAX) {
              feat[(int64_t)b * ldf + ch] = __uint_as_float(kb & 0xFFFFFFF0u) + __ldg(b3 + ch);   // (bias after the pool)
              argmax[(int64_t)b * 512 + ch] = bp + 15 - (int)(kb & 15u);
            } else {
              feat[(int64_t)b * ldf + ch] = bv + __ldg(b3 + ch);
            }
          }
        }
#pragma unroll
        for (int G = 0; G < 2; ++G) { best[G] = -INFINITY; bestp[G] = 0; }
      }
    }
  } else if (rank == 0) {
    // =========================================================== warp 16 of the leader CTA: MMA issue
    const uint32_t idesc = umma_idesc(256, 256);             // every MMA of this kernel: M=256 (pair), N=256, K=16
    auto l3_mma = [&](int chunk, uint32_t dcol, int k, bool acc) {   // one K=16 step of layer-3 channel chunk `chunk`
      const uint32_t koff = (k >> 2) * KBLOCK_BYTES + (k & 3) * 32;
      umma_bf16_2cta(tmem_base + dcol, umma_desc(sbase + SM_W3 + chunk * 65536 + koff), umma_desc(sbase + SM_H + koff), idesc,
                     acc ? 1u : 0u);
    };
    for (int it = 0; it < n_tiles && ok; ++it) {
      const uint32_t colp = region_col(it & 1), colq = region_col((it & 1) ^ 1);
      if (lane == 0) TSTAMP(24);
      // region p held layer-3 group 0 of the previous tile: E3 must have drained it before layer 2 overwrites it
      ok = mbar_wait(bar(BAR_H1_FULL), it & 1, err, 104) && mbar_wait(bar(BAR_G_EMPTY0), (it & 1) ^ 1, err, 108);
      if (!ok) break;
      tc_fence_after();
      if (lane == 0) TSTAMP(25);
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {                         // K = 128 = 2 k-blocks x 4 x UMMA_K(16)
          const uint32_t koff = (k >> 2) * KBLOCK_BYTES + (k & 3) * 32;
          umma_bf16_2cta(tmem_base + colp, umma_desc(sbase + SM_H + koff), umma_desc(sbase + SM_W2 + koff), idesc, k > 0);
        }
        umma_commit_mc(bar(BAR_ACC2_FULL));
      }
      __syncwarp();
      if (lane == 0) TSTAMP(26);
      // ---- group 0 (channel chunk 0) into region p^1 (held group 1 of the previous tile), k-block by k-block behind E2
      ok = mbar_wait(bar(BAR_G_EMPTY1), (it & 1) ^ 1, err, 106);
      if (!ok) break;
#pragma unroll 1
      for (int q = 0; q < 4 && ok; ++q) {
        const int kb = q;                                      // (each k-block comes from its own epilogue group)
        ok = mbar_wait(bar(BAR_H2_KB0 + kb), it & 1, err, 105);
        if (!ok) break;
        tc_fence_after();
        if (lane == 0) TSTAMP(27 + q);
        if (lane == 0) {
          for (int k = 0; k < 4; ++k) l3_mma(0, colq, kb * 4 + k, q > 0 || k > 0);
          if (q == 3) umma_commit_mc(bar(BAR_G_FULL0));
        }
        __syncwarp();
      }
      if (!ok) break;
      // ---- group 1 (channel chunk 1) into region p: E2 is done with the layer-2 accumulator (all four k-blocks arrived)
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          l3_mma(1, colp, k, k > 0);
          // the group walks K in order: after its 8th step nothing reads k-blocks 0-1 of H any more — exactly where the next
          // tile's H1 lives — so the next tile's layer 1 can be stored (and its layer-2 MMAs queued right behind this group)
          // about 1 k cycles before the group completes
          if (k == 7) umma_commit_mc(bar(BAR_L3_DONE));
        }
        umma_commit_mc(bar(BAR_G_FULL1));
      }
      __syncwarp();
      if (lane == 0) TSTAMP(32);
    }
  }

  // ---------------- teardown: everyone done with TMEM in both CTAs, then free it
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 16) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

extern "C" {

int pm_has_tcgen05(void) { return 1; }

// workspace: two packed weight images (rank 0 / rank 1) + an error word
size_t pm_pointnet_encode_forward_tc_ws_bytes(int, int, int) { return 2 * WPACK_PER_RANK + 4096 + (size_t)PM_NUM_SMS * 2048; }

int pm_pointnet_encode_forward_tc(const float* x, int64_t ldx, int B, int N, int C, const pm_encoder_params* p, int act,
                                  float* feat, int64_t ldf, int32_t* argmax, void* ws, size_t ws_bytes, pm_stream_t s) {
  PM_REQUIRE(C >= 1 && C <= 4, PM_ERR_UNSUPPORTED, "bf16 encoder: C=%d channels per point (supports 1..4; use PM_PREC_FP32)", C);
  PM_REQUIRE(N % PTS_PER_TILE == 0, PM_ERR_UNSUPPORTED, "bf16 encoder: N=%d must be a multiple of %d (use PM_PREC_FP32)", N, PTS_PER_TILE);
  PM_REQUIRE(ws && ws_bytes >= pm_pointnet_encode_forward_tc_ws_bytes(B, N, C), PM_ERR_ARG, "bf16 encoder: workspace too small");
  PM_REQUIRE(pm_aligned(ws, 256), PM_ERR_ALIGN, "bf16 encoder: workspace must be 256-byte aligned");
  PM_REQUIRE(pm_aligned(p->b2, 16), PM_ERR_ALIGN, "bf16 encoder: b2 must be 16-byte aligned");
  cudaStream_t st = pm_st(s);
  uint8_t* wpack = reinterpret_cast<uint8_t*>(ws);
  int32_t* err = reinterpret_cast<int32_t*>(wpack + 2 * WPACK_PER_RANK);
  cudaMemsetAsync(err, 0, sizeof(int32_t), st);
  const ErrSink sink{err, pm_tc_sticky_word()};
  pack_weights_kernel<<<64, 256, 0, st>>>(p->W2, p->W3, wpack);
  int n_clusters = B < PM_NUM_SMS / 2 ? B : PM_NUM_SMS / 2;
  dim3 grid(2 * n_clusters);
#define PM_TC_LAUNCH(ACTV)                                                                                                  \
  case ACTV: {                                                                                                              \
    static bool attr_set = false;                                                                                           \
    if (!attr_set) {                                                                                                        \
      cudaError_t e1 = cudaFuncSetAttribute(encoder_fwd_tc<ACTV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL);  \
      cudaError_t e2 = cudaFuncSetAttribute(encoder_fwd_tc<ACTV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL); \
      if (e1 != cudaSuccess || e2 != cudaSuccess) PM_FAIL(PM_ERR_CUDA, "cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2)); \
      attr_set = true;                                                                                                      \
    }                                                                                                                       \
    if (argmax)                                                                                                             \
      encoder_fwd_tc<ACTV, true><<<grid, TC_THREADS, SM_TOTAL, st>>>(x, ldx, B, N, C, wpack, p->W1, p->b1, p->b2, p->b3, feat, ldf, argmax, sink, wpack + 2 * WPACK_PER_RANK + 4096); \
    else                                                                                                                    \
      encoder_fwd_tc<ACTV, false><<<grid, TC_THREADS, SM_TOTAL, st>>>(x, ldx, B, N, C, wpack, p->W1, p->b1, p->b2, p->b3, feat, ldf, nullptr, sink, wpack + 2 * WPACK_PER_RANK + 4096); \
  } break;
  switch (act) {
    PM_TC_LAUNCH(PM_ACT_TANH)
    PM_TC_LAUNCH(PM_ACT_RELU)
    PM_TC_LAUNCH(PM_ACT_ELU)
    PM_TC_LAUNCH(PM_ACT_SELU)
    PM_TC_LAUNCH(PM_ACT_LRELU)
    PM_TC_LAUNCH(PM_ACT_SIGMOID)
    PM_TC_LAUNCH(PM_ACT_NONE)
    default: PM_FAIL(PM_ERR_ARG, "bf16 encoder: activation %d", act);
  }
#undef PM_TC_LAUNCH
  PM_CHECK_LAUNCH("pm_pointnet_encode_forward_tc");
  return PM_OK;
}

// test/diagnostic hook: the error word the last launch left in `ws` (0 = clean); synchronises the stream
int pm_pointnet_tc_last_error(const void* ws, pm_stream_t s) {
  int32_t h = -1;
  cudaMemcpyAsync(&h, reinterpret_cast<const uint8_t*>(ws) + 2 * WPACK_PER_RANK, sizeof(int32_t), cudaMemcpyDeviceToHost, pm_st(s));
  cudaStreamSynchronize(pm_st(s));
  return h;
}

}  // extern "C"
