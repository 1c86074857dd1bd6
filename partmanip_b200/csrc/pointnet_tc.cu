// K1 (bf16) — PointNet encoder forward on the 5th-gen tensor cores: tcgen05.mma, accumulators in TMEM,
// CTA pairs (cta_group::2) so that the two big weight matrices stay RESIDENT in shared memory.
// reference: algorithms/algo_utils/network.py:148-150 (per-point Linear-act-Linear-act-Linear) + :182 (max over points).
//
// Why a CTA pair: W2 (256x128) + W3 (512x256) are 320 KB in bf16 — more than one SM's 227 KB — and streaming W3
// from L2 once per 128-point tile would need ~2x the L2 bandwidth the chip has.  With cta_group::2 each CTA holds
// HALF of W3 (its 256 output channels, 128 KB) and half of W2 (128 channels, 32 KB); the hardware shares the
// per-CTA operand halves, so nothing is replicated and nothing is re-read from L2/HBM after the prologue.
//
// Per tile of 256 points (128 per CTA):
//   L1  CUDA cores : h1 = act(W1 x + b1)  (K = C <= 4: not a GEMM)       -> smem, bf16, 128B-swizzled K-major (A operand)
//   L2  tcgen05    : D2[256 pts x 256 ch] = H1 . W2^T   (M=256 N=256 K=128, 8 MMAs)   acc in TMEM cols [0,256)
//   E2  4 warps    : tcgen05.ld -> +b2 -> act -> bf16 -> smem H2 [128 pts x 256] (B operand of L3; overlays H1)
//   L3  tcgen05    : D3^T[256 ch x 128 pts] = W3 . H2^T for (2 channel chunks) x (2 point halves), 16 MMAs each,
//                    double-buffered in TMEM cols [256,384) / [384,512)
//   E3  4 warps    : lane = channel, columns = points: running max (+ first-index argmax) entirely in registers —
//                    the transposed orientation makes the symmetric max-pool a per-thread reduction, no shuffles.
// The (points x 128/256/512) activations never leave the SM; HBM sees 4C bytes per point in and 4 KB per cloud out.
#include "tc_common.cuh"

namespace {
using namespace pmtc;

constexpr int TC_THREADS = 288;          // warps 0-3: L1 + E2 | warps 4-7: E3 | warp 8: TMEM alloc + MMA issue
constexpr int PTS_PER_CTA = 128;
constexpr int PTS_PER_TILE = 256;

// ---- shared-memory map (bytes); every operand base is 1024-aligned (SWIZZLE_128B atoms)
constexpr uint32_t SM_W3 = 0;            // 2 chunks x 4 k-blocks x (128 rows x 128 B)      = 131072
constexpr uint32_t SM_W2 = 131072;       // 2 k-blocks x (128 rows x 128 B)                 =  32768
constexpr uint32_t SM_H = 163840;        // H2: 4 k-blocks x (128 rows x 128 B) = 65536; H1 overlays the first 32768
constexpr uint32_t SM_W1 = 229376;       // 128 x 4 fp32                                    =   2048
constexpr uint32_t SM_B1 = 231424;       // 128 fp32                                        =    512
constexpr uint32_t SM_BAR = 231936;      // 8 mbarriers (64 B) + tmem base (4 B)
constexpr uint32_t SM_TOTAL = 232064;
constexpr uint32_t KBLOCK_BYTES = 128 * 128;   // one 64-wide k-block of a 128-row operand

enum { BAR_H1_FULL = 0, BAR_ACC2_FULL, BAR_H2_FULL, BAR_ACC3_FULL0, BAR_ACC3_FULL1, BAR_ACC3_EMPTY0, BAR_ACC3_EMPTY1,
       BAR_L3_DONE, NUM_BARS };

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
template <int ACT, bool WANT_ARGMAX>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
encoder_fwd_tc(const float* __restrict__ x, int64_t ldx, int B, int N, int C, const uint8_t* __restrict__ wpack,
               const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ b2,
               const float* __restrict__ b3, float* __restrict__ feat, int64_t ldf,
               int32_t* __restrict__ argmax, int32_t* __restrict__ err) {
#ifdef PM_TC_TIMING
  long long* dbg = reinterpret_cast<long long*>(err + 16);
#define TSTAMP(slot) do { if (blockIdx.x < 2 && it == 3) dbg[blockIdx.x * 64 + (slot)] = clock64(); } while (0)
#else
#define TSTAMP(slot) do { } while (0)
#endif
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int tiles_per_cloud = N / PTS_PER_TILE;
  float* sW1 = reinterpret_cast<float*>(smem + SM_W1);
  float* sB1 = reinterpret_cast<float*>(smem + SM_B1);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 64);
  auto bar = [&](int i) { return sbase + SM_BAR + 8u * i; };

  // ---------------- prologue: resident weights, barriers, TMEM
  if ((sbase & 1023u) != 0 && tid == 0) atomicExch(err, 900);
  {
    const uint4* src = reinterpret_cast<const uint4*>(wpack + (size_t)rank * WPACK_PER_RANK);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (int i = tid; i < (int)(WPACK_PER_RANK / 16); i += TC_THREADS) dst[i] = __ldg(src + i);
    for (int i = tid; i < 128 * 4; i += TC_THREADS) sW1[i] = ((i & 3) < C) ? W1[(i >> 2) * C + (i & 3)] : 0.f;
    if (tid < 128) sB1[tid] = b1[tid];
  }
  if (tid == 0) {
    mbar_init(bar(BAR_H1_FULL), 2);
    mbar_init(bar(BAR_ACC2_FULL), 1);
    mbar_init(bar(BAR_H2_FULL), 2);
    mbar_init(bar(BAR_ACC3_FULL0), 1);
    mbar_init(bar(BAR_ACC3_FULL1), 1);
    mbar_init(bar(BAR_ACC3_EMPTY0), 2);
    mbar_init(bar(BAR_ACC3_EMPTY1), 2);
    mbar_init(bar(BAR_L3_DONE), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {   // one warp per CTA allocates all 512 TMEM columns for the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();          // weight images were written with generic-proxy stores; the MMA reads via the async proxy
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  bool ok = true;

  if (warp < 4) {
    // =========================================================== group A: layer 1 + layer-2 epilogue (thread = point row)
    const int row = tid;                                      // 0..127 == TMEM lane
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint32_t it = 0;
    for (int b = cluster_id; b < B && ok; b += n_clusters) {
      for (int j = 0; j < tiles_per_cloud && ok; ++j, ++it) {
        const float* xp = x + (int64_t)b * ldx + (int64_t)(j * PTS_PER_TILE + rank * PTS_PER_CTA + row) * C;
        if (tid == 0) TSTAMP(0);
        float xv[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = 0; c < C; ++c) xv[c] = __ldg(xp + c);
        // ---- layer 1 into registers (overlaps the tail of the previous tile's L3 MMAs)
        uint32_t h1[64];
#pragma unroll
        for (int q = 0; q < 64; ++q) {
          const float4 w0 = *reinterpret_cast<const float4*>(sW1 + (2 * q) * 4);
          const float4 w1 = *reinterpret_cast<const float4*>(sW1 + (2 * q + 1) * 4);
          const float a0 = fmaf(xv[3], w0.w, fmaf(xv[2], w0.z, fmaf(xv[1], w0.y, fmaf(xv[0], w0.x, sB1[2 * q]))));
          const float a1 = fmaf(xv[3], w1.w, fmaf(xv[2], w1.z, fmaf(xv[1], w1.y, fmaf(xv[0], w1.x, sB1[2 * q + 1]))));
          h1[q] = pack_bf16(act_fast<ACT>(a0), act_fast<ACT>(a1));
        }
        // the H region is still being read by the previous tile's layer-3 MMAs
        if (tid == 0) TSTAMP(1);
        if (it > 0) ok = mbar_wait(bar(BAR_L3_DONE), (it - 1) & 1, err, 101);
        if (!ok) break;
        if (tid == 0) TSTAMP(2);
#pragma unroll
        for (int c16 = 0; c16 < 16; ++c16) {                  // 16 chunks of 8 channels; k-block = c16 / 8
          const uint32_t off = SM_H + (c16 >> 3) * KBLOCK_BYTES + sw128(row, c16 & 7);
          *reinterpret_cast<uint4*>(smem + off) = make_uint4(h1[4 * c16], h1[4 * c16 + 1], h1[4 * c16 + 2], h1[4 * c16 + 3]);
        }
        fence_proxy_async();
        named_bar_sync(1, 128);
        if (tid == 0) mbar_arrive_cluster(bar(BAR_H1_FULL), 0);
        if (tid == 0) TSTAMP(3);
        // ---- layer-2 epilogue: acc2 row -> +b2 -> act -> bf16 -> H2 row
        ok = mbar_wait(bar(BAR_ACC2_FULL), it & 1, err, 102);
        if (!ok) break;
        tc_fence_after();
        if (tid == 0) TSTAMP(4);
#pragma unroll 1
        for (int cc = 0; cc < 8; ++cc) {                      // 8 x 32 channels
          uint32_t v[32];
          tmem_ld32(lane_taddr + cc * 32, v);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float a0 = __uint_as_float(v[2 * i]) + __ldg(b2 + cc * 32 + 2 * i);
            const float a1 = __uint_as_float(v[2 * i + 1]) + __ldg(b2 + cc * 32 + 2 * i + 1);
            pk[i] = pack_bf16(act_fast<ACT>(a0), act_fast<ACT>(a1));
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t off = SM_H + (cc >> 1) * KBLOCK_BYTES + sw128(row, (cc & 1) * 4 + q);
            *reinterpret_cast<uint4*>(smem + off) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          }
        }
        tc_fence_before();
        fence_proxy_async();
        named_bar_sync(1, 128);
        if (tid == 0) mbar_arrive_cluster(bar(BAR_H2_FULL), 0);
        if (tid == 0) TSTAMP(5);
      }
    }
  } else if (warp < 8) {
    // =========================================================== group B: layer-3 epilogue (thread = output channel)
    const int lrow = tid - 128;                               // 0..127 == TMEM lane
    const uint32_t lane_taddr = tmem_base + ((uint32_t)((warp - 4) * 32) << 16);
    uint32_t it = 0;
    for (int b = cluster_id; b < B && ok; b += n_clusters) {
      float best[2] = {-INFINITY, -INFINITY};
      int besti[2] = {0, 0};
      for (int j = 0; j < tiles_per_cloud && ok; ++j, ++it) {
#pragma unroll 1
        for (int s = 0; s < 4 && ok; ++s) {
          const int buf = s & 1, chunk = s >> 1, half = s & 1;
          if (tid == 128) TSTAMP(8 + 3 * s);
          ok = mbar_wait(bar(BAR_ACC3_FULL0 + buf), (it * 2 + chunk) & 1, err, 103);
          if (!ok) break;
          tc_fence_after();
          if (tid == 128) TSTAMP(9 + 3 * s);
          float bv = best[chunk];
          int bi = besti[chunk];
#pragma unroll 1
          for (int cc = 0; cc < 4; ++cc) {                    // 4 x 32 columns (points)
            uint32_t v[32];
            tmem_ld32(lane_taddr + 256 + buf * 128 + cc * 32, v);
            tmem_ld_wait();
            if (WANT_ARGMAX) {
              // column n -> point: columns [0,64) come from CTA 0's rows, [64,128) from CTA 1's (B operand N halves)
              const int pbase = j * PTS_PER_TILE + (cc >> 1) * PTS_PER_CTA + half * 64 + (cc & 1) * 32;
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float f = __uint_as_float(v[i]);
                if (f > bv) { bv = f; bi = pbase + i; }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) bv = fmaxf(bv, __uint_as_float(v[i]));
            }
          }
          best[chunk] = bv;
          besti[chunk] = bi;
          tc_fence_before();
          named_bar_sync(2, 128);
          if (tid == 128) mbar_arrive_cluster(bar(BAR_ACC3_EMPTY0 + buf), 0);
          if (tid == 128) TSTAMP(10 + 3 * s);
        }
      }
      if (ok) {
#pragma unroll
        for (int chunk = 0; chunk < 2; ++chunk) {
          const int ch = rank * 256 + chunk * 128 + lrow;
          feat[(int64_t)b * ldf + ch] = best[chunk] + __ldg(b3 + ch);
          if (WANT_ARGMAX) argmax[(int64_t)b * 512 + ch] = besti[chunk];
        }
      }
    }
  } else if (rank == 0) {
    // =========================================================== warp 8 of the leader CTA: MMA issue
    const uint32_t idesc_l2 = umma_idesc(256, 256), idesc_l3 = umma_idesc(256, 128);
    uint32_t it = 0;
    for (int b = cluster_id; b < B && ok; b += n_clusters) {
      for (int j = 0; j < tiles_per_cloud && ok; ++j, ++it) {
        if (lane == 0) TSTAMP(24);
        ok = mbar_wait(bar(BAR_H1_FULL), it & 1, err, 104);
        if (!ok) break;
        tc_fence_after();
        if (lane == 0) TSTAMP(25);
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {                       // K = 128 = 2 k-blocks x 4 x UMMA_K(16)
            const uint32_t koff = (k >> 2) * KBLOCK_BYTES + (k & 3) * 32;
            umma_bf16_2cta(tmem_base, umma_desc(sbase + SM_H + koff), umma_desc(sbase + SM_W2 + koff), idesc_l2, k > 0);
          }
          umma_commit_mc(bar(BAR_ACC2_FULL));
        }
        __syncwarp();
        if (lane == 0) TSTAMP(26);
        ok = mbar_wait(bar(BAR_H2_FULL), it & 1, err, 105);
        if (!ok) break;
        tc_fence_after();
        if (lane == 0) TSTAMP(27);
#pragma unroll 1
        for (int s = 0; s < 4 && ok; ++s) {
          const int buf = s & 1, chunk = s >> 1, half = s & 1;
          ok = mbar_wait(bar(BAR_ACC3_EMPTY0 + buf), ((it * 2 + chunk) & 1) ^ 1, err, 106);
          if (!ok) break;
          tc_fence_after();
          if (lane == 0) TSTAMP(28 + 2 * s);
          if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {                    // K = 256 = 4 k-blocks x 4 x UMMA_K
              const uint32_t koff = (k >> 2) * KBLOCK_BYTES + (k & 3) * 32;
              umma_bf16_2cta(tmem_base + 256 + buf * 128, umma_desc(sbase + SM_W3 + chunk * 65536 + koff),
                             umma_desc(sbase + SM_H + half * (64 * 128) + koff), idesc_l3, k > 0);
            }
            umma_commit_mc(bar(BAR_ACC3_FULL0 + buf));
            if (s == 3) umma_commit_mc(bar(BAR_L3_DONE));
          }
          __syncwarp();
          if (lane == 0) TSTAMP(29 + 2 * s);
        }
      }
    }
  }

  // ---------------- teardown: everyone done with TMEM in both CTAs, then free it
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

extern "C" {

int pm_has_tcgen05(void) { return 1; }

// workspace: two packed weight images (rank 0 / rank 1) + an error word
size_t pm_pointnet_encode_forward_tc_ws_bytes(int, int, int) { return 2 * WPACK_PER_RANK + 4096; }

int pm_pointnet_encode_forward_tc(const float* x, int64_t ldx, int B, int N, int C, const pm_encoder_params* p, int act,
                                  float* feat, int64_t ldf, int32_t* argmax, void* ws, size_t ws_bytes, pm_stream_t s) {
  PM_REQUIRE(C >= 1 && C <= 4, PM_ERR_UNSUPPORTED, "bf16 encoder: C=%d channels per point (supports 1..4; use PM_PREC_FP32)", C);
  PM_REQUIRE(N % PTS_PER_TILE == 0, PM_ERR_UNSUPPORTED, "bf16 encoder: N=%d must be a multiple of %d (use PM_PREC_FP32)", N, PTS_PER_TILE);
  PM_REQUIRE(ws && ws_bytes >= pm_pointnet_encode_forward_tc_ws_bytes(B, N, C), PM_ERR_ARG, "bf16 encoder: workspace too small");
  PM_REQUIRE(pm_aligned(ws, 256), PM_ERR_ALIGN, "bf16 encoder: workspace must be 256-byte aligned");
  cudaStream_t st = pm_st(s);
  uint8_t* wpack = reinterpret_cast<uint8_t*>(ws);
  int32_t* err = reinterpret_cast<int32_t*>(wpack + 2 * WPACK_PER_RANK);
  cudaMemsetAsync(err, 0, sizeof(int32_t), st);
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
      encoder_fwd_tc<ACTV, true><<<grid, TC_THREADS, SM_TOTAL, st>>>(x, ldx, B, N, C, wpack, p->W1, p->b1, p->b2, p->b3, feat, ldf, argmax, err); \
    else                                                                                                                    \
      encoder_fwd_tc<ACTV, false><<<grid, TC_THREADS, SM_TOTAL, st>>>(x, ldx, B, N, C, wpack, p->W1, p->b1, p->b2, p->b3, feat, ldf, nullptr, err); \
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
