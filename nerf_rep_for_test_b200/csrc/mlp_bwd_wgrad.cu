// a7, MLP backward part 2: weight and bias gradients of one NeRF model as nine split-K GEMMs on tcgen05.
//   dW_s[out, in] = sum over rows  dpre_s[row, out] * x_s[row, in]
// Both factors are the tile images the forward (x_s: `acts`) and the dgrad kernel (dpre_s: `dacts`) left in
// HBM (train_layout.cuh).  The contraction runs over the ROWS of those images, so the very same bytes are
// consumed MN-major here (tc_ptx.cuh: umma_desc_sw128_mn) -- no transpose anywhere.  The kernel is bound by
// streaming 10.7 KB per row from HBM, the tensor pipe is about one third busy.
//
// One CTA pair (cta_group::2, M = 256 = the output index) per TPC, persistent over the 128-row tiles
// t = cluster, cluster + n_clusters, ...; stage-major: for each of the nine stages the pair accumulates
// its share of the rows in TMEM, then reduces the accumulator into a global fp32 scratch with
// red.global.add (74 pairs x 9 stages, a few MB of atomics per model).  Per CTA (rank r):
//   A  = dpre_s blocks 2r, 2r+1      (its 128 output rows of D)
//   B  = x_s    blocks 2r, 2r+1      (its half of D's 256 columns)
// Small products ride along on extra accumulator columns instead of extra kernels:
//   * bias gradients: B = a constant tile of ones, N = 16                   (every stage)
//   * the PE / dir-PE input columns of pts_linears.0/.5 and views_linears.0: B = PE / DPE block
//   * the tail stage (8): A = D9 = [d_hv | g_raw | 0], so with B = H7 rows 0..127 are dW' (fused views x feature
//     weight, train_layout.cuh) and row 131 is dW_alpha; with B = HV rows 128..130 are dW_rgb
// The gradients of views_linears.0[:, :256], feature_linear and their biases follow from dW', db' by the chain
// rule in wgrad_tail_kernel (two 128x256x256 products on CUDA cores, once per step).
// Narrow B operands are fed with N = 2 x (real width): both CTAs supply the same columns and the upper
// half of the product is a duplicate that is never read (a CTA cannot start mid-row in a swizzled image).
//
// Warps: 0 = producer (bulk copies into a 2-slot ring), 1 = MMA issuer (leader) / relay (peer) + TMEM
// allocator, 2..5 = accumulator flush.
#include "mlp_tc_common.cuh"
#include "train_layout.cuh"

namespace nb {
namespace wg {
using namespace ptx;

constexpr int kThreads = 192;
constexpr int kRing = 2;
constexpr uint32_t kRegionA = 0, kRegionB = 32768, kRegionX1 = 65536, kRegionX2 = 81920;
constexpr uint32_t kSlotBytes = 98304;
constexpr uint32_t kOffOnes = kRing * kSlotBytes;          // 196608: [16 K-rows][128 B] of bf16 1.0
constexpr uint32_t kOffBar = kOffOnes + 2048;
constexpr uint32_t kSmemBytes = kOffBar + 256;

enum { BAR_FULL = 0, BAR_EMPTY = 2, BAR_ACCFULL = 4, BAR_ACCFREE = 5, BAR_COUNT = 6 };

// TMEM accumulator columns
constexpr uint32_t kColMain = 0, kColAux = 256, kColAux2 = 320, kColOnes = 448;

struct StagePlan {
  int a_blk;      // dacts block of this CTA's A (2 blocks)
  int b_blk;      // acts block of this CTA's main B (2 blocks), -1: no main product
  int x1_blk;     // extra block 1 (PE / DPE from acts), -1: none
  int x2_blk;     // extra block 2 (HV block `rank`), -1: none
};
__device__ __forceinline__ StagePlan plan(int s, int rank) {
  StagePlan p;
  p.a_blk = (s < 8 ? dact_pre(s) : kDactD9) + 2 * rank;
  p.b_blk = s == 0 ? -1 : ((s < 8 ? act_h(s - 1) : act_h(7)) + 2 * rank);
  p.x1_blk = (s == 0 || s == 5) ? kActPe : (s == 8 ? kActDpe : -1);
  p.x2_blk = s == 8 ? kActHv + rank : -1;
  return p;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// accumulator columns [tcol, tcol + 16*n16) of this thread's row -> += into dst[0 .. 16*n16)
__device__ __forceinline__ void flush_cols(uint32_t t_row, uint32_t tcol, int n16, float* dst) {
  for (int i = 0; i < n16; ++i) {
    uint32_t v[16];
    tmem_ld16(t_row + tcol + (uint32_t)i * 16u, v);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q)
      red_add_v4(dst + i * 16 + q * 4, __uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]), __uint_as_float(v[q * 4 + 2]),
                 __uint_as_float(v[q * 4 + 3]));
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
mlp_bwd_wgrad_kernel(const unsigned char* __restrict__ acts, const unsigned char* __restrict__ dacts, int n_tiles,
                     float* __restrict__ scratch) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  const uint32_t smem_base = smem_u32(smem_dyn);
  if ((smem_base & 1023u) != 0) __trap();
  const uint32_t bar_base = smem_base + kOffBar;
  const uint32_t tmem_slot = bar_base + BAR_COUNT * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  auto bar = [&](int i) { return bar_base + (uint32_t)i * 8u; };

  if (threadIdx.x == 0) {
    for (int i = 0; i < kRing; ++i) {
      mbar_init(bar(BAR_FULL + i), rank == 0 ? 2 : 1);   // leader: own producer + peer relay
      mbar_init(bar(BAR_EMPTY + i), 1);
    }
    mbar_init(bar(BAR_ACCFULL), 1);
    mbar_init(bar(BAR_ACCFREE), 256);                    // both CTAs' flush warps (leader's copy is used)
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 512; i += kThreads) reinterpret_cast<uint32_t*>(smem_dyn + kOffOnes)[i] = 0x3F803F80u;
  fence_proxy_async_smem();
  if (warp == 1) tmem_alloc_2cta(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int my_tiles = n_tiles > cluster_id ? (n_tiles - cluster_id + num_clusters - 1) / num_clusters : 0;
  const size_t act_tile_bytes = (size_t)kActBlocks * kBlockBytes, dact_tile_bytes = (size_t)kDactBlocks * kBlockBytes;

  if (my_tiles > 0) {
    if (warp == 0) {
      // =========================== producer ===========================
      if (lane == 0) {
        uint32_t seq = 0;
        for (int s = 0; s < kWgradStages; ++s) {
          const StagePlan p = plan(s, (int)rank);
          const uint32_t bytes = 32768u + (p.b_blk >= 0 ? 32768u : 0u) + (p.x1_blk >= 0 ? 16384u : 0u) + (p.x2_blk >= 0 ? 16384u : 0u);
          for (int i = 0; i < my_tiles; ++i, ++seq) {
            const size_t t = (size_t)cluster_id + (size_t)i * num_clusters;
            const uint32_t pos = seq % kRing, phase = (seq / kRing) & 1u;
            const uint32_t slot = smem_base + pos * kSlotBytes, full = bar(BAR_FULL + pos);
            mbar_wait(bar(BAR_EMPTY + pos), phase ^ 1u, 0x200 + s);
            mbar_arrive_expect_tx(full, bytes);
            bulk_g2s(slot + kRegionA, dacts + t * dact_tile_bytes + (size_t)p.a_blk * kBlockBytes, 32768u, full);
            if (p.b_blk >= 0) bulk_g2s(slot + kRegionB, acts + t * act_tile_bytes + (size_t)p.b_blk * kBlockBytes, 32768u, full);
            if (p.x1_blk >= 0) bulk_g2s(slot + kRegionX1, acts + t * act_tile_bytes + (size_t)p.x1_blk * kBlockBytes, 16384u, full);
            if (p.x2_blk >= 0) bulk_g2s(slot + kRegionX2, acts + t * act_tile_bytes + (size_t)p.x2_blk * kBlockBytes, 16384u, full);
          }
        }
      }
      __syncwarp();
    } else if (warp == 1 && rank == 1) {
      // =========================== relay (peer CTA): my part of ring slot `pos` has landed ===========================
      if (lane == 0) {
        uint32_t seq = 0;
        for (int s = 0; s < kWgradStages; ++s)
          for (int i = 0; i < my_tiles; ++i, ++seq) {
            const uint32_t pos = seq % kRing, phase = (seq / kRing) & 1u;
            mbar_wait(bar(BAR_FULL + pos), phase, 0x700 + s);
            mbar_arrive_remote(mapa(bar(BAR_FULL + pos), 0));
          }
      }
      __syncwarp();
    } else if (warp == 1) {
      // =========================== MMA issuer (leader CTA) ===========================
      uint32_t seq = 0;
      const uint32_t ones_addr = smem_base + kOffOnes;
      for (int s = 0; s < kWgradStages; ++s) {
        // the flush warps of both CTAs have drained the previous stage's accumulators
        mbar_wait_cluster(bar(BAR_ACCFREE), (uint32_t)(s & 1) ^ 1u, 0x400 + s);
        tc_fence_after();
        for (int i = 0; i < my_tiles; ++i, ++seq) {
          const uint32_t pos = seq % kRing, phase = (seq / kRing) & 1u;
          const uint32_t slot = smem_base + pos * kSlotBytes;
          mbar_wait_cluster(bar(BAR_FULL + pos), phase, 0x300 + s);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll 1
            for (int k = 0; k < 8; ++k) {   // K = 16 rows per MMA
              const uint32_t koff = (uint32_t)k * 2048u;
              const uint32_t acc = (i > 0 || k > 0) ? 1u : 0u;
              const uint64_t da = umma_desc_sw128_mn(slot + kRegionA + koff, 16384u);
              if (s > 0)
                umma_bf16_ss_2cta(tmem_base + kColMain, da, umma_desc_sw128_mn(slot + kRegionB + koff, 16384u),
                                  umma_idesc_bf16_mn(256, 256), acc);
              if (s == 0)         // pts_linears.0: input = PE (64, fed as 2 x 64)
                umma_bf16_ss_2cta(tmem_base + kColMain, da, umma_desc_sw128_mn(slot + kRegionX1 + koff, 16384u),
                                  umma_idesc_bf16_mn(256, 128), acc);
              else if (s == 5)    // pts_linears.5: PE columns of the skip concat
                umma_bf16_ss_2cta(tmem_base + kColAux, da, umma_desc_sw128_mn(slot + kRegionX1 + koff, 16384u),
                                  umma_idesc_bf16_mn(256, 128), acc);
              else if (s == 8) {  // tail: views_linears.0 dir-PE columns (32, fed as 2 x 32) and rgb_linear (B = HV, 128)
                umma_bf16_ss_2cta(tmem_base + kColAux, da, umma_desc_sw128_mn(slot + kRegionX1 + koff, 16384u),
                                  umma_idesc_bf16_mn(256, 64), acc);
                umma_bf16_ss_2cta(tmem_base + kColAux2, da, umma_desc_sw128_mn(slot + kRegionX2 + koff, 16384u),
                                  umma_idesc_bf16_mn(256, 128), acc);
              }
              umma_bf16_ss_2cta(tmem_base + kColOnes, da, umma_desc_sw128_mn(ones_addr, 16384u), umma_idesc_bf16_mn(256, 16), acc);
            }
            umma_commit_2cta(bar(BAR_EMPTY + pos), 3);                        // ring slot free in both CTAs
            if (i == my_tiles - 1) umma_commit_2cta(bar(BAR_ACCFULL), 3);     // accumulators complete
          }
          __syncwarp();
        }
      }
    } else {
      // =========================== flush: TMEM -> red.global.add into the fp32 scratch ===========================
      const int q = warp & 3;                                  // TMEM lane quadrant this warp may read
      const int out = (int)rank * 128 + q * 32 + lane;         // output index (row of D)
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
      const uint32_t acc_free_leader = mapa(bar(BAR_ACCFREE), 0);
      for (int s = 0; s < kWgradStages; ++s) {
        mbar_wait(bar(BAR_ACCFULL), (uint32_t)(s & 1), 0x100 + s);
        tc_fence_after();
        float* dst = scratch + ((size_t)s * 256 + (size_t)out) * kGradCols;
        flush_cols(t_row, kColMain, s == 0 ? 4 : 16, dst + kGradMain);          // s = 0: PE columns only (64)
        if (s == 5) flush_cols(t_row, kColAux, 4, dst + kGradAux);
        if (s == 8) {
          flush_cols(t_row, kColAux, 2, dst + kGradAux);
          if (out >= 128) flush_cols(t_row, kColAux2, 8, dst + kGradAux2);       // rows 128..130 hold dW_rgb
        }
        flush_cols(t_row, kColOnes, 1, dst + kGradOnes);
        tc_fence_before();
        mbar_arrive_remote(acc_free_leader);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

// scratch -> the gradient tensors that are plain copies (everything except views_w[:, :256], feature_w/b)
__global__ void wgrad_finalize_kernel(const float* __restrict__ scratch, nerfb200_mlp_grads g) {
  const int s = blockIdx.y;
  const float* sc = scratch + (size_t)s * 256 * kGradCols;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 256 * 320; i += gridDim.x * blockDim.x) {
    const int out = i / 320, c = i % 320;
    const float* row = sc + (size_t)out * kGradCols;
    if (s < 8) {
      if (c == 319) { g.pts_b[s][out] = row[kGradOnes]; continue; }
      if (s == 0) { if (c < kChX) g.pts_w[0][out * kChX + c] = row[kGradMain + c]; }
      else if (s == 5) {
        if (c < kChX) g.pts_w[5][out * 319 + c] = row[kGradAux + c];
        else if (c < 319) g.pts_w[5][out * 319 + c] = row[kGradMain + c - kChX];
      } else if (c < 256) g.pts_w[s][out * 256 + c] = row[kGradMain + c];
    } else {
      if (out < 128) {
        if (c >= 256 && c < 283) g.views_w[out * 283 + c] = row[kGradAux + c - 256];
        else if (c == 283) g.views_b[out] = row[kGradOnes];               // b' = Wv_a bf + bv  ->  d bv = d b'
      } else if (out < 131) {
        if (c < 128) g.rgb_w[(out - 128) * 128 + c] = row[kGradAux2 + c];
        else if (c == 128) g.rgb_b[out - 128] = row[kGradOnes];
      } else if (out == 131) {
        if (c < 256) g.alpha_w[c] = row[kGradMain + c];
        else if (c == 256) g.alpha_b[0] = row[kGradOnes];
      }
    }
  }
}

// chain rule through the fused tail W' = Wv_a Wf, b' = Wv_a bf + bv (Wv_a = views_linears.0.weight[:, :256]):
//   d Wv_a[n][j] = sum_k dW'[n][k] Wf[j][k] + db'[n] bf[j]      (blocks 0..127, one output row n each)
//   d Wf[j][k]   = sum_n Wv_a[n][j] dW'[n][k],  d bf[j] = sum_n Wv_a[n][j] db'[n]   (blocks 128..383, row j each)
__global__ void __launch_bounds__(1024) wgrad_tail_kernel(const float* __restrict__ scratch, nerfb200_mlp_weights w, nerfb200_mlp_grads g) {
  const float* sc = scratch + (size_t)8 * 256 * kGradCols;     // rows 0..127: dW' in cols 0..255, db' in col kGradOnes
  __shared__ float sh[256];
  __shared__ float part[4][256];
  const int t = threadIdx.x & 255, q = threadIdx.x >> 8, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (blockIdx.x < 128) {
    // d views_w[n][0:256] = dW'[n][:] Wf^T + db'[n] bf: warp = 8 output columns, the lanes split the contraction
    // (coalesced rows of Wf, all loads independent), shuffle tree at the end
    const int n = blockIdx.x;
    if (q == 0) sh[t] = sc[(size_t)n * kGradCols + kGradMain + t];
    __syncthreads();
    const float dbn = sc[(size_t)n * kGradCols + kGradOnes];
    float x[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) x[r] = sh[lane * 8 + r];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int col = warp * 8 + i;
      const float4* wf = reinterpret_cast<const float4*>(w.feature_w + (size_t)col * 256 + lane * 8);
      const float4 v0 = __ldg(wf), v1 = __ldg(wf + 1);
      float acc = x[0] * v0.x;
      acc = fmaf(x[1], v0.y, acc); acc = fmaf(x[2], v0.z, acc); acc = fmaf(x[3], v0.w, acc);
      acc = fmaf(x[4], v1.x, acc); acc = fmaf(x[5], v1.y, acc); acc = fmaf(x[6], v1.z, acc); acc = fmaf(x[7], v1.w, acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) g.views_w[(size_t)n * 283 + col] = fmaf(dbn, w.feature_b[col], acc);
    }
  } else {
    // d feature_w[j][0:256] = Wv[:, j]^T dW', d feature_b[j] = Wv[:, j]^T db': four thread groups x 32 rows each
    const int j = blockIdx.x - 128;
    if (threadIdx.x < 128) sh[threadIdx.x] = w.views_w[(size_t)threadIdx.x * 283 + j];      // column j of Wv_a
    __syncthreads();
    float acc = 0.f, accb = 0.f;
#pragma unroll 16
    for (int n = q * 32; n < q * 32 + 32; ++n) {
      acc = fmaf(sh[n], sc[(size_t)n * kGradCols + kGradMain + t], acc);
      if (t == 0) accb = fmaf(sh[n], sc[(size_t)n * kGradCols + kGradOnes], accb);
    }
    part[q][t] = acc;
    if (t == 0) sh[128 + q] = accb;
    __syncthreads();
    if (q == 0) {
      g.feature_w[(size_t)j * 256 + t] = (part[0][t] + part[1][t]) + (part[2][t] + part[3][t]);
      if (t == 0) g.feature_b[j] = (sh[128] + sh[129]) + (sh[130] + sh[131]);
    }
  }
}

}  // namespace wg

int launch_mlp_bwd_wgrad(const void* acts, const void* dacts, long long M, float* scratch, const nerfb200_mlp_weights* weights,
                         const nerfb200_mlp_grads* grads, cudaStream_t st) {
  using namespace wg;
  int dev = 0, sms = 0;
  NB_CUDA(cudaGetDevice(&dev));
  NB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static bool attr_set[64] = {};   // per-device, sticky: set once (also keeps the call out of CUDA-graph captures)
  NB_CHECK_ARG(dev >= 0 && dev < 64, "mlp_backward: device ordinal %d out of range", dev);
  if (!attr_set[dev]) {
    NB_CUDA(cudaFuncSetAttribute(mlp_bwd_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    attr_set[dev] = true;
  }
  NB_CUDA(cudaMemsetAsync(scratch, 0, kGradScratchFloats * sizeof(float), st));
  const int n_tiles = (int)((M + 127) / 128);
  const int clusters = n_tiles < sms / 2 ? n_tiles : sms / 2;
  if (n_tiles > 0) {
    mlp_bwd_wgrad_kernel<<<2 * clusters, kThreads, kSmemBytes, st>>>((const unsigned char*)acts, (const unsigned char*)dacts, n_tiles, scratch);
    NB_LAUNCH_OK("mlp_bwd_wgrad_kernel");
  }
  wgrad_finalize_kernel<<<dim3(40, kWgradStages), 256, 0, st>>>(scratch, *grads);
  NB_LAUNCH_OK("wgrad_finalize_kernel");
  wgrad_tail_kernel<<<384, 1024, 0, st>>>(scratch, *weights, *grads);
  NB_LAUNCH_OK("wgrad_tail_kernel");
  return 0;
}

}  // namespace nb
