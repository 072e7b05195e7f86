// a7, MLP backward part 1: the activation-gradient (dgrad) chain of the 8x256 NeRF MLP on tcgen05.
// The reference has no hand-written backward (autograd through network.py:49-74); this kernel walks the
// same graph in reverse for a tile of 128 rows per CTA (256 per CTA pair, cta_group::2), exactly like the
// forward kernel (mlp_bf16_tc2.cu) walks it forward:
//
//   prologue (CUDA cores)   d_hv   = relu'(hv)  . (g_rgb  * rgb_linear.W)                     [128]
//   bstage 0  (K = 128)     d_pre7 = relu'(h7) . (d_hv . W'[:, :256] + g_sigma * alpha_linear.W),  W' = Wv[:, :256] Wf
//                           (the fused tail of train_layout.cuh: feature_linear is linear, so it never appears alone)
//   bstage 1..7             d_pre(i-1) = relu'(h(i-1)) . (d_pre(i) . pts_linears.i.W[:, hidden cols]),  i = 7..1
//
// Every stage's output is (a) the next stage's A operand in shared memory and (b) the A operand of the
// weight-gradient GEMM, so the finished 64 KB tile image is also bulk-stored to the `dacts` store
// (train_layout.cuh).  relu' comes from the sign-bit planes the training forward wrote (32 B per row and
// stage instead of re-reading 512 B of activations).  Inputs are not differentiated (the hierarchical
// sampler is detached, training.py), so there is no stage for pts_linears.0 / the PE columns.
//
// Warp roles, barriers, the weight ring and the slot schedule are those of mlp_bf16_tc2.cu.
#include "mlp_tc_common.cuh"
#include "train_layout.cuh"

namespace nb {
namespace bwd {
using namespace ptx;

constexpr int kThreads = 320;
constexpr int kRing = 4;
constexpr uint32_t kABytes = 65536, kWStageBytes = 16384;
constexpr uint32_t kOffA = 0;
constexpr uint32_t kOffW = kOffA + 2 * kABytes;              // 131072
constexpr uint32_t kOffBar = kOffW + kRing * kWStageBytes;   // 196608
constexpr uint32_t kOffTail = kOffBar + 256;                 // fp32 rgb_w [3][128] | alpha_w [256]
constexpr uint32_t kSmemBytes = kOffTail + kBwdTailFloats * 4;

enum { BAR_WFULL = 0, BAR_WEMPTY = 4, BAR_AREADY = 8, BAR_ACCFULL = 10, BAR_COUNT = 12 };

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0u), "r"(sel));
  return d;
}

// 8 consecutive columns (8q .. 8q+7 of a 32-column group) -> four bf16x2 words, zeroed where the relu
// sign bit of the column is set.  Bit layout of `mb`: train_layout.cuh.
template <bool kMasked>
__device__ __forceinline__ uint4 pack8(const float (&x)[8], uint32_t mb, int q) {
  uint4 o;
  o.x = cvt_bf16x2<false>(x[0], x[1]); o.y = cvt_bf16x2<false>(x[2], x[3]);
  o.z = cvt_bf16x2<false>(x[4], x[5]); o.w = cvt_bf16x2<false>(x[6], x[7]);
  if (kMasked) {
    const uint32_t t0 = mb << (2 * q), t1 = mb << (2 * q + 1);
    o.x &= ~prmt(t0, 0x9988u); o.y &= ~prmt(t0, 0xBBAAu);
    o.z &= ~prmt(t1, 0x9988u); o.w &= ~prmt(t1, 0xBBAAu);
  }
  return o;
}

// MODE 0: linear (unused since the tail is fused); 1: + g_sigma * alpha_w, masked (d_pre7); 2: masked (d_pre6..0)
template <int MODE>
__device__ __forceinline__ void depi32(const uint32_t (&v)[32], unsigned char* out_row, int j0, int r7, uint32_t mb,
                                       float gs, const float* __restrict__ aw) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(v[q * 8 + i]);
    if (MODE == 1) {
      const float4 a0 = *reinterpret_cast<const float4*>(aw + q * 8), a1 = *reinterpret_cast<const float4*>(aw + q * 8 + 4);
      x[0] = fmaf(gs, a0.x, x[0]); x[1] = fmaf(gs, a0.y, x[1]); x[2] = fmaf(gs, a0.z, x[2]); x[3] = fmaf(gs, a0.w, x[3]);
      x[4] = fmaf(gs, a1.x, x[4]); x[5] = fmaf(gs, a1.y, x[5]); x[6] = fmaf(gs, a1.z, x[6]); x[7] = fmaf(gs, a1.w, x[7]);
    }
    *reinterpret_cast<uint4*>(out_row + (((j0 + q) ^ r7) << 4)) = pack8<MODE != 0>(x, mb, q);
  }
}

template <int MODE>
__device__ __forceinline__ void depi_stage256(uint32_t t_acc, unsigned char* a_row_base, int r7, const uint32_t (&mw)[8],
                                              float gs, const float* __restrict__ aw, unsigned char* bulk_g, uint32_t bulk_s) {
  uint32_t va[32], vb[32];
  tmem_ld32(t_acc, va);
  tmem_ld32(t_acc + 32u, vb);
  tmem_ld_wait();
  pin32(va);
  pin32(vb);
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    unsigned char* out_row = a_row_base + h * 16384;
    depi32<MODE>(va, out_row, 0, r7, mw[2 * h], gs, aw + h * 64);
    if (h < 3) tmem_ld32(t_acc + (uint32_t)(h * 64 + 64), va);
    depi32<MODE>(vb, out_row, 4, r7, mw[2 * h + 1], gs, aw + h * 64 + 32);
    // K-block h of this warp's rows -> activation-gradient store, while the epilogue goes on (mlp_tc_common.cuh)
    if (bulk_g != nullptr) bulk_store_warp_rows(bulk_g + h * 16384, bulk_s + (uint32_t)h * 16384u, 1);
    if (h < 3) {
      tmem_ld32(t_acc + (uint32_t)(h * 64 + 96), vb);
      tmem_ld_wait();
      pin32(va);
      pin32(vb);
    }
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
mlp_bwd_dgrad_kernel(const unsigned char* __restrict__ packed_bwd, const float* __restrict__ g_raw,
                     const uint32_t* __restrict__ masks, unsigned char* __restrict__ dacts, long long M, int num_quads) {
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
      mbar_init(bar(BAR_WFULL + i), rank == 0 ? 2 : 1);
      mbar_init(bar(BAR_WEMPTY + i), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(BAR_AREADY + s), 2);   // one arrival per epilogue group and CTA (its first thread, after the named barrier)
      mbar_init(bar(BAR_ACCFULL + s), 1);
    }
    fence_mbar_init();
  }
  {  // fp32 head weights -> shared memory (read by every epilogue thread, broadcast)
    const float* src = reinterpret_cast<const float*>(packed_bwd + kBwdTailOff);
    float* dst = reinterpret_cast<float*>(smem_dyn + kOffTail);
    for (int i = threadIdx.x; i < kBwdTailFloats; i += kThreads) dst[i] = src[i];
  }
  if (warp == 9) tmem_alloc_2cta(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int my_quads = num_quads > cluster_id ? (num_quads - cluster_id + num_clusters - 1) / num_clusters : 0;
  const long long n_tiles = (M + 127) / 128;

  if (warp < 8) {
    // =========================== epilogue groups ===========================
    const int slot = warp >> 2;
    const int w4 = warp & 3;
    const int row = w4 * 32 + lane;
    const int r7 = row & 7;
    unsigned char* a_row_base = smem_dyn + kOffA + (uint32_t)slot * kABytes + (uint32_t)row * 128u;
    const uint32_t t_acc = tmem_base + ((uint32_t)(w4 * 32) << 16) + (uint32_t)slot * 256u;
    const uint32_t b_ready_leader = mapa(bar(BAR_AREADY + slot), 0);
    const uint32_t b_full = bar(BAR_ACCFULL + slot);
    const bool leader = w4 == 0 && lane == 0;
    const float* rgb_w = reinterpret_cast<const float*>(smem_dyn + kOffTail) + kBwdTailRgbW;
    const float* alpha_w = reinterpret_cast<const float*>(smem_dyn + kOffTail) + kBwdTailAlphaW;
    const size_t mask_plane = (size_t)n_tiles * 128 * kMaskWords;
    uint32_t full_phase = 0;

    for (int it = 0; it < my_quads; ++it) {
      const long long tile = 4LL * ((long long)cluster_id + (long long)it * num_clusters) + 2 * slot + (long long)rank;
      const long long m = tile * 128 + row;
      const bool tile_ok = tile < n_tiles;
      unsigned char* dacts_tile = dacts + (size_t)(tile_ok ? tile : 0) * ((size_t)kDactBlocks * kBlockBytes);
      const size_t mask_row = ((size_t)(tile_ok ? tile : 0) * 128 + (size_t)row) * kMaskWords;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < M) g = __ldg(reinterpret_cast<const float4*>(g_raw) + m);
      uint4 hvm = make_uint4(0u, 0u, 0u, 0u);
      if (tile_ok) hvm = __ldg(reinterpret_cast<const uint4*>(masks + 8 * mask_plane + mask_row));
      bulk_store_warp_reads_done();    // the previous tile's last stage has left this warp's rows of the A tile
      {  // D9 tile: d_hv (blocks 0,1) | g_raw as bf16 + zero pad (block 2) | zeros (block 3)
        const uint32_t hv_words[4] = {hvm.x, hvm.y, hvm.z, hvm.w};
#pragma unroll
        for (int hb = 0; hb < 2; ++hb)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int n0 = hb * 64 + j * 8;
            float x[8];
#pragma unroll
            for (int i4 = 0; i4 < 2; ++i4) {
              const float4 w0 = *reinterpret_cast<const float4*>(rgb_w + n0 + i4 * 4);
              const float4 w1 = *reinterpret_cast<const float4*>(rgb_w + 128 + n0 + i4 * 4);
              const float4 w2 = *reinterpret_cast<const float4*>(rgb_w + 256 + n0 + i4 * 4);
              x[i4 * 4 + 0] = fmaf(g.z, w2.x, fmaf(g.y, w1.x, g.x * w0.x));
              x[i4 * 4 + 1] = fmaf(g.z, w2.y, fmaf(g.y, w1.y, g.x * w0.y));
              x[i4 * 4 + 2] = fmaf(g.z, w2.z, fmaf(g.y, w1.z, g.x * w0.z));
              x[i4 * 4 + 3] = fmaf(g.z, w2.w, fmaf(g.y, w1.w, g.x * w0.w));
            }
            *reinterpret_cast<uint4*>(a_row_base + hb * 16384 + ((j ^ r7) << 4)) = pack8<true>(x, hv_words[hb * 2 + (j >> 2)], j & 3);
          }
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        const uint4 gq = make_uint4(pack_bf16x2(g.x, g.y), pack_bf16x2(g.z, g.w), 0u, 0u);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          *reinterpret_cast<uint4*>(a_row_base + 2 * 16384 + ((j ^ r7) << 4)) = (j == 0) ? gq : zero;
          *reinterpret_cast<uint4*>(a_row_base + 3 * 16384 + ((j ^ r7) << 4)) = zero;
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + slot, 128);
      if (leader) mbar_arrive_remote(b_ready_leader);
      const uint32_t a_warp_s = smem_base + kOffA + (uint32_t)slot * kABytes + (uint32_t)w4 * 4096u;   // this warp's rows of block 0
      if (tile_ok) bulk_store_warp_rows(dacts_tile + (size_t)kDactD9 * kBlockBytes + (size_t)w4 * 4096u, a_warp_s, 4);

      for (int bs = 0; bs < kBwdStages; ++bs) {
        uint32_t mw[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        if (tile_ok) {   // relu sign bits of h(7-bs), fetched before the wait to hide the latency
          const uint4* mp = reinterpret_cast<const uint4*>(masks + (size_t)(7 - bs) * mask_plane + mask_row);
          const uint4 m0 = __ldg(mp), m1 = __ldg(mp + 1);
          mw[0] = m0.x; mw[1] = m0.y; mw[2] = m0.z; mw[3] = m0.w;
          mw[4] = m1.x; mw[5] = m1.y; mw[6] = m1.z; mw[7] = m1.w;
        }
        mbar_wait(b_full, full_phase, 0x100 + bs);
        full_phase ^= 1;
        tc_fence_after();
        bulk_store_warp_reads_done();    // the previous stage's stores have finished reading this warp's rows
        unsigned char* bulk_g = tile_ok ? dacts_tile + (size_t)dact_pre(7 - bs) * kBlockBytes + (size_t)w4 * 4096u : nullptr;
        if (bs == 0) depi_stage256<1>(t_acc, a_row_base, r7, mw, g.w, alpha_w, bulk_g, a_warp_s);
        else depi_stage256<2>(t_acc, a_row_base, r7, mw, 0.f, alpha_w, bulk_g, a_warp_s);
        tc_fence_before();
        fence_proxy_async_smem();
        named_bar_sync(1 + slot, 128);
        if (leader && bs + 1 < kBwdStages) mbar_arrive_remote(b_ready_leader);
      }
    }
    bulk_store_warp_drain();
  } else if (warp == 8) {
    // =========================== producer: this CTA's half of every W^T chunk ===========================
    if (lane == 0) {
      uint32_t seq = 0;
      const uint32_t half = kBwdChunkBytes >> 1;
      for (int it = 0; it < my_quads; ++it)
        for (int bs = 0; bs < kBwdStages; ++bs) {
          const unsigned char* src = packed_bwd + (size_t)bwd_stage_off(bs) + (size_t)rank * half;
          const int nch = bwd_chunks(bs);
          for (int c = 0; c < nch; ++c, ++seq) {
            const uint32_t pos = seq % kRing, phase = (seq / kRing) & 1u;
            mbar_wait(bar(BAR_WEMPTY + pos), phase ^ 1u, 0x200 + bs);
            mbar_arrive_expect_tx(bar(BAR_WFULL + pos), half);
            bulk_g2s(smem_base + kOffW + pos * kWStageBytes, src + (size_t)c * kBwdChunkBytes, half, bar(BAR_WFULL + pos));
          }
        }
    }
    __syncwarp();
  } else if (rank == 1) {
    // =========================== relay (peer CTA) ===========================
    if (lane == 0) {
      uint32_t seq = 0;
      for (int it = 0; it < my_quads; ++it)
        for (int bs = 0; bs < kBwdStages; ++bs) {
          const int nch = bwd_chunks(bs);
          for (int c = 0; c < nch; ++c, ++seq) {
            const uint32_t pos = seq % kRing, phase = (seq / kRing) & 1u;
            mbar_wait(bar(BAR_WFULL + pos), phase, 0x700 + bs);
            mbar_arrive_remote(mapa(bar(BAR_WFULL + pos), 0));
          }
        }
    }
    __syncwarp();
  } else {
    // =========================== MMA issuer (leader CTA) ===========================
    uint32_t seq = 0, ready_phase = 0;
    const uint32_t desc_hi = (uint32_t)(umma_desc_sw128(0) >> 32);
    const uint32_t lo_flags = (uint32_t)(umma_desc_sw128(0) & 0xFFFFFFFFu);
    const uint32_t a_lo0 = lo_flags | ((smem_base + kOffA) >> 4);
    const uint32_t w_lo0 = lo_flags | ((smem_base + kOffW) >> 4);
    const uint32_t idesc = umma_idesc_bf16(256, 256);
    for (int it = 0; it < my_quads; ++it) {
      for (int bs = 0; bs < kBwdStages; ++bs) {
        const int nch = bwd_chunks(bs);
#pragma unroll 1
        for (int slot = 0; slot < 2; ++slot) {
          mbar_wait_cluster(bar(BAR_AREADY + slot), ready_phase, 0x400 + bs * 2 + slot);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)slot * 256u;
#pragma unroll 1
          for (int c = 0; c < nch; ++c) {
            const uint32_t cs = seq + (uint32_t)c;
            const uint32_t pos = cs % kRing, phase = (cs / kRing) & 1u;
            if (slot == 0) {
              mbar_wait_cluster(bar(BAR_WFULL + pos), phase, 0x300 + bs);
              tc_fence_after();
            }
            const uint32_t a_lo = a_lo0 + (uint32_t)slot * (kABytes >> 4) + (uint32_t)c * 1024u;
            const uint32_t b_lo = w_lo0 + pos * (kWStageBytes >> 4);
            if (elect_one()) {
              const uint64_t hi64 = (uint64_t)desc_hi << 32;
              umma_bf16_ss_2cta(d_tmem, hi64 | (a_lo + 0u), hi64 | (b_lo + 0u), idesc, c > 0 ? 1u : 0u);
              umma_bf16_ss_2cta(d_tmem, hi64 | (a_lo + 2u), hi64 | (b_lo + 2u), idesc, 1u);
              umma_bf16_ss_2cta(d_tmem, hi64 | (a_lo + 4u), hi64 | (b_lo + 4u), idesc, 1u);
              umma_bf16_ss_2cta(d_tmem, hi64 | (a_lo + 6u), hi64 | (b_lo + 6u), idesc, 1u);
              if (slot == 1) umma_commit_2cta(bar(BAR_WEMPTY + pos), 3);
              if (c == nch - 1) umma_commit_2cta(bar(BAR_ACCFULL + slot), 3);
            }
            __syncwarp();
          }
        }
        seq += (uint32_t)nch;
        ready_phase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

}  // namespace bwd

int launch_mlp_bwd_dgrad(const void* packed_bwd, const float* g_raw, const void* masks, void* dacts, long long M,
                         cudaStream_t st) {
  using namespace bwd;
  int dev = 0, sms = 0;
  NB_CUDA(cudaGetDevice(&dev));
  NB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static bool attr_set[64] = {};   // per-device, sticky: set once (also keeps the call out of CUDA-graph captures)
  NB_CHECK_ARG(dev >= 0 && dev < 64, "mlp_backward: device ordinal %d out of range", dev);
  if (!attr_set[dev]) {
    NB_CUDA(cudaFuncSetAttribute(mlp_bwd_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    attr_set[dev] = true;
  }
  long long quads = (M + 511) / 512;
  int clusters = (int)(quads < sms / 2 ? quads : sms / 2);
  mlp_bwd_dgrad_kernel<<<2 * clusters, kThreads, kSmemBytes, st>>>((const unsigned char*)packed_bwd, g_raw,
                                                                   (const uint32_t*)masks, (unsigned char*)dacts, M, (int)quads);
  NB_LAUNCH_OK("mlp_bwd_dgrad_kernel");
  return 0;
}

}  // namespace nb
