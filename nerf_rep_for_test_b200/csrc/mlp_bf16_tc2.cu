// a3, bf16 performance mode, CTA-pair kernel: positional encoding fused into the 8x256 NeRF MLP on
// tcgen05 with cta_group::2 (two SMs of a TPC issue one 256-row MMA; each SM supplies its own 128
// A rows and HALF of the weight tile, so the shared-memory operand traffic per SM drops from 12 KB
// to 8 KB per K=16 step -- the 1-CTA kernel is bound by that traffic at 194 cycles per MMA, measured,
// against the 128-cycle tensor floor).
// Reference: volume_renderer.py:270-284 -> freq.py:23-26 -> network.py:49-74.
//
// One cluster of 2 CTAs per TPC, persistent over "quads" of 4 tiles x 128 rows.  Per CTA, 10 warps:
//   warps 0-3 / 4-7  epilogue groups of tile slot 0 / 1 (this CTA's 128 rows of the slot's 256)
//   warp 8           producer: own half (N/2 rows) of every weight K-chunk via cp.async.bulk into a
//                    4-stage ring, plus the stage's fp32 bias block into shared memory
//   warp 9           rank 0: MMA issuer (one thread, tcgen05.mma.cta_group::2) + TMEM allocator
//                    rank 1: relay ("my half of chunk c has landed" -> leader's mbarrier) + allocator
// Schedule per stage (mlp_layout.cuh): slot 0 runs ALL its K-chunks, then slot 1 runs the same chunks
// and releases them (the ring holds a whole stage: 4 x 16 KB), so every weight byte is used for
// 512 rows and each slot has a full slot-pass (2048 cycles) to drain its accumulator (TMEM read is
// the second bound: ~2000 cycles per 128x256 fp32 tile) while the tensor pipe works on the other slot.
#include <stdlib.h>

#include "mlp_tc_common.cuh"
#include "train_layout.cuh"

namespace nb {
namespace tc2 {
using namespace ptx;

constexpr int kThreads = 320;
constexpr int kRing = 4;
constexpr uint32_t kABytes = 65536, kPeBytes = 16384, kWStageBytes = 16384;
constexpr uint32_t kOffA = 0;
constexpr uint32_t kOffPe = kOffA + 2 * kABytes;            // 131072
constexpr uint32_t kOffW = kOffPe + 2 * kPeBytes;           // 163840
constexpr uint32_t kOffBar = kOffW + kRing * kWStageBytes;  // 229376
constexpr uint32_t kOffBias = kOffBar + 256;
constexpr uint32_t kSmemBytes = kOffBias + 2048;            // 231680 <= 232448

enum { BAR_WFULL = 0, BAR_WEMPTY = 4, BAR_AREADY = 8, BAR_ACCFULL = 10, BAR_BFULL = 12, BAR_BEMPTY = 14, BAR_COUNT = 16 };

// which on-chip buffer holds K-chunk c of a stage's input, and how many K=16 steps it has
// (`last` = the views stage: 9 in the unfused image, 8 = "8F" in the fused inference image)
__device__ __forceinline__ void chunk_src(int stage, int last, int c, bool& from_pe, int& kblock, int& ksteps) {
  ksteps = 4;
  if (stage == 0) { from_pe = true; kblock = 0; }
  else if (stage == 5) { from_pe = (c == 0); kblock = c - 1; }
  else if (stage == last) { from_pe = (c == 4); kblock = c; if (c == 4) ksteps = 2; }
  else { from_pe = false; kblock = c; }
}

// per-stage geometry of the packed image, unfused (10 stages) or fused (9 stages, mlp_layout.cuh)
template <bool kFused> struct Img {
  static constexpr int kN = kFused ? 9 : kStages;
  static constexpr int kLast = kN - 1;
  __device__ static __forceinline__ int chunks(int s) { return (kFused && s == 8) ? kFusedChunks : stage_chunks(s); }
  __device__ static __forceinline__ int n(int s) { return (kFused && s == 8) ? kFusedN : stage_n(s); }
  __device__ static __forceinline__ uint32_t chunk_bytes(int s) { return (kFused && s == 8) ? (uint32_t)kFusedChunkBytes : (uint32_t)bf16_chunk_bytes(s); }
  __device__ static __forceinline__ size_t stage_off(int s) { return (kFused && s == 8) ? (size_t)kFusedStageOff : (size_t)bf16_stage_off(s); }
  __device__ static __forceinline__ const float* bias(const unsigned char* packed, int s) {
    return (kFused && s == 8) ? reinterpret_cast<const float*>(packed + kFusedTailOff)
                              : reinterpret_cast<const float*>(packed + kBf16TailOff) + kTailBias + s * 256;
  }
};

template <bool kDump, bool kTimeline, bool kSave, bool kFused, bool kF16 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
mlp_bf16_tc2_kernel(const unsigned char* __restrict__ packed, const float* __restrict__ rays_o,
                    const float* __restrict__ rays_d, const float* __restrict__ z_vals, long long M, int S,
                    int num_quads, float* __restrict__ raw, float* __restrict__ stage_dump,
                    unsigned long long* __restrict__ tl, unsigned char* __restrict__ acts,
                    uint32_t* __restrict__ masks, const int* __restrict__ row_ids, const int* __restrict__ n_active,
                    int dbg = 0) {
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
      mbar_init(bar(BAR_WFULL + i), rank == 0 ? 2 : 1);   // leader: own producer + peer relay
      mbar_init(bar(BAR_WEMPTY + i), 1);
    }
    for (int s = 0; s < 2; ++s) {
      // both CTAs' epilogue groups signal the leader's copy: every thread in inference; in training only the
      // group's first thread, behind a 128-thread named barrier (the release fence of the remote arrive waits for
      // the warp's pending global stores -- the relu-mask rows here --, which is what made the first training forward
      // 2x slower than inference)
      mbar_init(bar(BAR_AREADY + s), kSave ? 2 : 256);
      mbar_init(bar(BAR_ACCFULL + s), 1);
      mbar_init(bar(BAR_BFULL + s), 1);
      mbar_init(bar(BAR_BEMPTY + s), 256);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc_2cta(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // peer barriers initialised before any remote arrive / multicast commit
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // sparse launch (empty-space skipping): only the rows listed in row_ids[0 .. *n_active) are evaluated;
  // the count lives on the device, so the persistent grid derives its trip count from it (no host sync)
  long long M_eff = M;
  if (row_ids != nullptr) {
    M_eff = *n_active;
    num_quads = (int)((M_eff + 511) / 512);
  }
  using I = Img<kFused>;
  const int my_quads = num_quads > cluster_id ? (num_quads - cluster_id + num_clusters - 1) / num_clusters : 0;
  const float* tail = reinterpret_cast<const float*>(packed + kBf16TailOff);

  if (warp < 8) {
    // =========================== epilogue groups ===========================
    const int slot = warp >> 2;
    const int w4 = warp & 3;
    const int row = w4 * 32 + lane;
    const int r7 = row & 7;
    unsigned char* a_row_base = smem_dyn + kOffA + (uint32_t)slot * kABytes + (uint32_t)row * 128u;
    const uint32_t pe_base = smem_base + kOffPe + (uint32_t)slot * kPeBytes;
    const uint32_t t_acc = tmem_base + ((uint32_t)(w4 * 32) << 16) + (uint32_t)slot * 256u;
    const uint32_t b_ready_leader = mapa(bar(BAR_AREADY + slot), 0);
    const uint32_t b_full = bar(BAR_ACCFULL + slot);
    uint32_t full_phase = 0;

    // rays / PE of a tile are computed one tile ahead so that only the stores sit on the critical path
    long long m = 0, tile_next = 0;
    bool valid = false;
    // training (kSave): finished operand tiles go to the activation store as tile images (train_layout.cuh), as per-warp
    // bulk stores issued while the epilogue produces them (mlp_tc_common.cuh, bulk_store_warp_rows)
    const long long n_tiles = (M + 127) / 128;
    const bool save_leader = kSave && w4 == 0 && lane == 0;
    float d[3] = {0.f, 0.f, 0.f};
    uint32_t pe_pk[32];
    auto prepare_tile = [&](int it) {
      const long long tile = 4LL * ((long long)cluster_id + (long long)it * num_clusters) + 2 * slot + (long long)rank;
      tile_next = tile;
      m = tile * 128 + row;
      valid = m < M_eff;
      if (valid && row_ids != nullptr) m = row_ids[m];   // compacted row -> original (ray, sample) row
      float p[3] = {0.f, 0.f, 0.f};
      d[0] = d[1] = d[2] = 0.f;
      if (valid) {
        long long ray = m / S;
        float z = z_vals[m];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          d[c] = rays_d[ray * 3 + c];
          p[c] = __fadd_rn(rays_o[ray * 3 + c], __fmul_rn(d[c], z));
        }
      }
      float f[64];
      pos_enc_row<kLx>(p, f);
      f[63] = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) pe_pk[i] = pack_16x2<kF16>(f[2 * i], f[2 * i + 1]);
    };
    if (my_quads > 0) prepare_tile(0);

    for (int it = 0; it < my_quads; ++it) {
      const long long m_cur = m;
      const bool valid_cur = valid;
      const bool tile_ok = kSave && tile_next < n_tiles && !(dbg & 1);
      const bool mask_ok = kSave && tile_next < n_tiles && !(dbg & 2);
      unsigned char* acts_tile = kSave ? acts + (size_t)tile_next * ((size_t)kActBlocks * kBlockBytes) : nullptr;
      const size_t mask_row = (size_t)tile_next * 128 + (size_t)row;
      const size_t mask_plane = (size_t)n_tiles * 128 * kMaskWords;
      float d_cur[3] = {d[0], d[1], d[2]};
      if (kSave) bulk_store_warp_reads_done();    // the previous tile's stores have left this warp's rows (PE tile, A tile)
      {  // xyz PE tile -> shared memory (swizzled 16-byte chunks), then hand the slot to the MMA issuer
        const uint32_t row_base = pe_base + (uint32_t)row * 128u;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v4(row_base + (uint32_t)((j ^ r7) << 4), pe_pk[4 * j], pe_pk[4 * j + 1], pe_pk[4 * j + 2], pe_pk[4 * j + 3]);
      }
      fence_proxy_async_smem();
      if (kSave) named_bar_sync(1 + slot, 128);
      if (!kSave || save_leader) mbar_arrive_remote(b_ready_leader);
      if (kSave && tile_ok)
        bulk_store_warp_rows(acts_tile + (size_t)kActPe * kBlockBytes + (size_t)w4 * 4096u, pe_base + (uint32_t)w4 * 4096u, 1);
      float sigma = 0.f;
      for (int stage = 0; stage < I::kN; ++stage) {
        const uint32_t bseq = (uint32_t)it * I::kN + (uint32_t)stage;
        const uint32_t bbuf = bseq & 1u;
        const float4* bias4 = reinterpret_cast<const float4*>(smem_dyn + kOffBias + bbuf * 1024u);
        if (stage == I::kLast && it + 1 < my_quads) prepare_tile(it + 1);   // overlaps the last stage's MMAs
        mbar_wait(bar(BAR_BFULL + bbuf), (bseq >> 1) & 1u, 0x500 + stage);
        mbar_wait(b_full, full_phase, 0x100 + stage);
        full_phase ^= 1;
        tc_fence_after();
        if (kTimeline && tl && blockIdx.x == 0 && it < 4 && row == 0) tl[((it * 10 + stage) * 2 + slot) * 4 + 2] = clock64();
        if (kDump) {   // diagnostic: fp32 post-activation outputs of rows 0..127 of the whole problem
          if (m_cur - row == 0) {
            const int ncb = stage == I::kLast ? 4 : 8;
            for (int cb = 0; cb < ncb; ++cb) {
              uint32_t v[32];
              tmem_ld32(t_acc + (uint32_t)cb * 32u, v);
              tmem_ld_wait();
              pin32(v);
              for (int i = 0; i < 32; ++i) {
                float x = __uint_as_float(v[i]) + tail[kTailBias + stage * 256 + cb * 32 + i];
                if (stage != 8) x = fmaxf(x, 0.f);
                stage_dump[((size_t)stage * 128 + row) * 256 + cb * 32 + i] = x;
              }
            }
          }
        }
        if (kSave) bulk_store_warp_reads_done();   // the previous stage's stores have finished reading this warp's rows of A
        if (stage < I::kLast) {
          uint32_t mw[8];
          unsigned char* bulk_g = nullptr;
          uint32_t bulk_s = 0u;
          if (kSave && tile_ok && stage < 8) {   // this warp's rows of the stage output (= the next A tile) in the activation store
            bulk_g = acts_tile + (size_t)act_h(stage) * kBlockBytes + (size_t)w4 * 4096u;
            bulk_s = smem_base + kOffA + (uint32_t)slot * kABytes + (uint32_t)w4 * 4096u;
          }
          if (!kFused && stage == 7) epi_stage256<1, kSave, kF16>(t_acc, bias4, a_row_base, r7, tail + kTailAlphaW, sigma, mw, bulk_g, bulk_s);
          else if (!kFused && stage == 8) epi_stage256<2, false, kF16>(t_acc, bias4, a_row_base, r7, nullptr, sigma);
          else epi_stage256<0, kSave, kF16>(t_acc, bias4, a_row_base, r7, nullptr, sigma, mw, bulk_g, bulk_s);
          if (stage == I::kLast - 1) {  // dir PE replaces the xyz PE tile (dead after stage 5) for the views stage
            float f[32];
            pos_enc_row<kLd>(d_cur, f);
#pragma unroll
            for (int i = kChD; i < 32; ++i) f[i] = 0.f;
            store_row_chunks<4, kF16>(pe_base, row, f);
          }
          const bool tl_on = kTimeline && kSave && tl && blockIdx.x == 0 && it < 4 && row == 0;
          if (tl_on) tl[720 + ((it * 10 + stage) * 2 + slot) * 4 + 0] = clock64();
          tc_fence_before();
          fence_proxy_async_smem();
          if (kSave) named_bar_sync(1 + slot, 128);
          if (tl_on) tl[720 + ((it * 10 + stage) * 2 + slot) * 4 + 1] = clock64();
          if (!kSave || save_leader) mbar_arrive_remote(b_ready_leader);
          mbar_arrive(bar(BAR_BEMPTY + bbuf));
          if (tl_on) tl[720 + ((it * 10 + stage) * 2 + slot) * 4 + 2] = clock64();
          if (kSave && tile_ok && stage == I::kLast - 1)   // the dir PE tile this stage's epilogue wrote (store_row_chunks above)
            bulk_store_warp_rows(acts_tile + (size_t)kActDpe * kBlockBytes + (size_t)w4 * 4096u, pe_base + (uint32_t)w4 * 4096u, 1);
          if (kSave && mask_ok && stage < 8) {   // relu sign bits of this stage for the dgrad epilogue (after the
            // hand-off: a global store in front of the arrive would sit under its release fence)
            uint4* dst = reinterpret_cast<uint4*>(masks + (size_t)stage * mask_plane + mask_row * kMaskWords);
            dst[0] = make_uint4(mw[0], mw[1], mw[2], mw[3]);
            dst[1] = make_uint4(mw[4], mw[5], mw[6], mw[7]);
          }
          if (kTimeline && tl && blockIdx.x == 0 && it < 4 && row == 0) tl[((it * 10 + stage) * 2 + slot) * 4 + 3] = clock64();
        } else {
          // stage 9: views_linears.0 (128 wide, relu) -> rgb_linear on CUDA cores (network.py:66-69)
          float r0 = 0.f, r1 = 0.f, r2 = 0.f;
          uint32_t hvm[4] = {0u, 0u, 0u, 0u};
#pragma unroll 2
          for (int cb = 0; cb < 4; ++cb) {
            uint32_t ch[4] = {0u, 0u, 0u, 0u};
            uint32_t v[32];
            tmem_ld32(t_acc + (uint32_t)cb * 32u, v);
            tmem_ld_wait();
            pin32(v);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              int n = cb * 32 + q * 4;
              float4 b4 = bias4[cb * 8 + q];
              float4 w0 = __ldg(reinterpret_cast<const float4*>(tail + kTailRgbW + n));
              float4 w1 = __ldg(reinterpret_cast<const float4*>(tail + kTailRgbW + 128 + n));
              float4 w2 = __ldg(reinterpret_cast<const float4*>(tail + kTailRgbW + 256 + n));
              const float x0 = __uint_as_float(v[q * 4 + 0]) + b4.x, x1 = __uint_as_float(v[q * 4 + 1]) + b4.y;
              const float x2 = __uint_as_float(v[q * 4 + 2]) + b4.z, x3 = __uint_as_float(v[q * 4 + 3]) + b4.w;
              float h0 = fmaxf(x0, 0.f);
              float h1 = fmaxf(x1, 0.f);
              float h2 = fmaxf(x2, 0.f);
              float h3 = fmaxf(x3, 0.f);
              r0 = fmaf(h0, w0.x, r0); r0 = fmaf(h1, w0.y, r0); r0 = fmaf(h2, w0.z, r0); r0 = fmaf(h3, w0.w, r0);
              r1 = fmaf(h0, w1.x, r1); r1 = fmaf(h1, w1.y, r1); r1 = fmaf(h2, w1.z, r1); r1 = fmaf(h3, w1.w, r1);
              r2 = fmaf(h0, w2.x, r2); r2 = fmaf(h1, w2.y, r2); r2 = fmaf(h2, w2.z, r2); r2 = fmaf(h3, w2.w, r2);
              if (kSave) {   // bf16 row of relu(views) into the (now dead) A tile, blocks 0..1, + its sign bits
                ch[0] = __funnelshift_l(__float_as_uint(x0), ch[0], 1); ch[1] = __funnelshift_l(__float_as_uint(x1), ch[1], 1);
                ch[2] = __funnelshift_l(__float_as_uint(x2), ch[2], 1); ch[3] = __funnelshift_l(__float_as_uint(x3), ch[3], 1);
                uint2 pk = make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
                const int col = n & 63;   // column inside block (n >> 6)
                *reinterpret_cast<uint2*>(a_row_base + (n >> 6) * 16384 + (((col >> 3) ^ r7) << 4) + (col & 7) * 2) = pk;
              }
            }
            if (kSave) hvm[cb] = ch[0] | (ch[1] << 8) | (ch[2] << 16) | (ch[3] << 24);
          }
          if (kSave) {
            if (mask_ok)
              *reinterpret_cast<uint4*>(masks + (size_t)8 * mask_plane + mask_row * kMaskWords) = make_uint4(hvm[0], hvm[1], hvm[2], hvm[3]);
            if (tile_ok)
              bulk_store_warp_rows(acts_tile + (size_t)kActHv * kBlockBytes + (size_t)w4 * 4096u,
                                   smem_base + kOffA + (uint32_t)slot * kABytes + (uint32_t)w4 * 4096u, 2);
          }
          if (kFused) {   // sigma_raw = accumulator column 128 (+ alpha_b, staged as bias element 128)
            uint32_t v[32];
            tmem_ld32(t_acc + 128u, v);
            tmem_ld_wait();
            pin32(v);
            // columns 128 / 129 = h7 . (hi / lo bf16 parts of alpha_linear.weight), pack.cu
            sigma = (__uint_as_float(v[0]) + __uint_as_float(v[1])) + bias4[32].x - tail[kTailAlphaB];   // alpha_b is added below
          }
          tc_fence_before();
          mbar_arrive(bar(BAR_BEMPTY + bbuf));
          if (valid_cur) {
            float4 o = make_float4(r0 + tail[kTailRgbB + 0], r1 + tail[kTailRgbB + 1], r2 + tail[kTailRgbB + 2],
                                   sigma + tail[kTailAlphaB]);
            *reinterpret_cast<float4*>(raw + m_cur * 4) = o;
          }
        }
      }
    }
    if (kSave) bulk_store_warp_drain();
  } else if (warp == 8) {
    // =========================== producer: bias block + this CTA's half of every weight chunk ==========
    uint32_t seq = 0, bseq = 0;
    for (int it = 0; it < my_quads; ++it) {
      for (int stage = 0; stage < I::kN; ++stage, ++bseq) {
        {
          const uint32_t bbuf = bseq & 1u;
          if (lane == 0) mbar_wait(bar(BAR_BEMPTY + bbuf), ((bseq >> 1) & 1u) ^ 1u, 0x600 + stage);
          __syncwarp();
          const float4* src4 = reinterpret_cast<const float4*>(I::bias(packed, stage));
          float4 v0 = __ldg(src4 + lane), v1 = __ldg(src4 + 32 + lane);
          const uint32_t dst = smem_base + kOffBias + bbuf * 1024u + (uint32_t)lane * 16u;
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(v0.x), "f"(v0.y), "f"(v0.z), "f"(v0.w) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 512u), "f"(v1.x), "f"(v1.y), "f"(v1.z), "f"(v1.w) : "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(BAR_BFULL + bbuf));
        }
        if (lane == 0) {
          const uint32_t chunk_bytes = I::chunk_bytes(stage);               // full N rows
          const uint32_t half = chunk_bytes >> 1;                           // this CTA's N/2 rows
          const unsigned char* src = packed + I::stage_off(stage) + (size_t)rank * half;
          const int nch = I::chunks(stage);
          for (int c = 0; c < nch; ++c, ++seq) {
            const uint32_t pos = seq % kRing, phase = (seq / kRing) & 1u;
            mbar_wait(bar(BAR_WEMPTY + pos), phase ^ 1u, 0x200 + stage);
            mbar_arrive_expect_tx(bar(BAR_WFULL + pos), half);
            bulk_g2s(smem_base + kOffW + pos * kWStageBytes, src + (size_t)c * chunk_bytes, half, bar(BAR_WFULL + pos));
          }
        }
        __syncwarp();
      }
    }
  } else if (rank == 1) {
    // =========================== relay (peer CTA): my half of chunk seq has landed ===========================
    if (lane == 0) {
      uint32_t seq = 0;
      for (int it = 0; it < my_quads; ++it)
        for (int stage = 0; stage < I::kN; ++stage) {
          const int nch = I::chunks(stage);
          for (int c = 0; c < nch; ++c, ++seq) {
            const uint32_t pos = seq % kRing, phase = (seq / kRing) & 1u;
            mbar_wait(bar(BAR_WFULL + pos), phase, 0x700 + stage);
            mbar_arrive_remote(mapa(bar(BAR_WFULL + pos), 0));
          }
        }
    }
    __syncwarp();
  } else {
    // =========================== MMA issuer (leader CTA) ===========================
    // The whole warp walks the schedule (waits are warp-uniform); one elected lane issues the MMAs and
    // commits.  A single thread's instruction latency is the limit here (measured: ~200 cycles per MMA
    // when every descriptor is rebuilt with shifts/masks inside a divergent lane-0 region), so the
    // descriptors are reduced to "constant high word + precomputed low word + 2*kstep".
    uint32_t seq = 0, ready_phase = 0;
    const uint32_t desc_hi = (uint32_t)(umma_desc_sw128(0) >> 32);
    const uint32_t lo_flags = (uint32_t)(umma_desc_sw128(0) & 0xFFFFFFFFu);   // LBO field
    const uint32_t a_lo0 = lo_flags | ((smem_base + kOffA) >> 4);     // + slot*4096 + kblock*1024 + 2*k
    const uint32_t pe_lo0 = lo_flags | ((smem_base + kOffPe) >> 4);   // + slot*1024 + 2*k
    const uint32_t w_lo0 = lo_flags | ((smem_base + kOffW) >> 4);     // + pos*1024 + 2*k
    for (int it = 0; it < my_quads; ++it) {
      for (int stage = 0; stage < I::kN; ++stage) {
        const int nch = I::chunks(stage);
        // kind::f16 operand format field: 1 = bf16, 0 = fp16 (bits 7-9 / 10-12 of the instruction descriptor)
        const uint32_t idesc = kF16 ? (umma_idesc_bf16(256, I::n(stage)) & ~((1u << 7) | (1u << 10))) : umma_idesc_bf16(256, I::n(stage));
        // chunk groups of at most kRing chunks: slot 0 runs the group, then slot 1 runs it and releases it
        for (int g0 = 0; g0 < nch; g0 += kRing) {
          const int g1 = (g0 + kRing < nch) ? g0 + kRing : nch;
#pragma unroll 1
          for (int slot = 0; slot < 2; ++slot) {
            if (g0 == 0) {
              if (kTimeline && tl && blockIdx.x == 0 && it < 4 && lane == 0) tl[((it * 10 + stage) * 2 + slot) * 4 + 0] = clock64();
              mbar_wait_cluster(bar(BAR_AREADY + slot), ready_phase, 0x400 + stage * 2 + slot);
              tc_fence_after();
              if (kTimeline && tl && blockIdx.x == 0 && it < 4 && lane == 0) tl[((it * 10 + stage) * 2 + slot) * 4 + 1] = clock64();
            }
            const uint32_t d_tmem = tmem_base + (uint32_t)slot * 256u;
#pragma unroll 1
            for (int c = g0; c < g1; ++c) {
              const uint32_t cs = seq + (uint32_t)(c - g0);
              const uint32_t pos = cs % kRing, phase = (cs / kRing) & 1u;
              if (slot == 0) {
                long long tw0 = 0;
                if (kTimeline && tl && blockIdx.x == 0 && it < 4 && lane == 0) tw0 = clock64();
                mbar_wait_cluster(bar(BAR_WFULL + pos), phase, 0x300 + stage);
                tc_fence_after();
                if (kTimeline && tl && blockIdx.x == 0 && it < 4 && lane == 0) {
                  tl[320 + ((it * 10 + stage) * 5 + c) * 2 + 0] = tw0;
                  tl[320 + ((it * 10 + stage) * 5 + c) * 2 + 1] = clock64();
                }
              }
              bool from_pe;
              int kblock, ksteps;
              chunk_src(stage, I::kLast, c, from_pe, kblock, ksteps);
              const uint32_t a_lo = from_pe ? (pe_lo0 + (uint32_t)slot * (kPeBytes >> 4))
                                            : (a_lo0 + (uint32_t)slot * (kABytes >> 4) + (uint32_t)kblock * 1024u);
              const uint32_t b_lo = w_lo0 + pos * (kWStageBytes >> 4);
              if (elect_one()) {
                const uint64_t hi64 = (uint64_t)desc_hi << 32;
                umma_bf16_ss_2cta(d_tmem, hi64 | (a_lo + 0u), hi64 | (b_lo + 0u), idesc, c > 0 ? 1u : 0u);
                umma_bf16_ss_2cta(d_tmem, hi64 | (a_lo + 2u), hi64 | (b_lo + 2u), idesc, 1u);
                if (ksteps == 4) {
                  umma_bf16_ss_2cta(d_tmem, hi64 | (a_lo + 4u), hi64 | (b_lo + 4u), idesc, 1u);
                  umma_bf16_ss_2cta(d_tmem, hi64 | (a_lo + 6u), hi64 | (b_lo + 6u), idesc, 1u);
                }
                if (slot == 1) umma_commit_2cta(bar(BAR_WEMPTY + pos), 3);                 // both CTAs' producers
                if (c == nch - 1) umma_commit_2cta(bar(BAR_ACCFULL + slot), 3);           // both CTAs' epilogue groups
              }
              __syncwarp();
            }
          }
          seq += (uint32_t)(g1 - g0);
        }
        ready_phase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA may exit (or free TMEM) while its peer still multicasts into it
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

}  // namespace tc2

int launch_mlp_bf16(const void* packed, const float* rays_o, const float* rays_d, const float* z_vals, int n_rays,
                    int n_samples, float* raw, float* stage_dump, void* acts, void* masks, const int* row_ids,
                    const int* n_active, bool f16, cudaStream_t st) {
  using namespace tc2;
  NB_CHECK_ARG(!f16 || (!acts && !stage_dump), "fp16 operands: inference launches only");
  int dev = 0, sms = 0;
  NB_CUDA(cudaGetDevice(&dev));
  NB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static bool attr_set[64] = {};   // the opt-in to 226 KB of dynamic shared memory is per device and sticky
  NB_CHECK_ARG(dev >= 0 && dev < 64, "mlp_forward: device ordinal %d out of range", dev);
  if (!attr_set[dev]) {
    NB_CUDA(cudaFuncSetAttribute(mlp_bf16_tc2_kernel<false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    NB_CUDA(cudaFuncSetAttribute(mlp_bf16_tc2_kernel<false, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    NB_CUDA(cudaFuncSetAttribute(mlp_bf16_tc2_kernel<true, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    NB_CUDA(cudaFuncSetAttribute(mlp_bf16_tc2_kernel<false, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    NB_CUDA(cudaFuncSetAttribute(mlp_bf16_tc2_kernel<false, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    NB_CUDA(cudaFuncSetAttribute(mlp_bf16_tc2_kernel<false, false, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    attr_set[dev] = true;
  }
  const char* tl_env = getenv("NERFB200_TIMELINE");
  unsigned long long* tl = nullptr;
  if (tl_env && !stage_dump && !f16) {
    cudaMalloc(&tl, (320 + 400 + 320) * 8);
    cudaMemset(tl, 0, (320 + 400 + 320) * 8);
  }
  long long M = (long long)n_rays * n_samples;
  long long quads = (M + 511) / 512;
  int clusters = (int)(quads < sms / 2 ? quads : sms / 2);
  if (const char* mc = getenv("NERFB200_MAX_CLUSTERS")) {   // debug: restrict the number of active TPCs
    int v = atoi(mc);
    if (v > 0 && v < clusters) clusters = v;
  }
  if (f16)
    mlp_bf16_tc2_kernel<false, false, false, true, true><<<2 * clusters, kThreads, kSmemBytes, st>>>(
        (const unsigned char*)packed, rays_o, rays_d, z_vals, M, n_samples, (int)quads, raw, nullptr, nullptr, nullptr, nullptr, row_ids, n_active);
  else if (acts && tl)
    mlp_bf16_tc2_kernel<false, true, true, true><<<2 * clusters, kThreads, kSmemBytes, st>>>(
        (const unsigned char*)packed, rays_o, rays_d, z_vals, M, n_samples, (int)quads, raw, nullptr, tl,
        (unsigned char*)acts, (uint32_t*)masks, nullptr, nullptr, getenv("NERFB200_DBG") ? atoi(getenv("NERFB200_DBG")) : 0);
  else if (acts)
    mlp_bf16_tc2_kernel<false, false, true, true><<<2 * clusters, kThreads, kSmemBytes, st>>>(
        (const unsigned char*)packed, rays_o, rays_d, z_vals, M, n_samples, (int)quads, raw, nullptr, nullptr,
        (unsigned char*)acts, (uint32_t*)masks, nullptr, nullptr);
  else if (stage_dump)
    mlp_bf16_tc2_kernel<true, false, false, false><<<2 * clusters, kThreads, kSmemBytes, st>>>(
        (const unsigned char*)packed, rays_o, rays_d, z_vals, M, n_samples, (int)quads, raw, stage_dump, nullptr, nullptr, nullptr, nullptr, nullptr);
  else if (tl)
    mlp_bf16_tc2_kernel<false, true, false, true><<<2 * clusters, kThreads, kSmemBytes, st>>>(
        (const unsigned char*)packed, rays_o, rays_d, z_vals, M, n_samples, (int)quads, raw, nullptr, tl, nullptr, nullptr, nullptr, nullptr);
  else
    mlp_bf16_tc2_kernel<false, false, false, true><<<2 * clusters, kThreads, kSmemBytes, st>>>(
        (const unsigned char*)packed, rays_o, rays_d, z_vals, M, n_samples, (int)quads, raw, nullptr, nullptr, nullptr, nullptr, row_ids, n_active);
  NB_LAUNCH_OK("mlp_bf16_tc2_kernel");
  if (tl) {   // debug only (NERFB200_TIMELINE=<file>): dump cluster 0's handshake timestamps
    unsigned long long host[320 + 400 + 320];
    cudaStreamSynchronize(st);
    cudaMemcpy(host, tl, sizeof(host), cudaMemcpyDeviceToHost);
    cudaFree(tl);
    FILE* f = fopen(tl_env, "w");
    if (f) {
      for (int i = 0; i < 4 * 10 * 2; ++i)
        fprintf(f, "%d %d %d %llu %llu %llu %llu\n", i / 20, (i / 2) % 10, i % 2, host[i * 4], host[i * 4 + 1],
                host[i * 4 + 2], host[i * 4 + 3]);
      for (int i = 0; i < 4 * 10 * 2; ++i)   // training epilogue detail: it stage 8<slot> after-compute after-barrier after-arrive 0
        fprintf(f, "%d %d 8%d %llu %llu %llu 0\n", i / 20, (i / 2) % 10, i % 2, host[720 + i * 4], host[720 + i * 4 + 1], host[720 + i * 4 + 2]);
      for (int i = 0; i < 200; ++i)   // weight-chunk waits of slot 0's pass: it stage 9 chunk start end 0 0
        fprintf(f, "%d %d 9%d %llu %llu 0 0\n", i / 50, (i / 5) % 10, i % 5, host[320 + i * 2], host[320 + i * 2 + 1]);
      fclose(f);
    }
  }
  return 0;
}

}  // namespace nb
