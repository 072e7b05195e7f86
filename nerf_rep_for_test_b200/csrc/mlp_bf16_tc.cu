// a3, bf16 performance mode: positional encoding fused into the 8x256 NeRF MLP on the 5th-gen
// tensor cores (tcgen05.mma kind::f16, fp32 accumulators in TMEM, weights streamed by the TMA
// engine, activations never leave the SM).
// Reference: volume_renderer.py:270-284 -> freq.py:23-26 -> network.py:49-74.
//
// 1-CTA variant (cta_group::1), kept for A/B comparison (NERFB200_TC_VARIANT=1); the default is the
// CTA-pair kernel of mlp_bf16_tc2.cu.  Persistent kernel, one CTA per SM, 10 warps:
//   warps 0-3  epilogue group of tile slot 0      warps 4-7  epilogue group of tile slot 1
//   warp 8     weight producer (cp.async.bulk -> 2-stage smem ring, mbarrier complete_tx)
//   warp 9     MMA issuer (one thread issues every tcgen05.mma), TMEM allocator
// A CTA works on two 128-row tiles at once (slot 0/1, 128 TMEM lanes x 256 fp32 columns each =
// all 512 columns).  The ten stages of mlp_layout.cuh are walked by both slots in lock step and
// the MMAs are interleaved per weight K-chunk: chunk c feeds slot 0's four K=16 MMAs, then slot
// 1's, then the ring stage is released -- every weight byte fetched from L2 is used for 256 rows.
// Slot 0 finishes a stage 4 MMAs before slot 1, so its epilogue (tcgen05.ld -> +bias -> ReLU ->
// bf16 -> swizzled st.shared of the next stage's A operand) overlaps slot 1's tail and vice versa.
//
// Shared memory (1024-B aligned, SWIZZLE_128B K-major everywhere):
//   A[slot]   64 KB  [128 rows][256 k] bf16 as four 16 KB K-blocks (row r at r*128 B)
//   PE[slot]  16 KB  [128][64] bf16: xyz PE for stages 0 and 5, later reused for the dir PE
//   W[2]      32 KB  weight ring stages, each one K-chunk [N rows][64 k] of the packed image
#include <cuda_bf16.h>
#include <stdlib.h>

#include "mlp_tc_common.cuh"

namespace nb {
using namespace ptx;

constexpr int kTcThreads = 320;
constexpr int kTileRows = 128;
constexpr int kWStages = 2;
constexpr uint32_t kABytes = 65536, kPeBytes = 16384, kWStageBytes = 32768;
constexpr uint32_t kOffA = 0;
constexpr uint32_t kOffPe = kOffA + 2 * kABytes;
constexpr uint32_t kOffW = kOffPe + 2 * kPeBytes;
constexpr uint32_t kOffBar = kOffW + kWStages * kWStageBytes;  // 229376
constexpr uint32_t kOffBias = kOffBar + 128;                   // 2 x 1 KB: fp32 bias of the stage in flight
constexpr uint32_t kSmemBytes = kOffBias + 2048;               // 231552 <= 232448 (227 KB)

// barrier slots (8 B each) at kOffBar
enum { BAR_WFULL = 0, BAR_WEMPTY = 2, BAR_AREADY = 4, BAR_ACCFULL = 6, BAR_BFULL = 8, BAR_BEMPTY = 10, BAR_COUNT = 12 };

template <bool kDump>
__global__ void __launch_bounds__(kTcThreads, 1)
mlp_bf16_tc_kernel(const unsigned char* __restrict__ packed, const float* __restrict__ rays_o,
                   const float* __restrict__ rays_d, const float* __restrict__ z_vals, long long M, int S,
                   int num_pairs, float* __restrict__ raw, float* __restrict__ stage_dump, int dbg,
                   unsigned long long* __restrict__ tl) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  const uint32_t smem_base = smem_u32(smem_dyn);
  if ((smem_base & 1023u) != 0) __trap();   // SWIZZLE_128B operands need 1024-byte alignment
  const uint32_t bar_base = smem_base + kOffBar;
  const uint32_t tmem_slot = bar_base + BAR_COUNT * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto bar = [&](int i) { return bar_base + (uint32_t)i * 8u; };

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWStages; ++i) {
      mbar_init(bar(BAR_WFULL + i), 1);
      mbar_init(bar(BAR_WEMPTY + i), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(BAR_AREADY + s), 128);
      mbar_init(bar(BAR_ACCFULL + s), 1);
      mbar_init(bar(BAR_BFULL + s), 1);
      mbar_init(bar(BAR_BEMPTY + s), 256);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int my_pairs = (num_pairs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const float* tail = reinterpret_cast<const float*>(packed + kBf16TailOff);

  if (warp < 8) {
    // =========================== epilogue groups ===========================
    const int slot = warp >> 2;
    const int w4 = warp & 3;
    const int row = w4 * 32 + lane;
    const uint32_t pe_base = smem_base + kOffPe + (uint32_t)slot * kPeBytes;
    const uint32_t t_acc = tmem_base + ((uint32_t)(w4 * 32) << 16) + (uint32_t)slot * 256u;
    const uint32_t b_ready = bar(BAR_AREADY + slot), b_full = bar(BAR_ACCFULL + slot);
    uint32_t full_phase = 0;
    for (int it = 0; it < my_pairs; ++it) {
      const long long tile = 2LL * ((long long)blockIdx.x + (long long)it * gridDim.x) + slot;
      const long long m = tile * kTileRows + row;
      const bool valid = m < M;
      float p[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 0.f};
      if (valid) {
        long long ray = m / S;
        float z = z_vals[m];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          d[c] = rays_d[ray * 3 + c];
          p[c] = __fadd_rn(rays_o[ray * 3 + c], __fmul_rn(d[c], z));
        }
      }
      {
        float f[64];
        pos_enc_row<kLx>(p, f);
        f[63] = 0.f;
        store_row_chunks<8>(pe_base, row, f);
      }
      fence_proxy_async_smem();
      mbar_arrive(b_ready);
      float sigma = 0.f;
      const int r7 = row & 7;
      unsigned char* a_row_base = smem_dyn + kOffA + (uint32_t)slot * kABytes + (uint32_t)row * 128u;
      for (int stage = 0; stage < kStages; ++stage) {
        // the stage's fp32 bias block, staged into shared memory by the producer warp
        const uint32_t bseq = (uint32_t)it * kStages + (uint32_t)stage;
        const uint32_t bbuf = bseq & 1u;
        const float4* bias4 = reinterpret_cast<const float4*>(smem_dyn + kOffBias + bbuf * 1024u);
        mbar_wait(bar(BAR_BFULL + bbuf), (bseq >> 1) & 1u, 0x500 + stage);
        mbar_wait(b_full, full_phase, 0x100 + stage);
        full_phase ^= 1;
        tc_fence_after();
        if (tl && blockIdx.x == 0 && it < 4 && row == 0) tl[((it * 10 + stage) * 2 + slot) * 4 + 2] = clock64();
        if (kDump && tile == 0) {   // diagnostic: fp32 post-activation outputs of rows 0..127
          const int ncb = stage == 9 ? 4 : 8;
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
        if ((dbg & 2) && it > 0) {
          tc_fence_before();
          if (stage < 9) { fence_proxy_async_smem(); mbar_arrive(b_ready); }
          mbar_arrive(bar(BAR_BEMPTY + bbuf));
          continue;
        }
        if (stage < 9) {
          if (stage == 7) epi_stage256<1>(t_acc, bias4, a_row_base, r7, tail + kTailAlphaW, sigma);
          else if (stage == 8) epi_stage256<2>(t_acc, bias4, a_row_base, r7, nullptr, sigma);
          else epi_stage256<0>(t_acc, bias4, a_row_base, r7, nullptr, sigma);
          if (stage == 8) {  // dir PE replaces the xyz PE tile (dead after stage 5) for stage 9
            float f[32];
            pos_enc_row<kLd>(d, f);
#pragma unroll
            for (int i = kChD; i < 32; ++i) f[i] = 0.f;
            store_row_chunks<4>(pe_base, row, f);
          }
          tc_fence_before();
          fence_proxy_async_smem();
          mbar_arrive(b_ready);
          mbar_arrive(bar(BAR_BEMPTY + bbuf));
          if (tl && blockIdx.x == 0 && it < 4 && row == 0) tl[((it * 10 + stage) * 2 + slot) * 4 + 3] = clock64();
        } else {
          // stage 9: views_linears.0 (128 wide, relu) -> rgb_linear on CUDA cores (network.py:66-69)
          float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll 2
          for (int cb = 0; cb < 4; ++cb) {
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
              float h0 = fmaxf(__uint_as_float(v[q * 4 + 0]) + b4.x, 0.f);
              float h1 = fmaxf(__uint_as_float(v[q * 4 + 1]) + b4.y, 0.f);
              float h2 = fmaxf(__uint_as_float(v[q * 4 + 2]) + b4.z, 0.f);
              float h3 = fmaxf(__uint_as_float(v[q * 4 + 3]) + b4.w, 0.f);
              r0 = fmaf(h0, w0.x, r0); r0 = fmaf(h1, w0.y, r0); r0 = fmaf(h2, w0.z, r0); r0 = fmaf(h3, w0.w, r0);
              r1 = fmaf(h0, w1.x, r1); r1 = fmaf(h1, w1.y, r1); r1 = fmaf(h2, w1.z, r1); r1 = fmaf(h3, w1.w, r1);
              r2 = fmaf(h0, w2.x, r2); r2 = fmaf(h1, w2.y, r2); r2 = fmaf(h2, w2.z, r2); r2 = fmaf(h3, w2.w, r2);
            }
          }
          tc_fence_before();
          mbar_arrive(bar(BAR_BEMPTY + bbuf));
          if (valid) {
            float4 o = make_float4(r0 + tail[kTailRgbB + 0], r1 + tail[kTailRgbB + 1], r2 + tail[kTailRgbB + 2],
                                   sigma + tail[kTailAlphaB]);
            *reinterpret_cast<float4*>(raw + m * 4) = o;
          }
        }
      }
    }
  } else if (warp == 8) {
    // =========================== weight producer ===========================
    uint32_t ring = 0, phase = 0, bseq = 0;
    for (int it = 0; it < my_pairs; ++it) {
      for (int stage = 0; stage < kStages; ++stage, ++bseq) {
        {  // fp32 bias block of this stage -> bias buffer bseq&1 (all 32 lanes, 2 x float4 each)
          const uint32_t bbuf = bseq & 1u;
          if (lane == 0) mbar_wait(bar(BAR_BEMPTY + bbuf), ((bseq >> 1) & 1u) ^ 1u, 0x600 + stage);
          __syncwarp();
          const float4* src4 = reinterpret_cast<const float4*>(tail + kTailBias + stage * 256);
          float4 v0 = __ldg(src4 + lane), v1 = __ldg(src4 + 32 + lane);
          const uint32_t dst = smem_base + kOffBias + bbuf * 1024u + (uint32_t)lane * 16u;
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(v0.x), "f"(v0.y), "f"(v0.z), "f"(v0.w) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 512u), "f"(v1.x), "f"(v1.y), "f"(v1.z), "f"(v1.w) : "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(BAR_BFULL + bbuf));
        }
        if (lane == 0) {
          const uint32_t bytes = (uint32_t)bf16_chunk_bytes(stage);
          const unsigned char* src = packed + bf16_stage_off(stage);
          const int nch = stage_chunks(stage);
          for (int c = 0; c < nch; ++c) {
            mbar_wait(bar(BAR_WEMPTY + ring), phase ^ 1, 0x200 + stage);
            if ((dbg & 1) && it > 0) { mbar_arrive(bar(BAR_WFULL + ring)); }
            else {
            mbar_arrive_expect_tx(bar(BAR_WFULL + ring), bytes);
            bulk_g2s(smem_base + kOffW + ring * kWStageBytes, src + (size_t)c * bytes, bytes, bar(BAR_WFULL + ring));
            }
            if (++ring == kWStages) { ring = 0; phase ^= 1; }
          }
        }
        __syncwarp();
      }
    }
  } else {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      uint32_t ring = 0, phase = 0, ready_phase = 0;
      for (int it = 0; it < my_pairs; ++it) {
        for (int stage = 0; stage < kStages; ++stage) {
          const int nch = stage_chunks(stage);
          const uint32_t idesc = umma_idesc_bf16(128, stage_n(stage));
          for (int c = 0; c < nch; ++c) {
            mbar_wait(bar(BAR_WFULL + ring), phase, 0x300 + stage);
            tc_fence_after();
            const uint32_t w_addr = smem_base + kOffW + ring * kWStageBytes;
            // which on-chip buffer holds K-chunk c of this stage's input (mlp_layout.cuh)
            bool from_pe;
            int kblock;
            if (stage == 0) { from_pe = true; kblock = 0; }
            else if (stage == 5) { from_pe = (c == 0); kblock = c - 1; }
            else if (stage == 9) { from_pe = (c == 4); kblock = c; }
            else { from_pe = false; kblock = c; }
            const int ksteps = (stage == 9 && c == 4) ? 2 : 4;
#pragma unroll 1
            for (int slot = 0; slot < 2; ++slot) {
              if (c == 0) {
                if (tl && blockIdx.x == 0 && it < 4) tl[((it * 10 + stage) * 2 + slot) * 4 + 0] = clock64();
                mbar_wait(bar(BAR_AREADY + slot), ready_phase, 0x400 + stage * 2 + slot);
                tc_fence_after();
                if (tl && blockIdx.x == 0 && it < 4) tl[((it * 10 + stage) * 2 + slot) * 4 + 1] = clock64();
              }
              const uint32_t a_addr = from_pe ? (smem_base + kOffPe + (uint32_t)slot * kPeBytes)
                                              : (smem_base + kOffA + (uint32_t)slot * kABytes + (uint32_t)kblock * 16384u);
              const uint32_t d_tmem = tmem_base + (uint32_t)slot * 256u;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (k < ksteps)
                  umma_bf16_ss(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(w_addr + k * 32), idesc,
                               (c > 0 || k > 0) ? 1u : 0u);
              }
              if (c == nch - 1) umma_commit(bar(BAR_ACCFULL + slot));
            }
            umma_commit(bar(BAR_WEMPTY + ring));
            if (++ring == kWStages) { ring = 0; phase ^= 1; }
          }
          ready_phase ^= 1;
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_mlp_bf16_1cta(const void* packed, const float* rays_o, const float* rays_d, const float* z_vals,
                    int n_rays, int n_samples, float* raw, float* stage_dump, cudaStream_t st) {
  int dev = 0, sms = 0;
  NB_CUDA(cudaGetDevice(&dev));
  NB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  NB_CUDA(cudaFuncSetAttribute(mlp_bf16_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  NB_CUDA(cudaFuncSetAttribute(mlp_bf16_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  const char* tl_env = getenv("NERFB200_TIMELINE");
  unsigned long long* tl = nullptr;
  if (tl_env && !stage_dump) {
    cudaMalloc(&tl, 4 * 10 * 2 * 4 * 8);
    cudaMemset(tl, 0, 4 * 10 * 2 * 4 * 8);
  }
  const char* dbg_env = getenv("NERFB200_DEBUG_FLAGS");
  int dbg = dbg_env ? atoi(dbg_env) : 0;
  long long M = (long long)n_rays * n_samples;
  long long tiles = (M + kTileRows - 1) / kTileRows;
  int pairs = (int)((tiles + 1) / 2);
  int grid = pairs < sms ? pairs : sms;
  if (stage_dump)
    mlp_bf16_tc_kernel<true><<<grid, kTcThreads, kSmemBytes, st>>>((const unsigned char*)packed, rays_o, rays_d, z_vals,
                                                                  M, n_samples, pairs, raw, stage_dump, dbg, nullptr);
  else
    mlp_bf16_tc_kernel<false><<<grid, kTcThreads, kSmemBytes, st>>>((const unsigned char*)packed, rays_o, rays_d, z_vals,
                                                                   M, n_samples, pairs, raw, nullptr, dbg, tl);
  if (tl) {   // debug only (NERFB200_TIMELINE=<file>): dump CTA 0's handshake timestamps
    unsigned long long host[4 * 10 * 2 * 4];
    cudaStreamSynchronize(st);
    cudaMemcpy(host, tl, sizeof(host), cudaMemcpyDeviceToHost);
    cudaFree(tl);
    FILE* f = fopen(tl_env, "w");
    if (f) {
      for (int i = 0; i < 4 * 10 * 2; ++i)
        fprintf(f, "%d %d %d %llu %llu %llu %llu\n", i / 20, (i / 2) % 10, i % 2, host[i * 4], host[i * 4 + 1], host[i * 4 + 2], host[i * 4 + 3]);
      fclose(f);
    }
  }
  NB_LAUNCH_OK("mlp_bf16_tc_kernel");
  return 0;
}

}  // namespace nb
