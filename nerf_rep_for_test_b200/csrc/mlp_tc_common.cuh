// Device helpers shared by the tcgen05 MLP kernels (forward, split-fp16 forward, dgrad): bf16 / fp16 packing,
// swizzled A-operand stores, positional encoding, and the accumulator epilogue.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "mlp_layout.cuh"
#include "tc_ptx.cuh"

namespace nb {
using namespace ptx;

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// operand element type of the single-pass kernel: bf16 (NERFB200_MODE_BF16) or fp16 (NERFB200_MODE_FP16, saturating)
template <bool kF16>
__device__ __forceinline__ uint32_t pack_16x2(float lo, float hi) {
  if (!kF16) return pack_bf16x2(lo, hi);
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// two fp32 -> packed bf16x2 (lo in bits 0-15), optionally with ReLU folded into the conversion
template <bool kRelu, bool kF16 = false>
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) {
  uint32_t d;
  if (kF16) {
    if (kRelu) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  } else {
    if (kRelu) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  }
  return d;
}
// registers written by an in-flight tcgen05.ld must not be touched before tcgen05.wait::ld; this
// empty asm pins the 32 destination registers as "defined here" once the wait has retired
__device__ __forceinline__ void pin32(uint32_t (&r)[32]) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
               "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
  asm volatile("" : "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
               "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]),
               "+r"(r[31]));
}

// Epilogue of 32 accumulator columns of one row: + bias (fp32x2 adds, bias broadcast from shared
// memory), activation folded into the bf16 pack, swizzled 16-byte stores into the next stage's A
// operand.  MODE 0: ReLU; 1: ReLU + alpha_linear partial dot on the fp32 values (stage 7); 2: linear.
// Plain C++ shared-memory accesses (not volatile asm) so the compiler batches the bias loads.
// kMask (training forward): additionally returns the relu sign bits of the 32 columns in the bit order
// of train_layout.cuh (column 4s+k -> bit 8k+7-s, set = negative), built with one funnel shift per value.
template <int MODE, bool kMask = false, bool kF16 = false>
__device__ __forceinline__ void epi32(const uint32_t (&v)[32], const float4* __restrict__ bias4, unsigned char* out_row,
                                      int j0, int r7, const float* __restrict__ alpha_w, float& sigma,
                                      uint32_t* mask_word = nullptr) {
  uint32_t ch0 = 0, ch1 = 0, ch2 = 0, ch3 = 0;
  float4 b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = bias4[i];
  float4 aw[8];
  if (MODE == 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) aw[i] = __ldg(reinterpret_cast<const float4*>(alpha_w) + i);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 b0 = b[2 * q], b1 = b[2 * q + 1];
    float2 x0 = __fadd2_rn(make_float2(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1])), make_float2(b0.x, b0.y));
    float2 x1 = __fadd2_rn(make_float2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3])), make_float2(b0.z, b0.w));
    float2 x2 = __fadd2_rn(make_float2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5])), make_float2(b1.x, b1.y));
    float2 x3 = __fadd2_rn(make_float2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7])), make_float2(b1.z, b1.w));
    if (MODE == 1) {
      const float4 a0 = aw[2 * q], a1 = aw[2 * q + 1];
      sigma = fmaf(fmaxf(x0.x, 0.f), a0.x, sigma); sigma = fmaf(fmaxf(x0.y, 0.f), a0.y, sigma);
      sigma = fmaf(fmaxf(x1.x, 0.f), a0.z, sigma); sigma = fmaf(fmaxf(x1.y, 0.f), a0.w, sigma);
      sigma = fmaf(fmaxf(x2.x, 0.f), a1.x, sigma); sigma = fmaf(fmaxf(x2.y, 0.f), a1.y, sigma);
      sigma = fmaf(fmaxf(x3.x, 0.f), a1.z, sigma); sigma = fmaf(fmaxf(x3.y, 0.f), a1.w, sigma);
    }
    if (kMask) {
      ch0 = __funnelshift_l(__float_as_uint(x0.x), ch0, 1); ch1 = __funnelshift_l(__float_as_uint(x0.y), ch1, 1);
      ch2 = __funnelshift_l(__float_as_uint(x1.x), ch2, 1); ch3 = __funnelshift_l(__float_as_uint(x1.y), ch3, 1);
      ch0 = __funnelshift_l(__float_as_uint(x2.x), ch0, 1); ch1 = __funnelshift_l(__float_as_uint(x2.y), ch1, 1);
      ch2 = __funnelshift_l(__float_as_uint(x3.x), ch2, 1); ch3 = __funnelshift_l(__float_as_uint(x3.y), ch3, 1);
    }
    constexpr bool kRelu = MODE != 2;
    uint4 o;
    o.x = cvt_bf16x2<kRelu, kF16>(x0.x, x0.y); o.y = cvt_bf16x2<kRelu, kF16>(x1.x, x1.y);
    o.z = cvt_bf16x2<kRelu, kF16>(x2.x, x2.y); o.w = cvt_bf16x2<kRelu, kF16>(x3.x, x3.y);
    *reinterpret_cast<uint4*>(out_row + (((j0 + q) ^ r7) << 4)) = o;
  }
  if (kMask) *mask_word = ch0 | (ch1 << 8) | (ch2 << 16) | (ch3 << 24);
}

// ---- activation store of the training kernels: tile images leave shared memory as per-warp bulk stores ----------------
// A tile image block is [128 rows][128 B] (train_layout.cuh); row r sits at r * 128 and the 128-byte swizzle permutes
// inside the row, so the 32 rows a warp owns (row = (warp & 3) * 32 + lane) are 4 KB contiguous in shared AND in global
// memory.  Once the warp has written its rows of a block, lane 0 hands them to the TMA engine (cp.async.bulk shared ->
// global); the store then streams out in the background.  Because a warp only ever rewrites its OWN rows, the only
// synchronisation is the warp's own cp.async.bulk.wait_group.read before it touches those rows again.
// History (round 1 / 2): the epilogue group's 128 threads copied the finished 64 KB tile out after the hand-off to the
// MMA issuer -- LDS.128 + coalesced STG.128 -- which put epilogue + copy on each slot's critical chain (in-kernel
// timeline: MMA 3300 + epilogue 3700 cycles per stage and slot, tensor pipe 36 %, stores 3.4 TB/s); a single 64 KB bulk
// store issued after the hand-off had the same window and stalled the next epilogue just the same.  Issuing 4 KB pieces
// as the epilogue produces them widens the window by the epilogue itself and frees the threads: training step
// 4.41 -> 4.03-4.13 ms with the forward alone converted (profiles/r02_experiments_not_merged.txt has the dead ends).
// `nblocks` consecutive blocks (stride 16 KB in both address spaces), each 4 KB of this warp's rows.
__device__ __forceinline__ void bulk_store_warp_rows(unsigned char* gdst, uint32_t ssrc, int nblocks) {
  fence_proxy_async_smem();           // this thread's rows -> visible to the async proxy
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    for (int b = 0; b < nblocks; ++b) bulk_s2g(gdst + (size_t)b * 16384, ssrc + (uint32_t)b * 16384u, 4096u);
    bulk_commit();
  }
}
// all bulk stores this warp has issued have finished READING shared memory: its rows may be rewritten
__device__ __forceinline__ void bulk_store_warp_reads_done() {
  if ((threadIdx.x & 31) == 0) bulk_wait_read0();
  __syncwarp();
}
// before the CTA exits: shared memory must outlive the stores
__device__ __forceinline__ void bulk_store_warp_drain() {
  if ((threadIdx.x & 31) == 0) bulk_wait0();
}

// one hidden stage (256 accumulator columns) with the TMEM loads double-buffered.  bulk_g / bulk_s (training, may be
// 0): global / shared address of THIS WARP's 32 rows of K-block 0 of the stage's tile image -- K-block h is stored as soon
// as the warp has written it (bulk_store_warp_rows).
template <int MODE, bool kMask = false, bool kF16 = false>
__device__ __forceinline__ void epi_stage256(uint32_t t_acc, const float4* __restrict__ bias4, unsigned char* a_row_base,
                                             int r7, const float* __restrict__ alpha_w, float& sigma,
                                             uint32_t* mw = nullptr, unsigned char* bulk_g = nullptr, uint32_t bulk_s = 0u) {
  uint32_t va[32], vb[32];
  tmem_ld32(t_acc, va);
  tmem_ld32(t_acc + 32u, vb);
  tmem_ld_wait();
  pin32(va);
  pin32(vb);
#pragma unroll
  for (int h = 0; h < 4; ++h) {   // K-block h of the A operand = columns 64h .. 64h+63
    unsigned char* out_row = a_row_base + h * 16384;
    epi32<MODE, kMask, kF16>(va, bias4 + h * 16, out_row, 0, r7, alpha_w + h * 64, sigma, mw + 2 * h);
    if (h < 3) tmem_ld32(t_acc + (uint32_t)(h * 64 + 64), va);
    epi32<MODE, kMask, kF16>(vb, bias4 + h * 16 + 8, out_row, 4, r7, alpha_w + h * 64 + 32, sigma, mw + 2 * h + 1);
    if (kMask && bulk_g != nullptr) bulk_store_warp_rows(bulk_g + h * 16384, bulk_s + (uint32_t)h * 16384u, 1);
    if (h < 3) {
      tmem_ld32(t_acc + (uint32_t)(h * 64 + 96), vb);
      tmem_ld_wait();
      pin32(va);
      pin32(vb);
    }
  }
}

// write `n8` 16-byte chunks (8 bf16 each) of one 128-byte swizzled row
template <int NCHUNK, bool kF16 = false>
__device__ __forceinline__ void store_row_chunks(uint32_t tile_base, int row, const float* f) {
  uint32_t row_base = tile_base + (uint32_t)row * 128u;
#pragma unroll
  for (int j = 0; j < NCHUNK; ++j) {
    uint32_t addr = row_base + (uint32_t)((j ^ (row & 7)) << 4);
    st_shared_v4(addr, pack_16x2<kF16>(f[j * 8 + 0], f[j * 8 + 1]), pack_16x2<kF16>(f[j * 8 + 2], f[j * 8 + 3]),
                 pack_16x2<kF16>(f[j * 8 + 4], f[j * 8 + 5]), pack_16x2<kF16>(f[j * 8 + 6], f[j * 8 + 7]));
  }
}

// [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)] in groups of 3 (freq.py:23-26).
// One accurate sincosf per coordinate, higher octaves by the double-angle recurrence (x*2^f is
// exact in fp32, so the recurrence is mathematically identical; its error growth, <= 2^9 * 6e-8,
// is far below the bf16 rounding of the operand).
template <int L>
__device__ __forceinline__ void pos_enc_row(const float (&x)[3], float* f) {
  float s[3], c[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    f[k] = x[k];
    sincosf(x[k], &s[k], &c[k]);
  }
#pragma unroll
  for (int l = 0; l < L; ++l) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      f[3 + l * 6 + k] = s[k];
      f[3 + l * 6 + 3 + k] = c[k];
      float s2 = 2.f * s[k] * c[k];
      float c2 = 1.f - 2.f * s[k] * s[k];
      s[k] = s2;
      c[k] = c2;
    }
  }
}

}  // namespace nb
