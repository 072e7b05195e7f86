// a3, fp32-ACCURATE tensor-core mode (NERFB200_MODE_FP32_TC): positional encoding fused into the 8x256 NeRF MLP on
// tcgen05 with split-fp16 operands.  Reference arithmetic: fp32 GEMMs, volume_renderer.py:270-284 -> freq.py:23-26 ->
// network.py:49-74.
//
// Every fp32 operand is carried as TWO fp16 numbers, x = x_hi + x_lo with x_hi = fp16(x), x_lo = fp16(x - x_hi)
// (22 significand bits; the products of fp16 numbers are exact in the tensor core's fp32 accumulator), and every
// K = 16 step issues THREE kind::f16 MMAs into the same TMEM accumulator:
//     a_hi * w_hi  +  a_hi * w_lo  +  a_lo * w_hi            (dropped: a_lo * w_lo, <= 2^-22 relative)
// Weights are pre-scaled per stage by a power of two so their residuals stay normal fp16 numbers (mlp_layout.cuh);
// activations are not scaled: |a| < 65504 is required (saturating conversion beyond), and below 2^-3 the residual
// becomes subnormal, i.e. the absolute error per activation is bounded by 2^-25 -- fp32's own rounding at |a| ~ 0.5.
// Measured against a float64 evaluation of the same network (tests/test_gpu_modes.py) the outputs are 2x (rgb) to 6x
// (sigma_raw) as far away as true fp32 arithmetic is (5e-6 against 8e-7 on a 256-term dot product of magnitude ~10): the tensor core aligns and truncates its fp32
// accumulator at each of the 48 MMAs of a 256-deep product, an FFMA chain rounds to nearest.  That is 50x inside the
// 1e-5 tolerance of the mode.  It costs 3 MMAs where single-pass bf16 costs one, and half of what 3xTF32 or six-term
// bf16 splitting would cost.  Heads (alpha_linear, rgb_linear), bias adds, the PE (full-range sincosf per octave,
// no recurrence) and o + d*z are plain fp32 on CUDA cores.  Nine stages: pts_linears.0-7 and views_linears.0 with
// feature_linear folded in at pack time (below).
//
// Structure: the CTA-pair machinery of mlp_bf16_tc2.cu (cta_group::2, M = 256 = 128 rows of each CTA, each CTA
// streams half of every weight chunk, 4-thread-role warp specialisation) with ONE tile slot per CTA: the hi and lo
// images of a [128][256] activation tile are 128 KB of shared memory, so there is no second slot to ping-pong with.
// What hides part of the epilogue instead (it is ~20 % of a stage: 48 MMAs vs one accumulator drain by eight warps) is
// a pipeline BETWEEN consecutive stages of the same tile: the accumulators of even / odd stages live in the two halves
// of TMEM (512 columns), the epilogue drains a stage in two phases -- K-blocks 0,1 of the next A operand first, then
// 2,3 -- and signals each phase separately, so the next stage's MMAs over K-blocks 0,1 run while K-blocks 2,3 are
// still being converted.  Per CTA, 10 warps:
//   warps 0-3 / 4-7  epilogue: row = (warp & 3) * 32 + lane, K-blocks {0, 2} / {1, 3} of the output (64 columns each)
//   warp 8           producer: this CTA's half of the hi and lo image of every weight K-chunk (2-stage ring of 32 KB)
//                    + the stage's fp32 bias block
//   warp 9           rank 0: MMA issuer + TMEM allocator; rank 1: relay of "my half has landed" + allocator
#include <cuda_fp16.h>
#include <stdlib.h>

#include "mlp_tc_common.cuh"

namespace nb {
namespace tcx {
using namespace ptx;

constexpr int kThreads = 320;
constexpr int kRing = 2;
constexpr uint32_t kAPartBytes = 65536;    // hi or lo image of the A tile: 4 K-blocks of [128 rows][64 fp16]
constexpr uint32_t kPePartBytes = 16384;   // hi or lo image of the PE tile
constexpr uint32_t kWPartBytes = 16384;    // hi or lo half-chunk of one ring stage (<= 128 rows x 128 B)
constexpr uint32_t kOffAHi = 0;
constexpr uint32_t kOffALo = kAPartBytes;
constexpr uint32_t kOffPeHi = 2 * kAPartBytes;                  // 131072
constexpr uint32_t kOffPeLo = kOffPeHi + kPePartBytes;          // 147456
constexpr uint32_t kOffW = kOffPeLo + kPePartBytes;             // 163840
constexpr uint32_t kOffBar = kOffW + kRing * 2 * kWPartBytes;   // 229376
constexpr uint32_t kOffBias = kOffBar + 256;
constexpr uint32_t kSmemBytes = kOffBias + 2048;                // 231680 <= 232448
// Executed stages: pts_linears.0-7 and the FUSED tail.  feature_linear has no activation, so -- as in the bf16 inference
// image -- the pack step folds it into views_linears.0 in fp32 (W' = Wv[:, :256] Wf, b' = Wv[:, :256] bf + bv, then split
// into hi / lo like every other weight) and writes the result into the slot of image stage 9; image stage 8 is not
// executed.  10 % fewer MMAs, one accumulator drain less; the change in rounding (h7 (Wf Wv) instead of (h7 Wf) Wv) is
// of the size of an fp32 summation-order change.
constexpr int kExec = 9;
constexpr int kLastStage = kStages - 1;                         // 9: image stage of the (fused) views_linears.0
__device__ __forceinline__ constexpr int exec_stage(int si) { return si < 8 ? si : kLastStage; }

// BAR_AREADY + 0: K-blocks 0,1 of the A operand (and the PE tile) are in place; + 1: K-blocks 2,3 (and the dir PE)
enum { BAR_WFULL = 0, BAR_WEMPTY = 2, BAR_AREADY = 4, BAR_ACCFULL = 6, BAR_BFULL = 7, BAR_BEMPTY = 9, BAR_COUNT = 11 };

// kind::f16 instruction descriptor: D fp32, A/B fp16 (format 0), both K-major, dense, M x N
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// x0, x1 -> packed fp16 pairs (x0 in bits 0-15) of the high parts and of the residuals
__device__ __forceinline__ void split_f16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - h.y), "f"(x0 - h.x));
}

// columns [J0, J0 + NJ) of [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)] in groups of three
// (freq.py:23-26); columns past 3 + 6 L are zero padding.  One full-range sincosf per (octave, coordinate) this column
// range needs.  The octave loop stays rolled (f lives in local memory): this runs once per 128-row tile, off the
// critical path, and thirty inlined sincosf bodies per call site would triple the kernel's code size.
template <int L, int J0, int NJ>
__device__ __forceinline__ void pe_cols(const float (&x)[3], float (&f)[NJ]) {
#pragma unroll
  for (int i = 0; i < NJ; ++i) f[i] = (J0 + i < 3) ? x[(J0 + i) % 3] : 0.f;
#pragma unroll 1
  for (int l = 0; l < L; ++l) {
    const float scale = (float)(1 << l);   // exact scaling by a power of two
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int js = 3 + 6 * l + k - J0, jc = js + 3;
      const bool need_s = js >= 0 && js < NJ, need_c = jc >= 0 && jc < NJ;
      if (need_s || need_c) {
        float sn, cs;
        sincosf(x[k] * scale, &sn, &cs);
        if (need_s) f[js] = sn;
        if (need_c) f[jc] = cs;
      }
    }
  }
}

template <int N>
__device__ __forceinline__ void split_row(const float (&f)[N], uint32_t* hi, uint32_t* lo) {
#pragma unroll
  for (int i = 0; i < N / 2; ++i) split_f16x2(f[2 * i], f[2 * i + 1], hi[i], lo[i]);
}

// 32 accumulator columns of one row: x = acc * 2^-e + bias (fp32), activation, hi/lo split, swizzled 16-byte stores
// into the hi and lo image of the next stage's A operand.  MODE 0: ReLU; 1: ReLU + alpha_linear partial dot on the
// fp32 values (stage 7); 2: linear (feature_linear).
template <int MODE>
__device__ __forceinline__ void epi32x(const uint32_t (&v)[32], const float4* __restrict__ bias4, unsigned char* hi_row,
                                       unsigned char* lo_row, int j0, int r7, float inv,
                                       const float* __restrict__ alpha_w, float& sigma) {
  float4 b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = bias4[i];
  float4 aw[8];
  if (MODE == 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) aw[i] = __ldg(reinterpret_cast<const float4*>(alpha_w) + i);
  }
  const float2 inv2 = make_float2(inv, inv);
  // alpha_linear: four independent partial sums per 32 columns, combined pairwise (a single 128-term FMA chain per
  // thread measured 5e-6 away from a float64 evaluation where blocked fp32 sums -- torch CPU, mlp_fp32_kernel -- are 8e-7)
  float sg0 = 0.f, sg1 = 0.f, sg2 = 0.f, sg3 = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 b0 = b[2 * q], b1 = b[2 * q + 1];
    float2 x0 = __ffma2_rn(make_float2(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1])), inv2, make_float2(b0.x, b0.y));
    float2 x1 = __ffma2_rn(make_float2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3])), inv2, make_float2(b0.z, b0.w));
    float2 x2 = __ffma2_rn(make_float2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5])), inv2, make_float2(b1.x, b1.y));
    float2 x3 = __ffma2_rn(make_float2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7])), inv2, make_float2(b1.z, b1.w));
    if (MODE != 2) {
      x0.x = fmaxf(x0.x, 0.f); x0.y = fmaxf(x0.y, 0.f); x1.x = fmaxf(x1.x, 0.f); x1.y = fmaxf(x1.y, 0.f);
      x2.x = fmaxf(x2.x, 0.f); x2.y = fmaxf(x2.y, 0.f); x3.x = fmaxf(x3.x, 0.f); x3.y = fmaxf(x3.y, 0.f);
    }
    if (MODE == 1) {
      const float4 a0 = aw[2 * q], a1 = aw[2 * q + 1];
      sg0 = fmaf(x0.x, a0.x, sg0); sg0 = fmaf(x0.y, a0.y, sg0);
      sg1 = fmaf(x1.x, a0.z, sg1); sg1 = fmaf(x1.y, a0.w, sg1);
      sg2 = fmaf(x2.x, a1.x, sg2); sg2 = fmaf(x2.y, a1.y, sg2);
      sg3 = fmaf(x3.x, a1.z, sg3); sg3 = fmaf(x3.y, a1.w, sg3);
    }
    uint4 hi, lo;
    split_f16x2(x0.x, x0.y, hi.x, lo.x);
    split_f16x2(x1.x, x1.y, hi.y, lo.y);
    split_f16x2(x2.x, x2.y, hi.z, lo.z);
    split_f16x2(x3.x, x3.y, hi.w, lo.w);
    const int off = ((j0 + q) ^ r7) << 4;
    *reinterpret_cast<uint4*>(hi_row + off) = hi;
    *reinterpret_cast<uint4*>(lo_row + off) = lo;
  }
  if (MODE == 1) sigma += (sg0 + sg1) + (sg2 + sg3);
}

// one K-block (64 accumulator columns at t_blk) of this thread's row -> the hi / lo image rows of that K-block
template <int MODE>
__device__ __forceinline__ void epi_block64x(uint32_t t_blk, const float4* __restrict__ bias4, unsigned char* hi_row,
                                             unsigned char* lo_row, int r7, float inv,
                                             const float* __restrict__ alpha_w, float& sigma) {
  uint32_t va[32], vb[32];
  tmem_ld32(t_blk, va);
  tmem_ld32(t_blk + 32u, vb);
  tmem_ld_wait();
  pin32(va);
  pin32(vb);
  epi32x<MODE>(va, bias4, hi_row, lo_row, 0, r7, inv, alpha_w, sigma);
  epi32x<MODE>(vb, bias4 + 8, hi_row, lo_row, 4, r7, inv, alpha_w + 32, sigma);
}

// which on-chip buffer holds K-chunk c of a stage's input, and how many K = 16 steps it has
__device__ __forceinline__ void chunk_src(int stage, int c, bool& from_pe, int& kblock, int& ksteps) {
  ksteps = 4;
  if (stage == 0) { from_pe = true; kblock = 0; }
  else if (stage == 5) { from_pe = (c == 0); kblock = c - 1; }
  else if (stage == kLastStage) { from_pe = (c == 4); kblock = c; if (c == 4) ksteps = 2; }
  else { from_pe = false; kblock = c; }
}

template <bool kDump>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
mlp_f16x2_tc2_kernel(const unsigned char* __restrict__ packed, const float* __restrict__ rays_o,
                     const float* __restrict__ rays_d, const float* __restrict__ z_vals, long long M, int S,
                     int num_pairs, float* __restrict__ raw, float* __restrict__ stage_dump,
                     const int* __restrict__ row_ids, const int* __restrict__ n_active) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  const uint32_t smem_base = smem_u32(smem_dyn);
  if ((smem_base & 1023u) != 0) __trap();
  const uint32_t bar_base = smem_base + kOffBar;
  const uint32_t tmem_slot = bar_base + 16 * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  auto bar = [&](int i) { return bar_base + (uint32_t)i * 8u; };

  if (threadIdx.x == 0) {
    for (int i = 0; i < kRing; ++i) {
      mbar_init(bar(BAR_WFULL + i), rank == 0 ? 2 : 1);   // leader: own producer + peer relay
      mbar_init(bar(BAR_WEMPTY + i), 1);
    }
    mbar_init(bar(BAR_AREADY + 0), 512);    // every epilogue thread of both CTAs signals the leader's copies
    mbar_init(bar(BAR_AREADY + 1), 512);
    mbar_init(bar(BAR_ACCFULL), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(BAR_BFULL + s), 1);
      mbar_init(bar(BAR_BEMPTY + s), 256);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc_2cta(tmem_slot, 512);   // two accumulators: even / odd stages
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // peer barriers initialised before any remote arrive / multicast commit
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // sparse launch (empty-space skipping): only the rows listed in row_ids[0 .. *n_active) are evaluated
  long long M_eff = M;
  if (row_ids != nullptr) {
    M_eff = *n_active;
    num_pairs = (int)((M_eff + 255) / 256);
  }
  const int my_pairs = num_pairs > cluster_id ? (num_pairs - cluster_id + num_clusters - 1) / num_clusters : 0;
  const float* tail = reinterpret_cast<const float*>(packed + kX2TailOff);

  if (warp < 8) {
    // =========================== epilogue: 2 threads per row (column halves) ===========================
    const int half = warp >> 2;
    const int w4 = warp & 3;
    const int row = w4 * 32 + lane;
    const int r7 = row & 7;
    unsigned char* a_hi_row = smem_dyn + kOffAHi + (uint32_t)row * 128u;   // + K-block * 16384
    unsigned char* a_lo_row = smem_dyn + kOffALo + (uint32_t)row * 128u;
    const uint32_t pe_hi_row = smem_base + kOffPeHi + (uint32_t)row * 128u;
    const uint32_t pe_lo_row = smem_base + kOffPeLo + (uint32_t)row * 128u;
    const uint32_t t_acc = tmem_base + ((uint32_t)(w4 * 32) << 16);
    const uint32_t b_ready0_leader = mapa(bar(BAR_AREADY + 0), 0);
    const uint32_t b_ready1_leader = mapa(bar(BAR_AREADY + 1), 0);
    // partial (rgb, sigma) sums of the upper column half; aliases the first 2 KB of the A tile, which is dead
    // between the last stage's accumulator and the next tile's stage-0 epilogue
    float4* exch = reinterpret_cast<float4*>(smem_dyn + kOffAHi);
    uint32_t full_phase = 0;

    // rays / PE of a tile are computed one tile ahead (during the last stage's MMAs): this thread's 32 of the 64 xyz-PE
    // columns and 16 of the 32 dir-PE columns, already split
    long long m = 0;
    bool valid = false;
    uint32_t pe_hi[16], pe_lo[16], dpe_hi[8], dpe_lo[8];
    auto prepare_tile = [&](int it) {
      const long long tile = 2LL * ((long long)cluster_id + (long long)it * num_clusters) + (long long)rank;
      m = tile * 128 + row;
      valid = m < M_eff;
      if (valid && row_ids != nullptr) m = row_ids[m];   // compacted row -> original (ray, sample) row
      float p[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 0.f};
      if (valid) {
        const long long ray = m / S;
        const float z = z_vals[m];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          d[c] = rays_d[ray * 3 + c];
          p[c] = __fadd_rn(rays_o[ray * 3 + c], __fmul_rn(d[c], z));   // :165, no FMA contraction
        }
      }
      float f[32], g[16];
      if (half == 0) {
        pe_cols<kLx, 0, 32>(p, f);
        pe_cols<kLd, 0, 16>(d, g);
      } else {
        pe_cols<kLx, 32, 32>(p, f);
        pe_cols<kLd, 16, 16>(d, g);
      }
      split_row<32>(f, pe_hi, pe_lo);
      split_row<16>(g, dpe_hi, dpe_lo);
    };
    // it = -1 is a prologue pass that only prepares tile 0, so that prepare_tile has ONE call site (code size)
    for (int it = -1; it < my_pairs; ++it) {
      const long long m_cur = m;
      const bool valid_cur = valid;
      const bool dump_tile = kDump && cluster_id == 0 && it == 0 && rank == 0;
      float sigma = 0.f;
      for (int si = (it < 0 ? kExec - 1 : 0); si < kExec; ++si) {
        const int stage = exec_stage(si);
        if (stage == 0) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {   // xyz PE tile, 16-byte chunks 4*half .. 4*half+3 of this row
            const uint32_t off = (uint32_t)(((half * 4 + q) ^ r7) << 4);
            st_shared_v4(pe_hi_row + off, pe_hi[4 * q], pe_hi[4 * q + 1], pe_hi[4 * q + 2], pe_hi[4 * q + 3]);
            st_shared_v4(pe_lo_row + off, pe_lo[4 * q], pe_lo[4 * q + 1], pe_lo[4 * q + 2], pe_lo[4 * q + 3]);
          }
          fence_proxy_async_smem();
          mbar_arrive_remote(b_ready0_leader);     // stage 0 reads the PE tile only: both phases are ready at once
          mbar_arrive_remote(b_ready1_leader);
        }
        if (stage == kLastStage && it + 1 < my_pairs) prepare_tile(it + 1);   // overlaps the last stage's MMAs
        if (it < 0) break;
        const uint32_t bseq = (uint32_t)it * kExec + (uint32_t)si;
        const uint32_t bbuf = bseq & 1u;
        const float4* bias4 = reinterpret_cast<const float4*>(smem_dyn + kOffBias + bbuf * 1024u);
        const float inv = __ldg(tail + kTailInvScale + stage);
        mbar_wait(bar(BAR_BFULL + bbuf), (bseq >> 1) & 1u, 0x500 + stage);
        mbar_wait(bar(BAR_ACCFULL), full_phase, 0x100 + stage);
        full_phase ^= 1;
        tc_fence_after();
        const uint32_t t_stage = t_acc + (uint32_t)(si & 1) * 256u;   // even / odd executed stages use the two TMEM halves
        if (kDump && dump_tile) {   // diagnostic: fp32 post-activation outputs of rows 0..127 of the whole problem
          const int ncb = stage == kLastStage ? 2 : 4;
          const int c0 = stage == kLastStage ? half * 64 : half * 128;
          for (int cb = 0; cb < ncb; ++cb) {
            uint32_t v[32];
            tmem_ld32(t_stage + (uint32_t)(c0 + cb * 32), v);
            tmem_ld_wait();
            pin32(v);
            for (int i = 0; i < 32; ++i) {
              const int n = c0 + cb * 32 + i;
              float x = fmaxf(fmaf(__uint_as_float(v[i]), inv, tail[kTailBias + stage * 256 + n]), 0.f);
              stage_dump[((size_t)stage * 128 + row) * 256 + n] = x;
            }
          }
        }
        if (stage < kLastStage) {
          // phase 0: K-block `half` (columns 64 half ..), phase 1: K-block 2 + half; each phase is handed to the MMA
          // issuer on its own, so the next stage's MMAs over K-blocks 0,1 overlap the conversion of K-blocks 2,3
#pragma unroll
          for (int ph = 0; ph < 2; ++ph) {
            const int blk = 2 * ph + half;
            const uint32_t tb = t_stage + (uint32_t)blk * 64u;
            const float4* b4 = bias4 + blk * 16;
            unsigned char* hr = a_hi_row + blk * 16384;
            unsigned char* lr = a_lo_row + blk * 16384;
            if (stage == 7) epi_block64x<1>(tb, b4, hr, lr, r7, inv, tail + kTailAlphaW + blk * 64, sigma);
            else epi_block64x<0>(tb, b4, hr, lr, r7, inv, nullptr, sigma);
            if (ph == 1 && stage == 7) {   // dir PE replaces the xyz PE tile (dead after stage 5): chunks 2*half, 2*half+1
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                const uint32_t off = (uint32_t)(((half * 2 + q) ^ r7) << 4);
                st_shared_v4(pe_hi_row + off, dpe_hi[4 * q], dpe_hi[4 * q + 1], dpe_hi[4 * q + 2], dpe_hi[4 * q + 3]);
                st_shared_v4(pe_lo_row + off, dpe_lo[4 * q], dpe_lo[4 * q + 1], dpe_lo[4 * q + 2], dpe_lo[4 * q + 3]);
              }
            }
            tc_fence_before();
            fence_proxy_async_smem();
            mbar_arrive_remote(ph == 0 ? b_ready0_leader : b_ready1_leader);
          }
          mbar_arrive(bar(BAR_BEMPTY + bbuf));
        } else {
          // stage 9: views_linears.0 (128 wide, relu) -> rgb_linear on CUDA cores (network.py:66-69); this thread
          // owns columns 64*half .. 64*half+63
          float r0 = 0.f, r1 = 0.f, r2 = 0.f;
          const float2 inv2 = make_float2(inv, inv);
#pragma unroll
          for (int cb = 0; cb < 2; ++cb) {
            uint32_t v[32];
            tmem_ld32(t_stage + (uint32_t)(half * 64 + cb * 32), v);
            tmem_ld_wait();
            pin32(v);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int n = half * 64 + cb * 32 + q * 4;
              const float4 b4 = bias4[n >> 2];
              const float4 w0 = __ldg(reinterpret_cast<const float4*>(tail + kTailRgbW + n));
              const float4 w1 = __ldg(reinterpret_cast<const float4*>(tail + kTailRgbW + 128 + n));
              const float4 w2 = __ldg(reinterpret_cast<const float4*>(tail + kTailRgbW + 256 + n));
              const float2 xa = __ffma2_rn(make_float2(__uint_as_float(v[q * 4 + 0]), __uint_as_float(v[q * 4 + 1])), inv2, make_float2(b4.x, b4.y));
              const float2 xb = __ffma2_rn(make_float2(__uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3])), inv2, make_float2(b4.z, b4.w));
              const float h0 = fmaxf(xa.x, 0.f), h1 = fmaxf(xa.y, 0.f), h2 = fmaxf(xb.x, 0.f), h3 = fmaxf(xb.y, 0.f);
              r0 = fmaf(h0, w0.x, r0); r0 = fmaf(h1, w0.y, r0); r0 = fmaf(h2, w0.z, r0); r0 = fmaf(h3, w0.w, r0);
              r1 = fmaf(h0, w1.x, r1); r1 = fmaf(h1, w1.y, r1); r1 = fmaf(h2, w1.z, r1); r1 = fmaf(h3, w1.w, r1);
              r2 = fmaf(h0, w2.x, r2); r2 = fmaf(h1, w2.y, r2); r2 = fmaf(h2, w2.z, r2); r2 = fmaf(h3, w2.w, r2);
            }
          }
          tc_fence_before();
          mbar_arrive(bar(BAR_BEMPTY + bbuf));
          if (half == 1) exch[row] = make_float4(r0, r1, r2, sigma);
          named_bar_sync(1, 256);
          if (half == 0 && valid_cur) {
            const float4 q = exch[row];
            float4 o = make_float4((r0 + q.x) + tail[kTailRgbB + 0], (r1 + q.y) + tail[kTailRgbB + 1],
                                   (r2 + q.z) + tail[kTailRgbB + 2], (sigma + q.w) + tail[kTailAlphaB]);
            *reinterpret_cast<float4*>(raw + m_cur * 4) = o;
          }
        }
      }
    }
  } else if (warp == 8) {
    // =========================== producer: bias block + this CTA's half of every weight chunk (hi, lo) ==========
    uint32_t seq = 0, bseq = 0;
    for (int it = 0; it < my_pairs; ++it) {
      for (int si = 0; si < kExec; ++si, ++bseq) {
        const int stage = exec_stage(si);
        {
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
          const uint32_t part = (uint32_t)bf16_chunk_bytes(stage);   // one image (hi or lo), all N rows
          const uint32_t half_bytes = part >> 1;                     // this CTA's N/2 rows
          const unsigned char* src = packed + x2_stage_off(stage) + (size_t)rank * half_bytes;
          const int nch = stage_chunks(stage);
          for (int c = 0; c < nch; ++c, ++seq) {
            const uint32_t pos = seq % kRing, phase = (seq / kRing) & 1u;
            mbar_wait(bar(BAR_WEMPTY + pos), phase ^ 1u, 0x200 + stage);
            mbar_arrive_expect_tx(bar(BAR_WFULL + pos), 2 * half_bytes);
            const uint32_t dst = smem_base + kOffW + pos * (2 * kWPartBytes);
            bulk_g2s(dst, src + (size_t)c * (2 * part), half_bytes, bar(BAR_WFULL + pos));
            bulk_g2s(dst + kWPartBytes, src + (size_t)c * (2 * part) + part, half_bytes, bar(BAR_WFULL + pos));
          }
        }
        __syncwarp();
      }
    }
  } else if (rank == 1) {
    // =========================== relay (peer CTA): my half of chunk seq has landed ===========================
    if (lane == 0) {
      uint32_t seq = 0;
      for (int it = 0; it < my_pairs; ++it)
        for (int si = 0; si < kExec; ++si) {
          const int stage = exec_stage(si);
          const int nch = stage_chunks(stage);
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
    // warp-uniform schedule, one elected lane issues; descriptors = constant high word + precomputed low word + 2*kstep
    uint32_t seq = 0, ready_phase = 0;
    const uint32_t desc_hi = (uint32_t)(umma_desc_sw128(0) >> 32);
    const uint32_t lo_flags = (uint32_t)(umma_desc_sw128(0) & 0xFFFFFFFFu);   // LBO field
    const uint32_t a_hi0 = lo_flags | ((smem_base + kOffAHi) >> 4);     // + kblock*1024 + 2*k
    const uint32_t a_lo0 = lo_flags | ((smem_base + kOffALo) >> 4);
    const uint32_t pe_hi0 = lo_flags | ((smem_base + kOffPeHi) >> 4);
    const uint32_t pe_lo0 = lo_flags | ((smem_base + kOffPeLo) >> 4);
    const uint32_t w0 = lo_flags | ((smem_base + kOffW) >> 4);          // + pos*2048 (+ 1024 for lo) + 2*k
    const uint64_t hi64 = (uint64_t)desc_hi << 32;
    for (int it = 0; it < my_pairs; ++it) {
      for (int si = 0; si < kExec; ++si) {
        const int stage = exec_stage(si);
        const int nch = stage_chunks(stage);
        const uint32_t idesc = umma_idesc_f16(256, stage_n(stage));
        const uint32_t d_tmem = tmem_base + (uint32_t)(si & 1) * 256u;
        // chunks whose A operand is complete after the epilogue's first phase (K-blocks 0,1 + PE tile); the rest
        // (K-blocks 2,3 and the dir PE of the last stage) wait for the second phase
        const int first_part = stage == 0 ? 1 : (stage == 5 ? 3 : 2);
        mbar_wait_cluster(bar(BAR_AREADY + 0), ready_phase, 0x400 + stage);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < nch; ++c, ++seq) {
          if (c == first_part) {
            mbar_wait_cluster(bar(BAR_AREADY + 1), ready_phase, 0x480 + stage);
            tc_fence_after();
          }
          const uint32_t pos = seq % kRing, phase = (seq / kRing) & 1u;
          mbar_wait_cluster(bar(BAR_WFULL + pos), phase, 0x300 + stage);
          tc_fence_after();
          bool from_pe;
          int kblock, ksteps;
          chunk_src(stage, c, from_pe, kblock, ksteps);
          const uint32_t ah = from_pe ? pe_hi0 : (a_hi0 + (uint32_t)kblock * 1024u);
          const uint32_t al = from_pe ? pe_lo0 : (a_lo0 + (uint32_t)kblock * 1024u);
          const uint32_t bh = w0 + pos * ((2 * kWPartBytes) >> 4);
          const uint32_t bl = bh + (kWPartBytes >> 4);
          if (elect_one()) {
#pragma unroll 1
            for (int k = 0; k < ksteps; ++k) {
              const uint32_t ko = 2u * (uint32_t)k;
              umma_bf16_ss_2cta(d_tmem, hi64 | (ah + ko), hi64 | (bh + ko), idesc, (c > 0 || k > 0) ? 1u : 0u);
              umma_bf16_ss_2cta(d_tmem, hi64 | (ah + ko), hi64 | (bl + ko), idesc, 1u);
              umma_bf16_ss_2cta(d_tmem, hi64 | (al + ko), hi64 | (bh + ko), idesc, 1u);
            }
            umma_commit_2cta(bar(BAR_WEMPTY + pos), 3);                      // both CTAs' producers
            if (c == nch - 1) umma_commit_2cta(bar(BAR_ACCFULL), 3);        // both CTAs' epilogue warps
          }
          __syncwarp();
        }
        if (nch <= first_part) {   // stage 0: a single PE chunk -- still consume the second phase of this stage
          mbar_wait_cluster(bar(BAR_AREADY + 1), ready_phase, 0x480 + stage);
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

}  // namespace tcx

int launch_mlp_f16x2(const void* packed, const float* rays_o, const float* rays_d, const float* z_vals, int n_rays,
                     int n_samples, float* raw, float* stage_dump, const int* row_ids, const int* n_active,
                     cudaStream_t st) {
  using namespace tcx;
  int dev = 0, sms = 0;
  NB_CUDA(cudaGetDevice(&dev));
  NB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static bool attr_set[64] = {};   // the opt-in to 226 KB of dynamic shared memory is per device and sticky
  NB_CHECK_ARG(dev >= 0 && dev < 64, "mlp_forward: device ordinal %d out of range", dev);
  if (!attr_set[dev]) {
    NB_CUDA(cudaFuncSetAttribute(mlp_f16x2_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    NB_CUDA(cudaFuncSetAttribute(mlp_f16x2_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    attr_set[dev] = true;
  }
  const long long M = (long long)n_rays * n_samples;
  const long long pairs = (M + 255) / 256;
  const int clusters = (int)(pairs < sms / 2 ? pairs : sms / 2);
  if (stage_dump)
    mlp_f16x2_tc2_kernel<true><<<2 * clusters, kThreads, kSmemBytes, st>>>(
        (const unsigned char*)packed, rays_o, rays_d, z_vals, M, n_samples, (int)pairs, raw, stage_dump, nullptr, nullptr);
  else
    mlp_f16x2_tc2_kernel<false><<<2 * clusters, kThreads, kSmemBytes, st>>>(
        (const unsigned char*)packed, rays_o, rays_d, z_vals, M, n_samples, (int)pairs, raw, nullptr, row_ids, n_active);
  NB_LAUNCH_OK("mlp_f16x2_tc2_kernel");
  return 0;
}

}  // namespace nb
