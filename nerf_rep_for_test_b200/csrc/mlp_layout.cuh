// Packed-weight layouts of the two MLP kernels (one NeRF model each; network.py:22-47).
//
// The MLP is treated as a chain of 10 dense "stages" whose inputs are concatenations of
// K-segments that live in on-chip buffers:
//   stage 0      pts_linears.0   in: pe(64)              -> 256 relu
//   stage 1..4   pts_linears.1-4 in: h(256)              -> 256 relu
//   stage 5      pts_linears.5   in: pe(64) | h(256)     -> 256 relu   (skip concat, network.py:57-58)
//   stage 6..7   pts_linears.6-7 in: h(256)              -> 256 relu
//   stage 8      feature_linear  in: h(256)              -> 256 (no activation)
//   stage 9      views_linears.0 in: feat(256) | dpe(32) -> 128 relu
// plus two tiny heads evaluated on CUDA cores: alpha_linear (256->1, on the stage-7 output) and
// rgb_linear (128->3, on the stage-9 output).  pe = 63 PE channels + 1 zero pad, dpe = 27 + 5.
#pragma once
#include "common.cuh"

namespace nb {

constexpr int kStages = 10;
constexpr int kPeK = 64;   // padded 63
constexpr int kDpeK = 32;  // padded 27

__host__ __device__ constexpr int stage_k(int s) { return s == 0 ? 64 : (s == 5 ? 320 : (s == 9 ? 288 : 256)); }
__host__ __device__ constexpr int stage_n(int s) { return s == 9 ? 128 : 256; }
__host__ __device__ constexpr bool stage_relu(int s) { return s != 8; }

// ---- FP32 layout (floats): per stage W^T [K][N] (k-major so a k-chunk is contiguous), then
// biases [10][256], alpha_w[256], alpha_b[4], rgb_w[3][128], rgb_b[4].
__host__ __device__ constexpr int f32_wt_off(int s) {
  int o = 0;
  for (int i = 0; i < s; ++i) o += stage_k(i) * stage_n(i);
  return o;
}
constexpr int kF32BiasOff = f32_wt_off(kStages);
constexpr int kF32AlphaWOff = kF32BiasOff + kStages * 256;
constexpr int kF32AlphaBOff = kF32AlphaWOff + 256;
constexpr int kF32RgbWOff = kF32AlphaBOff + 4;
constexpr int kF32RgbBOff = kF32RgbWOff + 3 * 128;
constexpr int kF32TotalFloats = kF32RgbBOff + 4;

// ---- BF16 layout (bytes): per stage the B operand of tcgen05.mma (N x K, K-major) cut into
// K-chunks of 64 (one 128-byte swizzle row per (n, chunk)); chunk image = N rows x 128 B, already
// XOR-swizzled (16-byte unit index ^= n & 7) so a flat bulk copy lands the canonical
// SWIZZLE_128B layout in shared memory.  Then the fp32 tail: biases [10][256], alpha_w[256],
// alpha_b[4], rgb_w[3][128], rgb_b[4].
__host__ __device__ constexpr int stage_chunks(int s) { return (stage_k(s) + 63) / 64; }  // 1,4,..,5,..,5(288->320)
__host__ __device__ constexpr int bf16_chunk_bytes(int s) { return stage_n(s) * 128; }
__host__ __device__ constexpr int bf16_stage_off(int s) {
  int o = 0;
  for (int i = 0; i < s; ++i) o += stage_chunks(i) * bf16_chunk_bytes(i);
  return o;
}
constexpr int kBf16TailOff = bf16_stage_off(kStages);  // bytes, multiple of 1024
constexpr int kTailBias = 0;
constexpr int kTailAlphaW = kStages * 256;
constexpr int kTailAlphaB = kTailAlphaW + 256;
constexpr int kTailRgbW = kTailAlphaB + 4;
constexpr int kTailRgbB = kTailRgbW + 3 * 128;
constexpr int kTailFloats = kTailRgbB + 4;
constexpr int kBf16TotalBytes = kBf16TailOff + kTailFloats * 4;

// ---- fused inference tail (BF16): feature_linear has no activation, so for inference it is folded
// into views_linears.0 (W' = Wv[:, :256] * Wf, b' = Wv[:, :256] * bf + bv) and alpha_linear rides along
// as output column 128 of the same MMA: stages 8 and 9 become ONE stage "8F" with
//   in: h7(256) | dpe(32)  ->  N = 144 = 128 (views, relu) + 16 (columns 128 + 129 = sigma_raw, alpha_linear's
//   weights split into a bf16 high and low part; rest 0).
// This removes 65 536 of the 593 408 MACs per row and one accumulator drain.  Stages 0..7 are shared
// with the unfused image; the 8F chunks and its fp32 bias row are appended after the unfused image.
constexpr int kFusedN = 144;
constexpr int kFusedChunks = 5;
constexpr int kFusedChunkBytes = kFusedN * 128;                                   // 18432
constexpr int kFusedStageOff = (kBf16TotalBytes + 1023) / 1024 * 1024;
constexpr int kFusedTailOff = kFusedStageOff + kFusedChunks * kFusedChunkBytes;   // fp32 [256]: b'(128), alpha_b, 0...
// fp32 scratch behind the image: the fused-tail product W' = Wv[:, :256] Wf ([128][256]) and b' ([128]), computed ONCE per
// re-pack by fused_tail_product_kernel and read by the forward image's converter and by the backward image's pack
constexpr int kFusedProdOff = (kFusedTailOff + 256 * 4 + 1023) / 1024 * 1024;
constexpr int kFusedProdFloats = 128 * 256 + 128;
constexpr int kBf16PackedBytes = kFusedProdOff + kFusedProdFloats * 4;

// ---- split-fp16 layout (NERFB200_MODE_FP32_TC, bytes): the geometry of the ten stages above, every K-chunk as TWO images of the
// bf16 geometry above -- [hi | lo], each N rows x 128 B, pre-swizzled -- holding w_hi = fp16(w * 2^e_s) and
// w_lo = fp16(w * 2^e_s - w_hi): 22 significand bits per weight.  e_s is a per-stage power of two chosen at pack time
// so that max|w| * 2^e_s lies in [2^13, 2^14) (the residuals of all weights within 2^-16 of the largest stay normal
// fp16 numbers); the epilogue multiplies the accumulator by 2^-e_s (exact).  The fp32 tail is the bf16 tail followed by
// inv_scale[16] = 2^-e_s and fwd_scale[16] = 2^e_s.
__host__ __device__ constexpr int x2_chunk_bytes(int s) { return 2 * bf16_chunk_bytes(s); }
__host__ __device__ constexpr int x2_stage_off(int s) { return 2 * bf16_stage_off(s); }
constexpr int kX2TailOff = 2 * kBf16TailOff;
constexpr int kTailInvScale = kTailFloats;
constexpr int kTailFwdScale = kTailFloats + 16;
constexpr int kX2TailFloats = kTailFloats + 32;
// The slot of stage 9 holds the FUSED tail (feature_linear folded into views_linears.0 in fp32: W' = Wv[:, :256] Wf in
// columns 0..255, Wv[:, 256:283] behind, bias row 9 = b' = Wv[:, :256] bf + bv); the slot of stage 8 is unused.  The
// fp32 product lives in a scratch area behind the tail ([128][256] + [128], fused_tail_product_kernel).
constexpr int kX2ProdOff = (kX2TailOff + kX2TailFloats * 4 + 1023) / 1024 * 1024;
constexpr int kX2PackedBytes = kX2ProdOff + kFusedProdFloats * 4;

// source element of stage s, output n, padded input k (returns false when the slot is padding)
struct SrcRef { int tensor; int col; };  // tensor: 0..7 pts, 8 feature, 9 views
__host__ __device__ inline bool stage_src(int s, int k, SrcRef* r) {
  if (s == 0) { r->tensor = 0; r->col = k; return k < kChX; }
  if (s == 5) {
    r->tensor = 5;
    if (k < 64) { r->col = k; return k < kChX; }
    r->col = kChX + (k - 64); return true;
  }
  if (s == 9) {
    r->tensor = 9;
    if (k < 256) { r->col = k; return true; }
    r->col = k; return (k - 256) < kChD;
  }
  r->tensor = s; r->col = k; return true;  // 1-4, 6-7 pts; 8 feature
}
__host__ __device__ constexpr int tensor_in_features(int t) { return t == 0 ? 63 : (t == 5 ? 319 : (t == 9 ? 283 : 256)); }

}  // namespace nb
