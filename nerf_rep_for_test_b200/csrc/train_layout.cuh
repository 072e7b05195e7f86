// Activation store of the training path (a7): what the forward keeps for the backward pass, what the
// dgrad kernel keeps for the wgrad kernel, and where.
//
// Everything is stored as "tile images": a tile = 128 consecutive MLP rows, a block = the tile's
// [128 rows][64 bf16 columns] cut, laid out exactly like the shared-memory operand of tcgen05.mma with
// SWIZZLE_128B (row r at r*128 B, 16-byte unit u of the row at ((u ^ (r & 7)) << 4)).  One block is
// 16 384 contiguous bytes, so
//   * the producing kernel sends a finished operand tile out as flat bulk stores (cp.async.bulk shared -> global, one
//     per warp and block: a warp's 32 rows are 4 KB contiguous; mlp_tc_common.cuh, bulk_store_warp_rows),
//   * the wgrad kernel streams blocks back with flat bulk copies (no tensor maps), and consumes the SAME
//     image either K-major (K = columns: forward / dgrad) or MN-major (K = rows: wgrad) -- only the
//     descriptor changes.
//
// The training path uses the FUSED TAIL of the inference kernel (mlp_layout.cuh): feature_linear has no
// activation, so views_linears.0(feature_linear(h7)) is one linear map W' = Wv[:, :256] Wf, b' = Wv[:, :256] bf + bv,
// and alpha_linear is one more output column of the same stage.  Forward and backward therefore have nine
// stages (pts_linears.0-7, "8F") and no feature plane; the gradients of Wv, Wf, bf, bv follow from dW', db'
// by the chain rule (wgrad finalize kernel).
// acts  (forward -> backward), per tile kActBlocks blocks:
//   PE(1) | DPE(1) | H0..H7 (4 each: relu(pts_linears.s)) | HV(2: relu(views))
// dacts (dgrad -> wgrad), per tile kDactBlocks blocks:
//   DPRE0..DPRE7 (4 each: dL/d(pre-activation of pts_linears.s)) | D9(4: d_hv(128) | g_raw(4)+0(60) | 0(64))
// masks (forward -> dgrad): relu sign bits, [9][rows_padded][8] uint32; plane s < 8 = pts_linears.s, plane 8 =
//   views (words 0..3).  Word g covers columns 32g..32g+31; column 32g + 4s + k sits at bit 8k + 7 - s and is
//   SET when the pre-activation is negative (gradient blocked).  That bit order lets the dgrad epilogue expand
//   two mask bits into a bf16x2 AND-mask with one SHL + one PRMT (sign-replicate mode).
#pragma once
#include "mlp_layout.cuh"

namespace nb {

constexpr int kTileRows = 128;
constexpr int kBlockBytes = 16384;

constexpr int kActPe = 0;
constexpr int kActDpe = 1;
__host__ __device__ constexpr int act_h(int s) { return 2 + 4 * s; }   // s = 0..7
constexpr int kActHv = 34;
constexpr int kActBlocks = 36;

__host__ __device__ constexpr int dact_pre(int s) { return 4 * s; }    // s = 0..7
constexpr int kDactD9 = 32;
constexpr int kDactBlocks = 36;

constexpr int kMaskPlanes = 9;
constexpr int kMaskWords = 8;

// ---- backward weight image (bf16): the B operands of the dgrad chain, K-major chunks of 64 like the
// forward image (mlp_layout.cuh) but holding W^T:  B[n = input index][k = output index].
//   bstage 0      W'[:, :256]^T = (Wv[:, :256] Wf)^T   N = 256 (h7 idx), K = 128   -> d_h7 (+ g_sigma * alpha_w)
//   bstage 1..7   pts_linears.{7..1}^T                 N = 256, K = 256 (layer 5: the 256 hidden columns only)
// followed by an fp32 tail: rgb_w [3][128], alpha_w [256].
constexpr int kBwdStages = 8;
__host__ __device__ constexpr int bwd_chunks(int b) { return b == 0 ? 2 : 4; }
constexpr int kBwdChunkBytes = 256 * 128;
__host__ __device__ constexpr int bwd_stage_off(int b) { return b == 0 ? 0 : (2 + 4 * (b - 1)) * kBwdChunkBytes; }
constexpr int kBwdTailOff = bwd_stage_off(kBwdStages);         // bytes
constexpr int kBwdTailRgbW = 0;                                // floats
constexpr int kBwdTailAlphaW = 384;
constexpr int kBwdTailFloats = 640;
constexpr int kBwdPackedBytes = kBwdTailOff + kBwdTailFloats * 4;

// ---- wgrad scratch (fp32, zeroed inside mlp_backward): per wgrad stage s = 0..8 a [256 rows (output
// index)][kGradCols] block that the wgrad kernel reduces into with red.global.add:
//   cols   0..255  main product   (s=0: dW0 in cols 0..62; s=8: rows 0..127 = dW'[:, :256], row 131 = d alpha_w)
//   cols 256..383  aux product    (s=5: dW5[:, :63]; s=8: dWv[:, 256:283] in 256..282)
//   cols 384..511  aux2           (s=8: rows 128..130 = dW_rgb [3][128])
//   cols 512..527  ones product   (col 512 = bias gradient; s=8: rows 0..127 = db', 128..130 = d b_rgb, 131 = d b_alpha)
constexpr int kWgradStages = 9;
constexpr int kGradCols = 528;
constexpr int kGradMain = 0, kGradAux = 256, kGradAux2 = 384, kGradOnes = 512;
constexpr size_t kGradScratchFloats = (size_t)kWgradStages * 256 * kGradCols;

}  // namespace nb
