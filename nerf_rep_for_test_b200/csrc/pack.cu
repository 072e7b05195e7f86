// Weight repacking: nn.Linear fp32 state_dict tensors (network.py:22-47, layout frozen) ->
// the streaming layouts of mlp_layout.cuh.  Re-run after every optimizer step.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "mlp_layout.cuh"
#include "train_layout.cuh"

namespace nb {

struct StageTable {
  int f32_off[kStages];
  int bf16_off[kStages];
};

__device__ __forceinline__ const float* tensor_w(const nerfb200_mlp_weights& w, int t) {
  return t < 8 ? w.pts_w[t] : (t == 8 ? w.feature_w : w.views_w);
}
__device__ __forceinline__ const float* stage_bias(const nerfb200_mlp_weights& w, int s) {
  return s < 8 ? w.pts_b[s] : (s == 8 ? w.feature_b : w.views_b);
}

__device__ __forceinline__ void pack_tail(const nerfb200_mlp_weights& w, float* tail, int i) {
  // i in [0, kTailFloats)
  float v = 0.f;
  if (i < kTailAlphaW) {
    int s = i / 256, n = i % 256;
    if (n < stage_n(s)) v = stage_bias(w, s)[n];
  } else if (i < kTailAlphaB) {
    v = w.alpha_w[i - kTailAlphaW];
  } else if (i < kTailRgbW) {
    v = (i == kTailAlphaB) ? w.alpha_b[0] : 0.f;
  } else if (i < kTailRgbB) {
    v = w.rgb_w[i - kTailRgbW];
  } else {
    int j = i - kTailRgbB;
    v = j < 3 ? w.rgb_b[j] : 0.f;
  }
  tail[i] = v;
}

__global__ void pack_f32_kernel(nerfb200_mlp_weights w, StageTable tab, float* __restrict__ dst) {
  int s = blockIdx.y;
  int K = stage_k(s), N = stage_n(s);
  if (s == kStages) {  // tail
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kTailFloats; i += gridDim.x * blockDim.x)
      pack_tail(w, dst + kF32BiasOff, i);
    return;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K * N; i += gridDim.x * blockDim.x) {
    int k = i / N, n = i % N;
    SrcRef r;
    float v = 0.f;
    if (stage_src(s, k, &r)) v = tensor_w(w, r.tensor)[(size_t)n * tensor_in_features(r.tensor) + r.col];
    dst[tab.f32_off[s] + i] = v;
  }
}

// 16-bit operand element: bf16 (round to nearest) or fp16 (round to nearest, saturating)
__device__ __forceinline__ unsigned short to_16(float v, bool f16) {
  if (f16) {
    v = fminf(fmaxf(v, -65504.f), 65504.f);
    return __half_as_ushort(__float2half_rn(v));
  }
  return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
__device__ __forceinline__ float from_16(unsigned short b, bool f16) {
  return f16 ? __half2float(__ushort_as_half(b)) : __bfloat162float(__ushort_as_bfloat16(b));
}

__global__ void pack_bf16_kernel(nerfb200_mlp_weights w, StageTable tab, unsigned char* __restrict__ dst, bool f16) {
  int s = blockIdx.y;
  if (s == kStages) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kTailFloats; i += gridDim.x * blockDim.x)
      pack_tail(w, reinterpret_cast<float*>(dst + kBf16TailOff), i);
    return;
  }
  int N = stage_n(s), K = stage_k(s), chunks = stage_chunks(s);
  int total = chunks * N * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int c = i / (N * 64);
    int n = (i / 64) % N;
    int kk = i % 64;
    int k = c * 64 + kk;
    SrcRef r;
    float v = 0.f;
    if (k < K && stage_src(s, k, &r)) v = tensor_w(w, r.tensor)[(size_t)n * tensor_in_features(r.tensor) + r.col];
    size_t off = (size_t)tab.bf16_off[s] + (size_t)c * (N * 128) + (size_t)n * 128 +
                 (size_t)(((kk >> 3) ^ (n & 7)) << 4) + (size_t)(kk & 7) * 2;
    *reinterpret_cast<unsigned short*>(dst + off) = to_16(v, f16);
  }
}

// W'[n][k] = sum_j Wv[n][j] Wf[j][k] (and the analogous b'): a 256-term fp32 dot per output element.  Four
// independent accumulators instead of one dependent FMA chain -- the re-pack runs after every optimizer step and the
// single chain made the two tail packs the slowest of the small kernels of the training step (31 us each).
__device__ __forceinline__ float dot256_strided(const float* __restrict__ a, const float* __restrict__ b, int b_stride) {
  float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
#pragma unroll 4
  for (int j = 0; j < 256; j += 4) {
    v0 = fmaf(a[j + 0], b[(size_t)(j + 0) * b_stride], v0);
    v1 = fmaf(a[j + 1], b[(size_t)(j + 1) * b_stride], v1);
    v2 = fmaf(a[j + 2], b[(size_t)(j + 2) * b_stride], v2);
    v3 = fmaf(a[j + 3], b[(size_t)(j + 3) * b_stride], v3);
  }
  return (v0 + v1) + (v2 + v3);
}

// W' = Wv[:, :256] Wf and b' = Wv[:, :256] bf + bv in fp32, once per re-pack.  Every element is the four-accumulator
// sum of dot256_strided (accumulator j % 4, increasing j, (v0+v1)+(v2+v3)) -- bit-identical to the per-element dots of the
// first version (38 + 40 us per model and step) -- with the four chains on four threads.  Block = an 8-row x 32-column
// tile of W' with both operand tiles in shared memory (41 KB read per block, 5 MB in all).  History: 32 blocks x 256
// threads reading Wf through __ldg took 26-39 us (a chain of L2 round trips per thread; ptxas keeps ~5 loads in flight
// whatever the source says); 128 blocks x 1024 threads with 64 loads per thread still 15-22 us (ncu: long-scoreboard
// stall 52 per issue, L2 at 3 % of its throughput: latency, not bandwidth).  The tile fills below are independent
// coalesced loads, eight per thread.
__global__ void __launch_bounds__(1024) fused_tail_product_kernel(nerfb200_mlp_weights w, float* __restrict__ prod) {
  __shared__ float wf[256][32];     // Wf[:, k0 .. k0+32); reused for the four partial sums
  __shared__ float a[8][256];       // Wv[n0 .. n0+8, :256]
  __shared__ float fb[256];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 8;
  const int kk = threadIdx.x & 31, nn = (threadIdx.x >> 5) & 7, q = threadIdx.x >> 8;
  for (int i = threadIdx.x; i < 256 * 32; i += 1024) wf[i >> 5][i & 31] = __ldg(w.feature_w + (size_t)(i >> 5) * 256 + k0 + (i & 31));
  for (int i = threadIdx.x; i < 8 * 256; i += 1024) a[i >> 8][i & 255] = w.views_w[(size_t)(n0 + (i >> 8)) * 283 + (i & 255)];
  if (threadIdx.x < 256) fb[threadIdx.x] = w.feature_b[threadIdx.x];
  __syncthreads();
  float acc = 0.f;
#pragma unroll 16
  for (int j = q; j < 256; j += 4) acc = fmaf(a[nn][j], wf[j][kk], acc);
  __syncthreads();
  float* part = &wf[0][0];          // [4][8][32]
  part[(q * 8 + nn) * 32 + kk] = acc;
  __syncthreads();
  if (q == 0)
    prod[(size_t)(n0 + nn) * 256 + k0 + kk] = (part[nn * 32 + kk] + part[(8 + nn) * 32 + kk]) + (part[(16 + nn) * 32 + kk] + part[(24 + nn) * 32 + kk]);
  if (blockIdx.x == 0 && threadIdx.x < 8)   // b'[n]
    prod[128 * 256 + n0 + threadIdx.x] = w.views_b[n0 + threadIdx.x] + dot256_strided(a[threadIdx.x], fb, 1);
}

// split-fp16 image (NERFB200_MODE_FP32_TC, mlp_layout.cuh): one block per stage finds max|w| of the stage's tensor and
// writes the power-of-two scale 2^e (max|w| * 2^e in [2^13, 2^14)) and its inverse into the tail
__global__ void stage_scale_kernel(nerfb200_mlp_weights w, float* __restrict__ tail, const float* __restrict__ prod) {
  const int s = blockIdx.x;   // stage index = tensor index (0..7 pts_linears, 8 feature_linear (unused), 9 fused views_linears.0)
  const float* W = tensor_w(w, s);
  const int count = stage_n(s) * tensor_in_features(s);
  float mx = 0.f;
  if (s == 9) {   // fused tail: W' = Wv[:, :256] Wf (prod, [128][256]) | Wv[:, 256:283]
    for (int i = threadIdx.x; i < 128 * 256; i += blockDim.x) mx = fmaxf(mx, fabsf(prod[i]));
    for (int i = threadIdx.x; i < 128 * kChD; i += blockDim.x) mx = fmaxf(mx, fabsf(W[(size_t)(i / kChD) * 283 + 256 + i % kChD]));
  } else {
    for (int i = threadIdx.x; i < count; i += blockDim.x) mx = fmaxf(mx, fabsf(W[i]));
  }
  __shared__ float red[32];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
    int e = 0;
    if (mx > 0.f && mx <= 3.0e38f) {
      int ex;
      frexpf(mx, &ex);   // mx = m * 2^ex, m in [0.5, 1)
      e = 14 - ex;
    }
    e = e < -60 ? -60 : (e > 60 ? 60 : e);
    tail[kTailInvScale + s] = ldexpf(1.f, -e);
    tail[kTailFwdScale + s] = ldexpf(1.f, e);
  }
}

__global__ void pack_f16x2_kernel(nerfb200_mlp_weights w, unsigned char* __restrict__ dst) {
  const int s = blockIdx.y;
  float* tail = reinterpret_cast<float*>(dst + kX2TailOff);
  const float* prod = reinterpret_cast<const float*>(dst + kX2ProdOff);   // fused_tail_product_kernel, earlier on the same stream
  if (s == kStages) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kTailFloats; i += gridDim.x * blockDim.x) {
      pack_tail(w, tail, i);
      if (i >= kTailBias + 9 * 256 && i < kTailBias + 9 * 256 + 128) tail[i] = prod[128 * 256 + (i - (kTailBias + 9 * 256))];   // b'
    }
    return;
  }
  if (s == 8) return;                            // feature_linear is folded into the stage-9 slot
  const float scale = tail[kTailFwdScale + s];   // written by stage_scale_kernel, earlier on the same stream
  const int N = stage_n(s), K = stage_k(s), chunks = stage_chunks(s);
  const int total = chunks * N * 64;
  const size_t part = (size_t)N * 128;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i / (N * 64), n = (i / 64) % N, kk = i % 64;
    const int k = c * 64 + kk;
    SrcRef r;
    float v = 0.f;
    if (k < K && stage_src(s, k, &r)) {
      v = (s == 9 && k < 256) ? prod[(size_t)n * 256 + k] : tensor_w(w, r.tensor)[(size_t)n * tensor_in_features(r.tensor) + r.col];
      v *= scale;
    }
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    const size_t off = (size_t)x2_stage_off(s) + (size_t)c * (2 * part) + (size_t)n * 128 +
                       (size_t)(((kk >> 3) ^ (n & 7)) << 4) + (size_t)(kk & 7) * 2;
    *reinterpret_cast<__half*>(dst + off) = hi;
    *reinterpret_cast<__half*>(dst + off + part) = lo;
  }
}

// fused stage 8F (mlp_layout.cuh): one thread per (n, k) of the [144][320] operand + the bias row; the product comes
// from fused_tail_product_kernel (earlier on the same stream)
__global__ void pack_bf16_fused_kernel(nerfb200_mlp_weights w, unsigned char* __restrict__ dst, bool f16) {
  const float* prod = reinterpret_cast<const float*>(dst + kFusedProdOff);
  const int total = kFusedN * kFusedChunks * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total + 256; i += gridDim.x * blockDim.x) {
    if (i >= total) {   // bias row
      int n = i - total;
      float v = 0.f;
      if (n < 128) {
        v = prod[128 * 256 + n];
      } else if (n == 128) {
        v = w.alpha_b[0];
      }
      reinterpret_cast<float*>(dst + kFusedTailOff)[n] = v;
      continue;
    }
    int c = i / (kFusedN * 64);
    int n = (i / 64) % kFusedN;
    int kk = i % 64;
    int k = c * 64 + kk;
    float v = 0.f;
    if (n < 128) {
      if (k < 256) {
        v = prod[(size_t)n * 256 + k];
      } else if (k - 256 < kChD) {
        v = w.views_w[(size_t)n * 283 + k];
      }
    } else if (n == 128 && k < 256) {
      v = w.alpha_w[k];
    } else if (n == 129 && k < 256) {
      // low part of alpha_linear.weight: column 128 holds bf16(w), column 129 the residual, and the epilogue adds
      // the two accumulator columns.  A rounded WEIGHT is a deterministic perturbation of the model (its error is
      // correlated over all samples, unlike activation rounding) and sigma is the output the image is most
      // sensitive to; two of the 16 pad columns of the N = 144 tail make alpha_linear's weights ~16-bit for free.
      const float a = w.alpha_w[k];
      v = a - from_16(to_16(a, f16), f16);
    }
    size_t off = (size_t)kFusedStageOff + (size_t)c * kFusedChunkBytes + (size_t)n * 128 +
                 (size_t)(((kk >> 3) ^ (n & 7)) << 4) + (size_t)(kk & 7) * 2;
    *reinterpret_cast<unsigned short*>(dst + off) = to_16(v, f16);
  }
}

// backward image (train_layout.cuh): W^T chunks for the dgrad chain + fp32 head weights
// prod (may be NULL): the W' product fused_tail_product_kernel left behind the forward image of the SAME weights
__global__ void pack_bf16_bwd_kernel(nerfb200_mlp_weights w, unsigned char* __restrict__ dst, const float* __restrict__ prod) {
  const int b = blockIdx.y;
  if (b == kBwdStages) {
    float* tail = reinterpret_cast<float*>(dst + kBwdTailOff);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kBwdTailFloats; i += gridDim.x * blockDim.x)
      tail[i] = i < kBwdTailAlphaW ? w.rgb_w[i] : w.alpha_w[i - kBwdTailAlphaW];
    return;
  }
  const int total = bwd_chunks(b) * 256 * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int c = i / (256 * 64), n = (i / 64) % 256, kk = i % 64;
    if (b == 0) { n = i % 256; kk = (i / 256) % 64; c = i / (256 * 64); }   // n fastest: coalesced reads of Wf rows
    const int k = c * 64 + kk;   // output index of the layer = K of the dgrad GEMM
    float v = 0.f;
    if (b == 0) {   // fused tail W'[k][n] = sum_j Wv[k][j] Wf[j][n]  (same fp32 arithmetic as pack_bf16_fused_kernel)
      v = prod != nullptr ? prod[(size_t)k * 256 + n] : dot256_strided(w.views_w + (size_t)k * 283, w.feature_w + n, 256);
    } else {
      const int layer = 8 - b;   // 7..1
      v = layer == 5 ? w.pts_w[5][(size_t)k * 319 + kChX + n] : w.pts_w[layer][(size_t)k * 256 + n];
    }
    size_t off = (size_t)bwd_stage_off(b) + (size_t)c * kBwdChunkBytes + (size_t)n * 128 +
                 (size_t)(((kk >> 3) ^ (n & 7)) << 4) + (size_t)(kk & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(dst + off) = __float2bfloat16_rn(v);
  }
}

}  // namespace nb

using namespace nb;

extern "C" size_t nerfb200_packed_weights_bytes(int mode) {
  if (mode == NERFB200_MODE_FP32) return (size_t)kF32TotalFloats * 4;
  if (mode == NERFB200_MODE_BF16 || mode == NERFB200_MODE_FP16) return (size_t)kBf16PackedBytes;
  if (mode == NERFB200_MODE_FP32_TC) return (size_t)kX2PackedBytes;
  return 0;
}

extern "C" int nerfb200_pack_weights(const nerfb200_mlp_weights* w, int mode, void* packed, void* stream) {
  NB_CHECK_ARG(w && packed, "pack_weights: null pointer");
  NB_CHECK_ARG(mode == NERFB200_MODE_FP32 || mode == NERFB200_MODE_BF16 || mode == NERFB200_MODE_FP32_TC ||
                   mode == NERFB200_MODE_FP16, "pack_weights: unknown mode %d", mode);
  NB_CHECK_ARG(((uintptr_t)packed & 1023) == 0, "pack_weights: packed buffer must be 1024-byte aligned");
  for (int i = 0; i < 8; ++i) NB_CHECK_ARG(w->pts_w[i] && w->pts_b[i], "pack_weights: null pts_linears.%d", i);
  NB_CHECK_ARG(w->views_w && w->views_b && w->feature_w && w->feature_b && w->alpha_w && w->alpha_b && w->rgb_w &&
                   w->rgb_b, "pack_weights: null head tensor");
  StageTable tab;
  for (int s = 0; s < kStages; ++s) {
    tab.f32_off[s] = f32_wt_off(s);
    tab.bf16_off[s] = bf16_stage_off(s);
  }
  dim3 grid(64, kStages + 1);
  if (mode == NERFB200_MODE_FP32)
    pack_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*w, tab, (float*)packed);
  else if (mode == NERFB200_MODE_FP32_TC) {
    float* prod = reinterpret_cast<float*>((unsigned char*)packed + kX2ProdOff);
    fused_tail_product_kernel<<<dim3(8, 16), 1024, 0, (cudaStream_t)stream>>>(*w, prod);
    NB_LAUNCH_OK("fused_tail_product_kernel");
    stage_scale_kernel<<<kStages, 256, 0, (cudaStream_t)stream>>>(*w, reinterpret_cast<float*>((unsigned char*)packed + kX2TailOff), prod);
    NB_LAUNCH_OK("stage_scale_kernel");
    pack_f16x2_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*w, (unsigned char*)packed);
  } else {
    const bool f16 = mode == NERFB200_MODE_FP16;
    fused_tail_product_kernel<<<dim3(8, 16), 1024, 0, (cudaStream_t)stream>>>(
        *w, reinterpret_cast<float*>((unsigned char*)packed + kFusedProdOff));
    NB_LAUNCH_OK("fused_tail_product_kernel");
    pack_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*w, tab, (unsigned char*)packed, f16);
    NB_LAUNCH_OK("pack_bf16_kernel");
    pack_bf16_fused_kernel<<<96, 256, 0, (cudaStream_t)stream>>>(*w, (unsigned char*)packed, f16);
  }
  NB_LAUNCH_OK("pack_weights_kernel");
  return 0;
}

extern "C" size_t nerfb200_packed_bwd_bytes(void) { return (size_t)kBwdPackedBytes; }

extern "C" int nerfb200_pack_weights_bwd2(const nerfb200_mlp_weights* w, const void* packed_fwd, void* packed_bwd, void* stream) {
  NB_CHECK_ARG(w && packed_bwd, "pack_weights_bwd: null pointer");
  NB_CHECK_ARG(((uintptr_t)packed_bwd & 1023) == 0 && ((uintptr_t)packed_fwd & 1023) == 0, "pack_weights_bwd: buffers must be 1024-byte aligned");
  for (int i = 1; i < 8; ++i) NB_CHECK_ARG(w->pts_w[i], "pack_weights_bwd: null pts_linears.%d", i);
  NB_CHECK_ARG(w->views_w && w->feature_w && w->alpha_w && w->rgb_w, "pack_weights_bwd: null head tensor");
  const float* prod = packed_fwd ? reinterpret_cast<const float*>((const unsigned char*)packed_fwd + kFusedProdOff) : nullptr;
  pack_bf16_bwd_kernel<<<dim3(64, kBwdStages + 1), 256, 0, (cudaStream_t)stream>>>(*w, (unsigned char*)packed_bwd, prod);
  NB_LAUNCH_OK("pack_bf16_bwd_kernel");
  return 0;
}

extern "C" int nerfb200_pack_weights_bwd(const nerfb200_mlp_weights* w, void* packed_bwd, void* stream) {
  return nerfb200_pack_weights_bwd2(w, nullptr, packed_bwd, stream);
}
