// a3, fp32 parity mode: positional encoding fused into the 8x256 NeRF MLP on CUDA cores.
// Reference: volume_renderer.py:270-284 -> freq.py:23-26 -> network.py:49-74.
//
// This is the 1e-5-relative parity mode (true fp32 FFMA, full-range sinf/cosf), not the
// performance mode (mlp_bf16_tc2.cu).  One CTA owns a tile of 64 sample points and walks the ten
// stages of mlp_layout.cuh with activations resident in shared memory (ONE [64][256] fp32 buffer: a
// stage's outputs overwrite its inputs after the barrier that ends its k loop -- the accumulators live
// in registers) while W^T streams from L2 through a cp.async double buffer in k-chunks of 8.  That is
// 104 KB per CTA, so two CTAs share an SM and one's barriers / weight waits hide behind the other's
// FMAs (the first version ping-ponged two activation buffers with 16-wide chunks: 184 KB, one CTA per
// SM, 8 warps).  Each warp owns
// 8 rows, each lane 8 (or 4) output columns: per k, 2 broadcast LDS.128 of activations + 2
// conflict-free LDS.128 of weights feed 64 FFMAs.  HBM sees only rays, z and the 16 B/row result.
#include "mlp_layout.cuh"

namespace nb {

constexpr int kTileM = 64;
constexpr int kThreads = 256;
constexpr int kKC = 8;

struct SmemF32 {
  float pe[kTileM][kPeK];      // 16 KB
  float dpe[kTileM][kDpeK];    //  8 KB
  float h[kTileM][kW];         // 64 KB
  float wbuf[2][kKC][kW];      // 16 KB
  float sigma[kTileM];
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int N>
__device__ __forceinline__ void load_w_chunk(float* dst, const float* __restrict__ src, int tid) {
  // kKC rows x N floats, contiguous in global
  constexpr int kVec = kKC * N / 4;
  for (int i = tid; i < kVec; i += kThreads) cp_async16(dst + i * 4, src + i * 4);
}

// out[r][n] = act(bias[n] + sum_k in[r][k] * WT[k][n]); input = in0 (K0 cols) ++ in1 (K1 cols)
template <int N>
__device__ __forceinline__ void dense_stage(const float* in0, int ld0, int K0, const float* in1, int ld1,
                                            int K1, const float* __restrict__ WT,
                                            const float* __restrict__ bias, float* out, int ldo, bool relu,
                                            float (*wbuf)[kKC][kW], int tid) {
  constexpr int CPT = N / 32;   // columns per thread: 8 or 4
  constexpr int NV = CPT / 4;   // float4 groups per thread: 2 or 1
  const int warp = tid >> 5, lane = tid & 31;
  const int row0 = warp * 8;
  // accumulators as fp32 pairs: Blackwell's packed FFMA2 (fma.rn.f32x2) does two independent IEEE fmas per issue
  // slot -- same arithmetic, same order per accumulator, half the FMA instructions (the scalar version was issue-bound)
  float2 acc[8][CPT / 2];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < CPT / 2; ++c) acc[r][c] = make_float2(0.f, 0.f);
  const int nchunks = (K0 + K1) / kKC;
  float* wb0 = &wbuf[0][0][0];
  float* wb1 = &wbuf[1][0][0];
  load_w_chunk<N>(wb0, WT, tid);
  cp_async_commit();
  for (int c = 0; c < nchunks; ++c) {
    float* cur = (c & 1) ? wb1 : wb0;
    if (c + 1 < nchunks) {
      load_w_chunk<N>((c & 1) ? wb0 : wb1, WT + (size_t)(c + 1) * kKC * N, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    int k = c * kKC;
    const float* in = (k < K0) ? (in0 + k) : (in1 + (k - K0));
    const int ld = (k < K0) ? ld0 : ld1;
#pragma unroll
    for (int kk = 0; kk < kKC; kk += 4) {
      float4 a[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) a[r] = *reinterpret_cast<const float4*>(in + (size_t)(row0 + r) * ld + kk);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 wv[CPT / 2];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          float4 t = *reinterpret_cast<const float4*>(cur + (kk + i) * N + v * 128 + lane * 4);
          wv[v * 2 + 0] = make_float2(t.x, t.y);
          wv[v * 2 + 1] = make_float2(t.z, t.w);
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float av = i == 0 ? a[r].x : (i == 1 ? a[r].y : (i == 2 ? a[r].z : a[r].w));
          const float2 av2 = make_float2(av, av);
#pragma unroll
          for (int cc = 0; cc < CPT / 2; ++cc) acc[r][cc] = __ffma2_rn(av2, wv[cc], acc[r][cc]);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    float4 b = *reinterpret_cast<const float4*>(bias + v * 128 + lane * 4);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      float4 o;
      o.x = acc[r][v * 2 + 0].x + b.x; o.y = acc[r][v * 2 + 0].y + b.y;
      o.z = acc[r][v * 2 + 1].x + b.z; o.w = acc[r][v * 2 + 1].y + b.w;
      if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      *reinterpret_cast<float4*>(out + (size_t)(row0 + r) * ldo + v * 128 + lane * 4) = o;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads, 2)
mlp_fp32_kernel(const float* __restrict__ packed, const float* __restrict__ rays_o,
                const float* __restrict__ rays_d, const float* __restrict__ z_vals, long long M, int S,
                float* __restrict__ raw) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemF32& sm = *reinterpret_cast<SmemF32*>(smem_raw);
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * kTileM;

  // ---- positional encoding (freq.py:23-26: p_fn(x * 2^f), sin before cos, groups of 3) ----
  {
    int r = tid & 63, part = tid >> 6;  // 4 threads per row
    long long m = m0 + r;
    float p[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 0.f};
    if (m < M) {
      long long ray = m / S;
      float z = z_vals[m];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        d[c] = rays_d[ray * 3 + c];
        p[c] = __fadd_rn(rays_o[ray * 3 + c], __fmul_rn(d[c], z));  // :165, no FMA contraction
      }
    }
    if (part == 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) { sm.pe[r][c] = p[c]; sm.dpe[r][c] = d[c]; }
      sm.pe[r][63] = 0.f;
#pragma unroll
      for (int c = kChD; c < kDpeK; ++c) sm.dpe[r][c] = 0.f;
    }
    // 30 (freq,coord) pairs for xyz, 12 for dirs, split over the 4 threads of the row
    for (int q = part; q < 30; q += 4) {
      int f = q / 3, c = q % 3;
      float a = p[c] * (float)(1 << f);  // exact scaling by a power of two
      sm.pe[r][3 + f * 6 + c] = sinf(a);
      sm.pe[r][3 + f * 6 + 3 + c] = cosf(a);
    }
    for (int q = part; q < 12; q += 4) {
      int f = q / 3, c = q % 3;
      float a = d[c] * (float)(1 << f);
      sm.dpe[r][3 + f * 6 + c] = sinf(a);
      sm.dpe[r][3 + f * 6 + 3 + c] = cosf(a);
    }
  }
  __syncthreads();

  const float* bias = packed + kF32BiasOff;
  float* hA = &sm.h[0][0];   // in-place stages: "A" and "B" are the same buffer
  float* hB = hA;
  float* pe = &sm.pe[0][0];
  float* dpe = &sm.dpe[0][0];
#define NB_WT(s) (packed + f32_wt_off(s))
  dense_stage<256>(pe, kPeK, 64, nullptr, 0, 0, NB_WT(0), bias + 0 * 256, hA, kW, true, sm.wbuf, tid);
  dense_stage<256>(hA, kW, 256, nullptr, 0, 0, NB_WT(1), bias + 1 * 256, hB, kW, true, sm.wbuf, tid);
  dense_stage<256>(hB, kW, 256, nullptr, 0, 0, NB_WT(2), bias + 2 * 256, hA, kW, true, sm.wbuf, tid);
  dense_stage<256>(hA, kW, 256, nullptr, 0, 0, NB_WT(3), bias + 3 * 256, hB, kW, true, sm.wbuf, tid);
  dense_stage<256>(hB, kW, 256, nullptr, 0, 0, NB_WT(4), bias + 4 * 256, hA, kW, true, sm.wbuf, tid);
  dense_stage<256>(pe, kPeK, 64, hA, kW, 256, NB_WT(5), bias + 5 * 256, hB, kW, true, sm.wbuf, tid);
  dense_stage<256>(hB, kW, 256, nullptr, 0, 0, NB_WT(6), bias + 6 * 256, hA, kW, true, sm.wbuf, tid);
  dense_stage<256>(hA, kW, 256, nullptr, 0, 0, NB_WT(7), bias + 7 * 256, hB, kW, true, sm.wbuf, tid);
  // alpha_linear on the stage-7 output (network.py:61), 4 threads per row
  {
    int r = tid >> 2, q = tid & 3;
    const float* aw = packed + kF32AlphaWOff;
    float s = 0.f;
    for (int k = q * 64; k < q * 64 + 64; ++k) s = fmaf(sm.h[r][k], aw[k], s);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (q == 0) sm.sigma[r] = s + packed[kF32AlphaBOff];
  }
  dense_stage<256>(hB, kW, 256, nullptr, 0, 0, NB_WT(8), bias + 8 * 256, hA, kW, false, sm.wbuf, tid);
  dense_stage<128>(hA, kW, 256, dpe, kDpeK, 32, NB_WT(9), bias + 9 * 256, hB, kW, true, sm.wbuf, tid);
#undef NB_WT
  // rgb_linear (network.py:69) + output [rgb_raw, sigma_raw] (network.py:70)
  {
    int r = tid >> 2, q = tid & 3;
    const float* rw = packed + kF32RgbWOff;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int k = q * 32; k < q * 32 + 32; ++k) {
      float h = sm.h[r][k];
      s0 = fmaf(h, rw[k], s0); s1 = fmaf(h, rw[128 + k], s1); s2 = fmaf(h, rw[256 + k], s2);
    }
#pragma unroll
    for (int d = 1; d < 4; d <<= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, d);
      s1 += __shfl_xor_sync(0xffffffffu, s1, d);
      s2 += __shfl_xor_sync(0xffffffffu, s2, d);
    }
    long long m = m0 + r;
    if (q == 0 && m < M) {
      const float* rb = packed + kF32RgbBOff;
      float4 o = make_float4(s0 + rb[0], s1 + rb[1], s2 + rb[2], sm.sigma[r]);
      *reinterpret_cast<float4*>(raw + m * 4) = o;
    }
  }
}

int launch_mlp_fp32(const void* packed, const float* rays_o, const float* rays_d, const float* z_vals,
                    int n_rays, int n_samples, float* raw, cudaStream_t st) {
  int dev = 0;
  NB_CUDA(cudaGetDevice(&dev));
  NB_CHECK_ARG(dev >= 0 && dev < 64, "mlp_forward: device ordinal %d out of range", dev);
  static bool attr_set[64] = {};   // the opt-in to 104 KB of dynamic shared memory is per device and sticky
  if (!attr_set[dev]) {
    NB_CUDA(cudaFuncSetAttribute(mlp_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemF32)));
    attr_set[dev] = true;
  }
  long long M = (long long)n_rays * n_samples;
  int blocks = ceil_div(M, kTileM);
  mlp_fp32_kernel<<<blocks, kThreads, sizeof(SmemF32), st>>>((const float*)packed, rays_o, rays_d, z_vals, M, n_samples, raw);
  NB_LAUNCH_OK("mlp_fp32_kernel");
  return 0;
}

}  // namespace nb
