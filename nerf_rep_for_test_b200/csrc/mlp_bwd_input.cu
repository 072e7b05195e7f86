// a7 with the reference's non-detached sampler (volume_renderer.py:181-183), bf16 training path: the gradient of
// the loss with respect to the sample depths z THROUGH THE MLP INPUT,
//     g_pe = dpre0 . W0  +  dpre5 . W5[:, :63]          (pts_linears.0 and the skip concat, network.py:53-59)
//     g_x  = g_pe[0:3] + sum_l 2^l (g_sin_l cos(2^l x) - g_cos_l sin(2^l x))      (freq.py:23-26)
//     g_z  = g_x . d                                                              (x = o + d z, :165)
// read from what nerfb200_mlp_backward's activation-gradient chain already left in its workspace: the bf16 tile
// images DPRE0 and DPRE5 (train_layout.cuh).  6 % of the MACs of one MLP pass and 1 KB of HBM reads per row, so it is
// a plain mma.sync kernel rather than one more tcgen05 stage: persistent CTAs walk 128-row tiles in two 64-row halves,
// the tile images copied verbatim into a double-buffered shared-memory stage (cp.async; the SWIZZLE_128B image is exactly
// what ldmatrix wants), [W0; W5[:, :63]]^T converted to bf16 once per CTA, fp32 accumulation, then the positional-
// encoding backward with the sines and cosines recomputed in fp32 from z.
#include <cuda_bf16.h>

#include "train_layout.cuh"

namespace nb {
namespace bwdin {

constexpr int kThreads = 256;
constexpr int kK = 512;                       // 256 (dpre0) + 256 (dpre5)
constexpr int kHalfRows = 64;                 // rows per pipeline step (half a 128-row tile)
constexpr int kHalfBlockBytes = kHalfRows * 128;              // 8 KB: 64 rows of one [128][64] block
constexpr int kABufBytes = 8 * kHalfBlockBytes;               // 64 KB: 2 planes x 4 blocks
constexpr int kBStride = kK * 2 + 16;         // bytes per row of Bt[n][k] (padded: conflict-free ldmatrix)
constexpr int kOffA = 0;                      // two buffers (double-buffered cp.async)
constexpr int kOffB = 2 * kABufBytes;         // 131072
constexpr int kOffStage = kOffB + 64 * kBStride;              // 197 632: fp32 [64][65] partial / final g_pe tile
constexpr int kStageStride = 65;
constexpr int kOffPart = kOffStage + kHalfRows * kStageStride * 4;   // 214 272: [64][3] per-coordinate contributions
constexpr int kSmemBytes = kOffPart + kHalfRows * 3 * 4 + 256;       // 215 296 <= 232 448

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Round-2 revision: 64-row half tiles, double-buffered -- the copy of the next half tile overlaps the MMAs of this one
// (the first version loaded a whole 128 KB tile, then computed: 401 us per fine launch, 2.0 TB/s); warps split the
// contraction in halves (dpre0 / dpre5) so every B fragment feeds the same number of MMAs with half the ldmatrix
// traffic per warp; the positional-encoding backward runs on 192 threads (row x coordinate) with two sincosf per
// coordinate and the double-angle recurrence in between instead of 30 sincosf on 128 threads.
__global__ void __launch_bounds__(kThreads, 1)
mlp_bwd_input_kernel(const unsigned char* __restrict__ dacts, const float* __restrict__ w0, const float* __restrict__ w5,
                     const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ z_vals,
                     long long M, int S, int n_tiles, float* __restrict__ g_z) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_base = smem_addr(smem + kOffA), b_base = smem_addr(smem + kOffB);
  float* stage = reinterpret_cast<float*>(smem + kOffStage);
  float* part = reinterpret_cast<float*>(smem + kOffPart);
  const int my_tiles = n_tiles > (int)blockIdx.x ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int n_steps = 2 * my_tiles;            // half tiles of this CTA, in order
  // cp.async of half tile `step` into buffer step & 1: 8 blocks (dpre0 0..3, dpre5 0..3) x 8 KB
  auto prefetch = [&](int step) {
    const long long tile = (long long)blockIdx.x + (long long)(step >> 1) * gridDim.x;
    const unsigned char* src = dacts + (size_t)tile * kDactBlocks * kBlockBytes + (size_t)(step & 1) * kHalfBlockBytes;
    const uint32_t dst = a_base + (uint32_t)(step & 1) * kABufBytes;
    for (int c = tid; c < kABufBytes / 16; c += kThreads) {
      const int blk = c >> 9;                  // 512 16-byte chunks per 8 KB half block
      const unsigned char* g = src + (size_t)(blk < 4 ? dact_pre(0) + blk : dact_pre(5) + blk - 4) * kBlockBytes + (size_t)(c & 511) * 16;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)c * 16u), "l"(g));
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (n_steps > 0) prefetch(0);
  // Bt[n][k] = bf16 of W0[k][n] (k < 256) | W5[k - 256][n] (k >= 256), n < 63; row 63 = 0
  for (int e = tid; e < 64 * kK; e += kThreads) {
    const int k = e >> 6, n = e & 63;          // consecutive threads walk n: coalesced reads of a weight row
    float v = 0.f;
    if (n < kChX) v = k < 256 ? w0[k * kChX + n] : w5[(k - 256) * (kChX + 256) + n];
    *reinterpret_cast<__nv_bfloat16*>(smem + kOffB + n * kBStride + k * 2) = __float2bfloat16_rn(v);
  }
  const int mt = warp & 3, kh = warp >> 2;     // 16-row tile of the half, half of the contraction
  const int mi = lane >> 3, r8 = lane & 7;
  const int a_row = mt * 16 + (mi & 1) * 8 + r8;
  for (int step = 0; step < n_steps; ++step) {
    if (step + 1 < n_steps) {
      prefetch(step + 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();                           // this half tile (and, on the first pass, Bt) is in shared memory
    const uint32_t a_buf = a_base + (uint32_t)(step & 1) * kABufBytes + (uint32_t)kh * 4u * kHalfBlockBytes;
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[j][q] = 0.f;
#pragma unroll 4
    for (int kk = 0; kk < 16; ++kk) {          // 256 of the 512 k per warp
      uint32_t a[4];
      const int unit = (kk & 3) * 2 + (mi >> 1);
      ldmatrix_x4(a_buf + (uint32_t)(kk >> 2) * kHalfBlockBytes + (uint32_t)a_row * 128u + (uint32_t)((unit ^ (a_row & 7)) << 4), a);
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t b[4];
        const int n = (jp * 2 + (mi >> 1)) * 8 + r8;
        ldmatrix_x4(b_base + (uint32_t)n * kBStride + (uint32_t)(kh * 256 + kk * 16 + (mi & 1) * 8) * 2u, b);
        mma_bf16(acc[jp * 2], a, b[0], b[1]);
        mma_bf16(acc[jp * 2 + 1], a, b[2], b[3]);
      }
    }
    // reduce the two contraction halves through the staging tile
    const int r = mt * 16 + (lane >> 2);
    if (kh == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = j * 8 + (lane & 3) * 2;
        stage[r * kStageStride + c] = acc[j][0];
        stage[r * kStageStride + c + 1] = acc[j][1];
        stage[(r + 8) * kStageStride + c] = acc[j][2];
        stage[(r + 8) * kStageStride + c + 1] = acc[j][3];
      }
    }
    __syncthreads();
    if (kh == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = j * 8 + (lane & 3) * 2;
        stage[r * kStageStride + c] += acc[j][0];
        stage[r * kStageStride + c + 1] += acc[j][1];
        stage[(r + 8) * kStageStride + c] += acc[j][2];
        stage[(r + 8) * kStageStride + c + 1] += acc[j][3];
      }
    }
    __syncthreads();
    // positional-encoding backward: thread = (row, coordinate)
    const long long tile = (long long)blockIdx.x + (long long)(step >> 1) * gridDim.x;
    const long long m0 = tile * kTileRows + (long long)(step & 1) * kHalfRows;
    if (tid < kHalfRows * 3) {
      const int row = tid / 3, c = tid - row * 3;
      const long long m = m0 + row;
      float contrib = 0.f;
      if (m < M) {
        const long long ray = m / S;
        const float d = rays_d[ray * 3 + c];
        const float x = __fadd_rn(rays_o[ray * 3 + c], __fmul_rn(d, z_vals[m]));
        const float* g = stage + row * kStageStride;
        float sn = 0.f, cs = 1.f;
        float gx = g[c], f = 1.f;
#pragma unroll
        for (int l = 0; l < kLx; ++l) {
          // angle doubling (2^l x is exact in fp32), restarted from an accurate sincosf every five octaves: the
          // recurrence's error doubles per step and the high octaves carry the weight 2^l
          if (l % 5 == 0) sincosf(x * f, &sn, &cs);
          gx += f * (g[3 + 6 * l + c] * cs - g[3 + 6 * l + 3 + c] * sn);
          const float s2 = 2.f * sn * cs, c2 = 1.f - 2.f * sn * sn;
          sn = s2; cs = c2; f *= 2.f;
        }
        contrib = gx * d;
      }
      part[row * 3 + c] = contrib;
    }
    __syncthreads();
    if (tid < kHalfRows) {
      const long long m = m0 + tid;
      if (m < M) g_z[m] = (part[tid * 3] + part[tid * 3 + 1]) + part[tid * 3 + 2];
    }
    // the next iteration's first __syncthreads orders these reads before the staging tile / this A buffer are reused
  }
}

}  // namespace bwdin
}  // namespace nb

using namespace nb;

extern "C" int nerfb200_mlp_backward_input(const nerfb200_mlp_weights* weights, const void* workspace, const float* rays_o,
                                           const float* rays_d, const float* z_vals, int n_rays, int n_samples, float* g_z,
                                           void* stream) {
  NB_CHECK_ARG(n_rays >= 0 && n_samples >= 1, "mlp_backward_input: bad sizes");
  if (n_rays == 0) return 0;
  NB_CHECK_ARG(weights && weights->pts_w[0] && weights->pts_w[5], "mlp_backward_input: null pts_linears.0 / .5 weight");
  NB_CHECK_ARG(workspace && rays_o && rays_d && z_vals && g_z, "mlp_backward_input: null pointer");
  NB_CHECK_ARG(((uintptr_t)workspace & 127) == 0, "mlp_backward_input: workspace must be 128-byte aligned");
  const long long M = (long long)n_rays * n_samples;
  const long long tiles = (M + kTileRows - 1) / kTileRows;
  NB_CHECK_ARG(tiles < (1LL << 31), "mlp_backward_input: too many rows");
  int dev = 0, sms = 0;
  NB_CUDA(cudaGetDevice(&dev));
  NB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static bool attr_set[64] = {};
  NB_CHECK_ARG(dev >= 0 && dev < 64, "mlp_backward_input: device ordinal %d out of range", dev);
  if (!attr_set[dev]) {
    NB_CUDA(cudaFuncSetAttribute(bwdin::mlp_bwd_input_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bwdin::kSmemBytes));
    attr_set[dev] = true;
  }
  const int grid = (int)(tiles < sms ? tiles : sms);
  bwdin::mlp_bwd_input_kernel<<<grid, bwdin::kThreads, bwdin::kSmemBytes, (cudaStream_t)stream>>>(
      (const unsigned char*)workspace, weights->pts_w[0], weights->pts_w[5], rays_o, rays_d, z_vals, M, n_samples, (int)tiles, g_z);
  NB_LAUNCH_OK("mlp_bwd_input_kernel");
  return 0;
}
