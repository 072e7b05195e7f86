// a7 with the reference's non-detached sampler (volume_renderer.py:181-183), bf16 training path: the gradient of
// the loss with respect to the sample depths z THROUGH THE MLP INPUT,
//     g_pe = dpre0 . W0  +  dpre5 . W5[:, :63]          (pts_linears.0 and the skip concat, network.py:53-59)
//     g_x  = g_pe[0:3] + sum_l 2^l (g_sin_l cos(2^l x) - g_cos_l sin(2^l x))      (freq.py:23-26)
//     g_z  = g_x . d                                                              (x = o + d z, :165)
// read from what nerfb200_mlp_backward's activation-gradient chain already left in its workspace: the bf16 tile
// images DPRE0 and DPRE5 (train_layout.cuh).  6 % of the MACs of one MLP pass and 1 KB of HBM reads per row, so it is
// a plain mma.sync kernel rather than one more tcgen05 stage: one CTA per 128-row tile (persistent), the two 64 KB
// tile images copied verbatim into shared memory (cp.async; the SWIZZLE_128B image is exactly what ldmatrix wants),
// [W0; W5[:, :63]]^T converted to bf16 once per CTA, fp32 accumulation, then one thread per row for the positional-
// encoding backward with the sines and cosines recomputed in fp32 from z.
#include <cuda_bf16.h>

#include "train_layout.cuh"

namespace nb {
namespace bwdin {

constexpr int kThreads = 256;
constexpr int kK = 512;                       // 256 (dpre0) + 256 (dpre5)
constexpr int kBStride = kK * 2 + 16;         // bytes per row of Bt[n][k] (padded: conflict-free ldmatrix)
constexpr int kOffA = 0;                      // 2 x 4 blocks of 16 KB
constexpr int kOffB = 8 * kBlockBytes;        // 131072
constexpr int kSmemBytes = kOffB + 64 * kBStride;   // 197 632
constexpr int kStageStride = 65;              // floats per staged output row (aliases the A region)

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

__global__ void __launch_bounds__(kThreads, 1)
mlp_bwd_input_kernel(const unsigned char* __restrict__ dacts, const float* __restrict__ w0, const float* __restrict__ w5,
                     const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ z_vals,
                     long long M, int S, int n_tiles, float* __restrict__ g_z) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // Bt[n][k] = bf16 of W0[k][n] (k < 256) | W5[k - 256][n] (k >= 256), n < 63; row 63 = 0
  for (int e = tid; e < 64 * kK; e += kThreads) {
    const int k = e >> 6, n = e & 63;          // consecutive threads walk n: coalesced reads of a weight row
    float v = 0.f;
    if (n < kChX) v = k < 256 ? w0[k * kChX + n] : w5[(k - 256) * (kChX + 256) + n];
    *reinterpret_cast<__nv_bfloat16*>(smem + kOffB + n * kBStride + k * 2) = __float2bfloat16_rn(v);
  }
  const uint32_t a_base = smem_addr(smem + kOffA), b_base = smem_addr(smem + kOffB);
  float* stage = reinterpret_cast<float*>(smem + kOffA);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();   // previous tile's staging reads done (and Bt complete on the first pass)
    const unsigned char* src = dacts + (size_t)tile * kDactBlocks * kBlockBytes;
    for (int c = tid; c < 8 * kBlockBytes / 16; c += kThreads) {
      const int blk = c >> 10;                 // 1024 16-byte chunks per block
      const unsigned char* g = src + (size_t)(blk < 4 ? dact_pre(0) + blk : dact_pre(5) + blk - 4) * kBlockBytes + (size_t)(c & 1023) * 16;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a_base + (uint32_t)c * 16u), "l"(g));
    }
    asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[j][q] = 0.f;
    const int mi = lane >> 3, r8 = lane & 7;
    const int a_row = warp * 16 + (mi & 1) * 8 + r8;
#pragma unroll 4
    for (int kk = 0; kk < kK / 16; ++kk) {
      uint32_t a[4];
      const int unit = (kk & 3) * 2 + (mi >> 1);
      ldmatrix_x4(a_base + (uint32_t)(kk >> 2) * kBlockBytes + (uint32_t)a_row * 128u + (uint32_t)((unit ^ (a_row & 7)) << 4), a);
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {         // two 8-column tiles per ldmatrix.x4
        uint32_t b[4];
        const int n = (jp * 2 + (mi >> 1)) * 8 + r8;
        ldmatrix_x4(b_base + (uint32_t)n * kBStride + (uint32_t)(kk * 16 + (mi & 1) * 8) * 2u, b);
        mma_bf16(acc[jp * 2], a, b[0], b[1]);
        mma_bf16(acc[jp * 2 + 1], a, b[2], b[3]);
      }
    }
    __syncthreads();   // every warp is done with the A images: reuse them as the fp32 staging tile
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = warp * 16 + (lane >> 2), c = j * 8 + (lane & 3) * 2;
      stage[r * kStageStride + c] = acc[j][0];
      stage[r * kStageStride + c + 1] = acc[j][1];
      stage[(r + 8) * kStageStride + c] = acc[j][2];
      stage[(r + 8) * kStageStride + c + 1] = acc[j][3];
    }
    __syncthreads();
    if (tid < kTileRows) {
      const long long m = (long long)tile * kTileRows + tid;
      if (m < M) {
        const long long ray = m / S;
        const float z = z_vals[m];
        const float* g = stage + tid * kStageStride;
        float out = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float d = rays_d[ray * 3 + c];
          const float x = __fadd_rn(rays_o[ray * 3 + c], __fmul_rn(d, z));
          float gx = g[c];
          for (int l = 0; l < kLx; ++l) {
            float sn, cs;
            const float f = (float)(1 << l);
            sincosf(x * f, &sn, &cs);
            gx += f * (g[3 + 6 * l + c] * cs - g[3 + 6 * l + 3 + c] * sn);
          }
          out += gx * d;
        }
        g_z[m] = out;
      }
    }
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
