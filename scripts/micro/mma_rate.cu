// Microbenchmark: sustained tcgen05.mma issue rate from one thread, fixed smem operands.
// usage: mma_rate <ctas> ; prints cycles per MMA for several shapes.
#include <cstdio>
#include <cstdlib>
#include "../../nerf_rep_for_test_b200/csrc/tc_ptx.cuh"
using namespace nb::ptx;

template <int CG, int U>
__global__ void __launch_bounds__(128, 1) k(int n_mma, int N, int mode, unsigned long long* out) {
  const int same_d = 0;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint32_t base = smem_u32(smem);
  uint32_t bar = base + 196608, slot = bar + 8;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank = 0;
  if (CG == 2) rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 196608 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;  // small bf16 values
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 16, 1); mbar_init(bar + 32, 1); fence_mbar_init(); }
  if (warp == 0) { if (CG == 2) tmem_alloc_2cta(slot, 512); else tmem_alloc(slot, 512); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (warp == 1 && lane == 0 && rank == 0) {
    uint32_t idesc = nb::ptx::umma_idesc_bf16(CG == 2 ? 256 : 128, N);
    long long t0 = clock64();
    for (int i0 = 0; i0 < n_mma; i0 += U) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u;
        uint32_t a = base + (uint32_t)((i & 3) * 32) + (uint32_t)(((i >> 2) & 3) * 16384);
        uint32_t b = base + 65536 + (uint32_t)((i & 3) * 32) + (uint32_t)(((i >> 2) & 3) * 32768);
        uint32_t d = tmem + (uint32_t)(((i >> 4) & 1) * 256);
        if (CG == 2) umma_bf16_ss_2cta(d, umma_desc_sw128(a), umma_desc_sw128(b), idesc, 1u);
        else umma_bf16_ss(d, umma_desc_sw128(a), umma_desc_sw128(b), idesc, 1u);
      }
      if (mode & 1) { if (CG == 2) umma_commit_2cta(bar + 16, 1); else umma_commit(bar + 16); }
      if (mode & 2) { mbar_wait(bar + 32, 1, 2); }
    }
    long long t1 = clock64();
    if (CG == 2) umma_commit_2cta(bar, 1); else umma_commit(bar);
    mbar_wait(bar, 0, 1);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) { if (CG == 2) tmem_dealloc_2cta(tmem, 512); else tmem_dealloc(tmem, 512); }
}


int main(int argc, char** argv) {
  int ctas = argc > 1 ? atoi(argv[1]) : 148;
  unsigned long long* out; cudaMalloc(&out, 16);
  int smem = 196608 + 64;
  int n = 4096;
  n = argc > 2 ? atoi(argv[2]) : 16384;
  auto run = [&](auto kern, int cg, int U, int mode) {
    unsigned long long h[2];
    for (int rep = 0; rep < 2; ++rep) {
      cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(ctas & ~1); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cg; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      cudaLaunchKernelEx(&cfg, kern, n, 256, mode, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    }
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("cta_group %d unroll %2d mode %d: %.1f cyc/MMA\n", cg, U, mode, (double)h[1] / n);
  };
  for (int mode : {0, 1, 3}) {
    run(k<2, 1>, 2, 1, mode); run(k<2, 2>, 2, 2, mode); run(k<2, 4>, 2, 4, mode); run(k<2, 8>, 2, 8, mode); run(k<2, 16>, 2, 16, mode);
  }
  run(k<1, 4>, 1, 4, 0); run(k<1, 16>, 1, 16, 0);
  return 0;
}
