// Microbenchmark: tcgen05.ld throughput (TMEM -> registers) for 4 or 8 warps per SM.
#include <cstdio>
#include <cstdlib>
#include "../../nerf_rep_for_test_b200/csrc/tc_ptx.cuh"
using namespace nb::ptx;

__device__ __forceinline__ void pin(uint32_t (&r)[32]) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
               "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
  asm volatile("" : "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
               "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]));
}

template <int BATCH>
__global__ void __launch_bounds__(256, 1) k(int iters, int nwarps, unsigned long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc((uint32_t)__cvta_generic_to_shared(&slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem = slot;
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps) {
    uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 256u;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      uint32_t v[BATCH][32];
#pragma unroll
      for (int b = 0; b < BATCH; ++b) tmem_ld32(base + (uint32_t)(((i * BATCH + b) & 7) * 32), v[b]);
      tmem_ld_wait();
#pragma unroll
      for (int b = 0; b < BATCH; ++b) { pin(v[b]); acc ^= v[b][0] ^ v[b][31]; }
    }
    t1 = clock64();
  }
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
  if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  unsigned long long* out; cudaMalloc(&out, 16);
  uint32_t* sink; cudaMalloc(&sink, 4096);
  int iters = 4096;
  for (int nw : {1, 4, 8}) {
    for (int batch : {1, 2, 4}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (batch == 1) k<1><<<148, 256>>>(iters, nw, out, sink);
        if (batch == 2) k<2><<<148, 256>>>(iters, nw, out, sink);
        if (batch == 4) k<4><<<148, 256>>>(iters, nw, out, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      unsigned long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
      double per = (double)h / (iters * batch);
      printf("warps %d batch %d: %.1f cycles per LDTM.x32 per warp -> %.1f B/clk/warp, %.1f B/clk/SM\n", nw, batch, per, 4096.0 / per, 4096.0 / per * nw);
    }
  }
  return 0;
}
