// Microbenchmark: how fast can ONE CTA per SM drain 64 KB tiles from shared memory to global memory (the activation
// store of the training forward), with every SM doing it at once?  Variants: the kernel's copy loop (T threads, batches
// of B 16-byte chunks per thread), and cp.async.bulk shared -> global with G bulk copies of 64/G KB in flight.
// usage: store_rate ; prints bytes per clock per SM and aggregate TB/s.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

template <int T, int B>
__global__ void __launch_bounds__(T, 1) copy_loop(unsigned char* __restrict__ g, int tiles, size_t stride_tiles, unsigned long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  for (int i = threadIdx.x; i < 65536 / 4; i += T) ((uint32_t*)smem)[i] = i;
  __syncthreads();
  long long t0 = clock64();
  for (int t = 0; t < tiles; ++t) {
    unsigned char* dst = g + ((size_t)t * gridDim.x + blockIdx.x) * 65536;
#pragma unroll 1
    for (int k = 0; k < 65536 / (T * 16 * B); ++k) {
      uint4 v[B];
#pragma unroll
      for (int j = 0; j < B; ++j) v[j] = *reinterpret_cast<const uint4*>(smem + (k * B + j) * T * 16 + threadIdx.x * 16);
#pragma unroll
      for (int j = 0; j < B; ++j) *reinterpret_cast<uint4*>(dst + (k * B + j) * T * 16 + threadIdx.x * 16) = v[j];
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int G>
__global__ void __launch_bounds__(128, 1) bulk_store(unsigned char* __restrict__ g, int tiles, unsigned long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) ((uint32_t*)smem)[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    for (int t = 0; t < tiles; ++t) {
      unsigned char* dst = g + ((size_t)t * gridDim.x + blockIdx.x) * 65536;
#pragma unroll
      for (int c = 0; c < G; ++c)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * (65536 / G)), "r"(s + c * (65536 / G)), "r"(65536 / G) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the tile may be rewritten
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <typename F>
void run(const char* name, F launch, int tiles, int sms) {
  unsigned long long* out;
  cudaMalloc(&out, sms * 8);
  launch(out);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  launch(out);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long h[256];
  cudaMemcpy(h, out, sms * 8, cudaMemcpyDeviceToHost);
  double cyc = 0;
  for (int i = 0; i < sms; ++i) cyc += (double)h[i];
  cyc /= sms;
  printf("%-44s %8.0f cycles / 64 KB tile  %6.1f B/clk/SM  %6.2f TB/s aggregate (%s)\n", name, cyc / tiles, 65536.0 * tiles / cyc,
         65536.0 * tiles * sms / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int tiles = 256;                       // 148 x 256 x 64 KB = 2.4 GB per launch
  unsigned char* g;
  cudaMalloc(&g, (size_t)tiles * sms * 65536);
#define RUN_LOOP(T, B) { cudaFuncSetAttribute(copy_loop<T, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536); \
    run("copy loop " #T " threads, batch " #B, [&](unsigned long long* o) { copy_loop<T, B><<<sms, T, 65536>>>(g, tiles, 0, o); }, tiles, sms); }
  RUN_LOOP(128, 4) RUN_LOOP(128, 8) RUN_LOOP(128, 16) RUN_LOOP(256, 4) RUN_LOOP(256, 8) RUN_LOOP(512, 4) RUN_LOOP(1024, 4)
#define RUN_BULK(G) { cudaFuncSetAttribute(bulk_store<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536); \
    run("cp.async.bulk s2g, " #G " copies per tile", [&](unsigned long long* o) { bulk_store<G><<<sms, 128, 65536>>>(g, tiles, o); }, tiles, sms); }
  RUN_BULK(1) RUN_BULK(4) RUN_BULK(16)
  return 0;
}
