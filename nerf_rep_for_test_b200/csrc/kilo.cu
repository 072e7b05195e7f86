// a9: the KiloNeRF-style path of BASELINE configs[4] -- occupancy-grid ray marching, thousands of
// 32-wide micro-MLPs, integration with early ray termination -- rebuilt for B200.
// Reference semantics: cuda/generate_inputs.cu:11-35 (ray directions), :60-126 (march),
// cuda/network_eval.cu:24-254 (micro-MLP), cuda/integrate.cu:9-57 (integrate + ERT), :84-97 (background).
// The reference's path never runs; what is kept is its arithmetic and its data contract
// (query index = ray * max_depth + depth, occupancy grid of network ids, 6212 floats per network packed
// [bias | W in-major] per layer, multi-pass state (depth_indices, transmittance, active mask)).
//
// What is different (B200-first):
//  * no host round trips: the pass loop is a fixed sequence of launches on one stream; every kernel reads
//    the device-side counters and returns at once when the previous pass left nothing to do;
//  * grouping samples by network is a device-side counting sort (shared-memory histograms, one scan block,
//    one scatter) instead of thrust::sort_by_key + gather/scatter (cuda/reorder.cu:12-48);
//  * the micro-MLP kernel is persistent over (network, 256-sample chunk) work items, so a few crowded
//    networks do not serialise on one block (the reference maps block i <-> network i); the five layers run
//    on the tensor cores with split-fp16 operands (fp32-accurate, see eval_tc_kernel) and the hidden
//    activations never leave registers; results are written straight back to the ray-major slot (no
//    scatter pass);
//  * cos/sin by one accurate sincosf per coordinate + the double-angle recurrence instead of 20 fast-math
//    intrinsics on arguments up to 512 rad.
// 12 160 FLOP per sample against 16 B in / 16 B out: the path is compute-bound, and by design evaluates ~10x
// fewer samples than the dense path.  Round 1 ran the micro-MLP on CUDA cores (packed FFMA2, two samples per
// thread: 3.0 ms for the bench frame's 8.06 M samples, 44 % of the FFMA peak); the tensor-core kernel below takes
// 1.25 ms (profiles/r02_kilo_ab.txt).
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace nb {
namespace kilo {

// micro-MLP geometry (network_eval.cu:48-52): hidden width 32, 63 position / 27 direction embedding channels
constexpr int kParamSize = 6212;
constexpr int kOffL0 = 0;                                   // bias[32] | W[63][32]
constexpr int kOffL1 = kOffL0 + 32 + 63 * 32;               // bias[32] | W[32][32]
constexpr int kOffL2 = kOffL1 + 32 + 32 * 32;               // bias[33] | W[32][33]
constexpr int kOffL3 = kOffL2 + 33 + 32 * 33;               // bias[32] | W[59][32]
constexpr int kOffL4 = kOffL3 + 32 + 59 * 32;               // bias[3]  | W[32][3]
static_assert(kOffL4 + 3 + 32 * 3 == kParamSize, "micro-MLP parameter layout");

constexpr int kChunk = 256;        // samples per work item of the micro-MLP kernel (8 warps x 32)

// counters[] (device int32): 0 = queries of this pass, 1 = work items of this pass, 2 = rays still active
// after this pass, 3 = rays active before this pass
enum { C_QUERIES = 0, C_ITEMS = 1, C_ACTIVE_NEXT = 2, C_ACTIVE = 3, C_COUNT = 8 };

__global__ void rays_d_kernel(nerfb200_kilo_camera cam, float* __restrict__ dirs, float* __restrict__ dists, float dbp) {
  const int n = cam.H * cam.W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int x = i % cam.W, y = i / cam.W;
    const float v[3] = {__fdiv_rn((float)x - cam.cx, cam.fx), -__fdiv_rn((float)y - cam.cy, cam.fy), -1.f};
    float d[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int k = 0; k < 3; ++k) d[k] = __fadd_rn(d[k], __fmul_rn(v[j], cam.c2w[k * 3 + j]));
    dirs[i * 3 + 0] = d[0]; dirs[i * 3 + 1] = d[1]; dirs[i * 3 + 2] = d[2];
    if (dists) dists[i] = __fmul_rn(dbp, sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2]))));
  }
}

// generate_inputs.cu:60-126, one thread per ray.  Unfilled slots get assigned = -1 (and query = -1).
__global__ void march_kernel(nerfb200_kilo_grid g, const float* __restrict__ origin3, const float* __restrict__ dirs,
                             const int16_t* __restrict__ grid, int32_t* __restrict__ query, int16_t* __restrict__ assigned,
                             uint8_t* __restrict__ active, int32_t* __restrict__ depth_idx, int n_rays, float dbp, int spp,
                             int max_depth, float min_distance, int initial, int32_t* __restrict__ counters) {
  if (!initial && counters && counters[C_ACTIVE] == 0) return;
  const float o[3] = {origin3[0], origin3[1], origin3[2]};
  float voxel[3];
  int stride[3] = {g.res[1] * g.res[2], g.res[2], 1};
#pragma unroll
  for (int c = 0; c < 3; ++c) voxel[c] = __fdiv_rn(__fsub_rn(g.gmax[c], g.gmin[c]), (float)g.res[c]);
  float inv_voxel[3], lo_in[3], hi_in[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    inv_voxel[c] = __fdiv_rn(1.f, voxel[c]);
    lo_in[c] = __fadd_rn(g.gmin[c], 0.001f);
    hi_in[c] = __fsub_rn(g.gmax[c], 0.001f);
  }
  int emitted = 0;
  for (int ray = blockIdx.x * blockDim.x + threadIdx.x; ray < n_rays; ray += gridDim.x * blockDim.x) {
    int out = 0;
    const bool act = initial ? true : active[ray] != 0;
    int32_t* q = query + (size_t)ray * spp;
    int16_t* a = assigned + (size_t)ray * spp;
    if (act) {
      const float d[3] = {dirs[ray * 3], dirs[ray * 3 + 1], dirs[ray * 3 + 2]};
      int depth = initial ? 0 : depth_idx[ray];
      float dist = __fadd_rn(min_distance, __fmul_rn((float)depth, dbp));
      while (depth < max_depth && out < spp) {
        float p[3];
        bool inside = true;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          p[c] = __fadd_rn(o[c], __fmul_rn(dist, d[c]));
          inside = inside && (lo_in[c] < p[c]) && (p[c] < hi_in[c]);
        }
        int net = -1;
        if (inside) {
          // cell index = (int)((p - gmin) / voxel), the reference's IEEE division (generate_inputs.cu:100-104).  The
          // product with the reciprocal is within 3 ulp of that quotient, so the two truncate to the same integer
          // unless the product lies within 3 ulp of one; only then is the division itself evaluated (bit-exact
          // either way; the three divisions were ~40 % of the loop's instructions).
          int flat = 0;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float x = __fsub_rn(p[c], g.gmin[c]);
            float qv = __fmul_rn(x, inv_voxel[c]);
            const float fr = qv - truncf(qv), tol = qv * 4e-7f;
            if (fr < tol || fr > 1.f - tol) qv = __fdiv_rn(x, voxel[c]);
            flat += (int)qv * stride[c];
          }
          net = (int)grid[flat];
        }
        if (net != -1) {
          a[out] = (int16_t)net;
          q[out] = ray * max_depth + depth;
          ++out;
        }
        ++depth;
        dist = __fadd_rn(dist, dbp);
      }
      if (out < spp) active[ray] = 0;
      else { active[ray] = 1; depth_idx[ray] = depth; }
    }
    emitted += out;
    for (; out < spp; ++out) { a[out] = -1; q[out] = -1; }
  }
  if (counters) {   // total queries of this pass (warp-aggregated)
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) emitted += __shfl_xor_sync(0xffffffffu, emitted, s);
    if ((threadIdx.x & 31) == 0 && emitted) atomicAdd(&counters[C_QUERIES], emitted);
  }
}

// ---- counting sort by network id ---------------------------------------------------------------
__global__ void hist_kernel(const int16_t* __restrict__ assigned, size_t n_slots, int num_networks, int32_t* __restrict__ count,
                            const int32_t* __restrict__ counters) {
  if (counters[C_QUERIES] == 0) return;
  extern __shared__ int32_t sh[];
  for (int i = threadIdx.x; i < num_networks; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const size_t per = (n_slots + gridDim.x - 1) / gridDim.x;
  const size_t lo = (size_t)blockIdx.x * per, hi = lo + per < n_slots ? lo + per : n_slots;
  for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const int net = assigned[i];
    if (net >= 0) atomicAdd(&sh[net], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < num_networks; i += blockDim.x)
    if (sh[i]) atomicAdd(&count[i], sh[i]);
}

// one block: exclusive scans of count[] (sample offsets) and of ceil(count/kChunk) (work-item offsets)
__global__ void scan_kernel(const int32_t* __restrict__ count, int num_networks, int32_t* __restrict__ start,
                            int32_t* __restrict__ cursor, int32_t* __restrict__ item_start, int32_t* __restrict__ counters) {
  if (counters[C_QUERIES] == 0) { if (threadIdx.x == 0) counters[C_ITEMS] = 0; return; }
  __shared__ int32_t s_tot[2][32];
  __shared__ int32_t s_carry[2];
  if (threadIdx.x == 0) s_carry[0] = s_carry[1] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int base = 0; base < num_networks; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int c = i < num_networks ? count[i] : 0;
    int v0 = c, v1 = (c + kChunk - 1) / kChunk;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const int t0 = __shfl_up_sync(0xffffffffu, v0, s), t1 = __shfl_up_sync(0xffffffffu, v1, s);
      if (lane >= s) { v0 += t0; v1 += t1; }
    }
    if (lane == 31) { s_tot[0][warp] = v0; s_tot[1][warp] = v1; }
    __syncthreads();
    int w0 = 0, w1 = 0;
    for (int w = 0; w < warp; ++w) { w0 += s_tot[0][w]; w1 += s_tot[1][w]; }
    const int c0 = s_carry[0], c1 = s_carry[1];
    if (i < num_networks) {
      start[i] = c0 + w0 + v0 - c;
      cursor[i] = c0 + w0 + v0 - c;
      item_start[i] = c1 + w1 + v1 - (c + kChunk - 1) / kChunk;
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) {
      int t0 = 0, t1 = 0;
      for (int w = 0; w < nw; ++w) { t0 += s_tot[0][w]; t1 += s_tot[1][w]; }
      s_carry[0] = c0 + t0;
      s_carry[1] = c1 + t1;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    item_start[num_networks] = s_carry[1];
    start[num_networks] = s_carry[0];
    counters[C_ITEMS] = s_carry[1];
  }
}

// sorted[start[net] + k] = slot, for every filled slot (order inside a network is irrelevant: results are
// written back per slot).  Two shared-memory passes per block: local histogram -> one global cursor bump
// per non-empty bin -> local ranks.
__global__ void scatter_kernel(const int16_t* __restrict__ assigned, size_t n_slots, int num_networks, int32_t* __restrict__ cursor,
                               int32_t* __restrict__ sorted, const int32_t* __restrict__ counters) {
  if (counters[C_QUERIES] == 0) return;
  extern __shared__ int32_t sh[];
  int32_t* base = sh;                 // [num_networks]
  for (int i = threadIdx.x; i < num_networks; i += blockDim.x) base[i] = 0;
  __syncthreads();
  const size_t per = (n_slots + gridDim.x - 1) / gridDim.x;
  const size_t lo = (size_t)blockIdx.x * per, hi = lo + per < n_slots ? lo + per : n_slots;
  for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const int net = assigned[i];
    if (net >= 0) atomicAdd(&base[net], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < num_networks; i += blockDim.x) {
    const int c = base[i];
    base[i] = c ? atomicAdd(&cursor[i], c) : 0;
  }
  __syncthreads();
  for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const int net = assigned[i];
    if (net >= 0) sorted[atomicAdd(&base[net], 1)] = (int32_t)i;
  }
}

// ---- micro-MLP ---------------------------------------------------------------------------------
// [x, cos(2^0 x) .. cos(2^(L-1) x), sin(2^0 x) .. sin(2^(L-1) x)]  (network_eval.cu:127-140)
template <int L>
__device__ __forceinline__ void fourier(float x, float* e) {
  float s, c;
  sincosf(x, &s, &c);
  e[0] = x;
#pragma unroll
  for (int l = 0; l < L; ++l) {
    e[1 + l] = c;
    e[1 + L + l] = s;
    const float s2 = 2.f * s * c, c2 = 1.f - 2.f * s * s;
    s = s2; c = c2;
  }
}

struct EvalCam {
  float c2w[9];
  float origin[3];
  float cx, cy, fx, fy;
  int W, max_depth;
  float min_distance, dbp;
};

// ---- micro-MLP on the tensor cores (split-fp16 operands, fp32 accumulate) -------------------------------------------
// The five layers (network_eval.cu:127-254) as warp-level mma.sync.m16n8k16 products: M = 32 samples per warp (two 16-row
// tiles, lane l prepares sample l), N = the layer's outputs in tiles of 8, K = its inputs in steps of 16.  fp32 accuracy
// on 16-bit tensor-core operands as in mlp_f16x2_tc2.cu: every operand is carried as two fp16 numbers x = hi + lo,
// hi = fp16(x), lo = fp16(x - hi) (22 significand bits; products of fp16 numbers are exact in the fp32 accumulator) and
// every K step issues       a_lo * b_hi  +  a_hi * b_lo  +  a_hi * b_hi          (dropped: a_lo * b_lo, 2^-22 relative).
// (First version: 3xTF32 on m16n8k8 -- the legacy tf32 HMMA runs at ~410 MAC/clk/SM, so three of them per K = 8 only
// matched the FFMA2 peak: 2.05 ms against 3.0 ms for the CUDA-core kernel of round 1.  fp16 halves the instruction count per MAC.)
// Range: a network's weights are scaled by one power of two S (max |w| S < 2^14, found when the network is staged) so
// their residuals stay normal fp16 numbers; the accumulator is un-scaled in fp32 (h = acc / S + bias).  Activations are
// not scaled: |h| < 65504 is required (saturating conversion beyond), below 2^-3 the residual is subnormal, i.e. the
// absolute error of an activation is bounded by 2^-25.
// The accumulator fragment of a layer IS the A fragment of the next one (thread (g, t) = (lane / 4, lane % 4) holds
// columns 2t, 2t+1 of output tile i for rows g, g+8 = K slots 2t, 2t+1 (tile 2j) and 2t+8, 2t+9 (tile 2j+1) of K step j),
// so hidden activations never leave registers.  Weights are staged in shared memory already split, packed and in
// fragment order: one conflict-free LDS.128 (b0.hi, b1.hi, b0.lo, b1.lo) per (K step, N tile) and lane, no conversion
// work in the inner loop.  Inputs that do not come from a previous layer (the 63 position / 27 direction embedding
// channels, computed per lane by sincosf + the double-angle recurrence) go through a
// feature-major staging tile per warp ([64 features][36 floats], conflict-free both ways).  density (layer-2 output 0)
// rides in a fifth N tile so that the 32 feature outputs stay aligned with the K steps of layer 3.
constexpr int kTcThreads = 256;
constexpr int kTcWarps = kTcThreads / 32;
// fragment tiles (512 B = 32 lanes x uint4 each): layer 0: 4 K steps x 4 N tiles, 1: 2 x 4, 2: 2 x 5 (tiles 0-3 = feature
// 1..32, tile 4 column 0 = density), 3: 4 x 4 (K steps 0-1 feature, 2-3 direction embedding), 4: 2 x 1
constexpr int kT0 = 0, kT1 = kT0 + 16, kT2 = kT1 + 8, kT3 = kT2 + 10, kT4 = kT3 + 16, kTiles = kT4 + 2;   // 52
constexpr int kFragFloats = kTiles * 128;       // 6656 32-bit words
constexpr int kTb0 = kFragFloats, kTb1 = kTb0 + 32, kTb2 = kTb1 + 32, kTb3 = kTb2 + 40, kTb4 = kTb3 + 32;
constexpr int kTcWFloats = kTb4 + 8;            // 6800
constexpr int kStageStride = 36;
constexpr int kStageFloats = 64 * kStageStride;
constexpr int kTcMisc = 8 + kTcWarps;           // dom[6], scale, 1/scale, one partial maximum per warp
constexpr size_t kTcSmemBytes = (size_t)(kTcWFloats + kTcWarps * kStageFloats + kTcMisc) * 4;   // 64 128 B

// x0, x1 -> packed fp16 pairs (x0 in bits 0-15) of the high parts and of the residuals
__device__ __forceinline__ void split_h2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - h.y), "f"(x0 - h.x));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// one K step (16 inputs) of a layer for both row tiles: acc[m][i] += a[m] * Wfrag[i].  a[m] = fp32 values of
// {row g: slots 2t, 2t+1 | row g+8: same | row g: slots 2t+8, 2t+9 | row g+8: same}; wf = this K step's NT fragment tiles
template <int NT>
__device__ __forceinline__ void mma_kstep(float (&acc)[2][NT][4], const float (&a)[2][8], const uint4* __restrict__ wf, int lane) {
  uint32_t ahi[2][4], alo[2][4];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int r = 0; r < 4; ++r) split_h2(a[m][2 * r], a[m][2 * r + 1], ahi[m][r], alo[m][r]);
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const uint4 b = wf[i * 32 + lane];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      mma_f16(acc[m][i], alo[m], b.x, b.y);
      mma_f16(acc[m][i], ahi[m], b.z, b.w);
      mma_f16(acc[m][i], ahi[m], b.x, b.y);
    }
  }
}
template <int NT>
__device__ __forceinline__ void zero_acc(float (&acc)[2][NT][4]) {
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int i = 0; i < NT; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[m][i][e] = 0.f;
}
// h = acc / S + bias (optionally through relu), in place
template <int NT, bool kRelu>
__device__ __forceinline__ void finish_layer(float (&acc)[2][NT][4], const float* __restrict__ bias, float inv_s, int t) {
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const float2 v = *reinterpret_cast<const float2*>(bias + 8 * i + 2 * t);
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float h = fmaf(acc[m][i][e], inv_s, (e & 1) ? v.y : v.x);
        acc[m][i][e] = kRelu ? fmaxf(h, 0.f) : h;
      }
  }
}
// finished output tiles 2j, 2j+1 of the previous layer -> A operand of K step j
template <int NT>
__device__ __forceinline__ void frag_c_to_a(float (&a)[2][8], const float (&c)[2][NT][4], int j) {
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    a[m][0] = c[m][2 * j][0];     a[m][1] = c[m][2 * j][1];
    a[m][2] = c[m][2 * j][2];     a[m][3] = c[m][2 * j][3];
    a[m][4] = c[m][2 * j + 1][0]; a[m][5] = c[m][2 * j + 1][1];
    a[m][6] = c[m][2 * j + 1][2]; a[m][7] = c[m][2 * j + 1][3];
  }
}
// staged (feature-major) inputs -> A operand of staged K step js
__device__ __forceinline__ void frag_stage_to_a(float (&a)[2][8], const float* __restrict__ stg, int js, int g, int t) {
  const float* r = stg + (16 * js + 2 * t) * kStageStride + g;
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    a[m][0] = r[m * 16];                          a[m][1] = r[kStageStride + m * 16];
    a[m][2] = r[m * 16 + 8];                      a[m][3] = r[kStageStride + m * 16 + 8];
    a[m][4] = r[8 * kStageStride + m * 16];       a[m][5] = r[9 * kStageStride + m * 16];
    a[m][6] = r[8 * kStageStride + m * 16 + 8];   a[m][7] = r[9 * kStageStride + m * 16 + 8];
  }
}
// packed parameters of one network -> split, scaled, fragment-ordered shared-memory image of one layer.
// index_of(in, i, g): parameter index of W[in][output g of tile i], or -1 for padding
template <int KS, int NT, typename F>
__device__ __forceinline__ void fill_frags(uint4* __restrict__ dst, const float* __restrict__ src, float scale, F index_of) {
  for (int idx = threadIdx.x; idx < KS * NT * 32; idx += kTcThreads) {
    const int lane = idx & 31, tile = idx >> 5;
    const int i = tile % NT, j = tile / NT, g = lane >> 2, t = lane & 3;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int p = index_of(16 * j + 2 * t + (e & 1) + (e >> 1) * 8, i, g);
      v[e] = p >= 0 ? __ldg(src + p) * scale : 0.f;
    }
    uint4 o;
    split_h2(v[0], v[1], o.x, o.z);
    split_h2(v[2], v[3], o.y, o.w);
    dst[idx] = o;
  }
}

__global__ void __launch_bounds__(kTcThreads, 2)
eval_tc_kernel(const int32_t* __restrict__ query, const int32_t* __restrict__ sorted, const int32_t* __restrict__ start,
               const int32_t* __restrict__ item_start, int num_networks, const float* __restrict__ params,
               const float* __restrict__ domain_mins, const float* __restrict__ domain_maxs, EvalCam cam,
               float4* __restrict__ rgb_sigma, const int32_t* __restrict__ counters) {
  extern __shared__ __align__(16) float tc_smem[];
  float* w = tc_smem;
  const uint4* wf = reinterpret_cast<const uint4*>(tc_smem);
  float* misc = tc_smem + kTcWFloats + kTcWarps * kStageFloats;   // dom[6] | S | 1/S | partial maxima
  const float* dom = misc;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  float* stg = tc_smem + kTcWFloats + warp * kStageFloats;
  const int n_items = counters[C_ITEMS];
  int cached_net = -1;
  const int per_block = (n_items + (int)gridDim.x - 1) / (int)gridDim.x;
  const int item_end = min(n_items, ((int)blockIdx.x + 1) * per_block);
  // network of an item = last n with item_start[n] <= item: one binary search for the block's first item, then a
  // forward walk (consecutive items mostly share the network; the search was 12 dependent L2 round trips per item)
  int net = 0, net_items_end = 0;
  for (int item = (int)blockIdx.x * per_block; item < item_end; ++item) {
    if (cached_net < 0) {
      int lo = 0, hi = num_networks;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (item_start[mid] <= item) lo = mid; else hi = mid;
      }
      net = lo;
      net_items_end = item_start[net + 1];
    }
    while (item >= net_items_end) net_items_end = item_start[++net + 1];
    if (net != cached_net) {
      __syncthreads();
      const float* src = params + (size_t)net * kParamSize;
      float mx = 0.f;
      for (int i = threadIdx.x; i < kParamSize; i += kTcThreads) mx = fmaxf(mx, fabsf(__ldg(src + i)));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (lane == 0) misc[8 + warp] = mx;
      __syncthreads();
      mx = 0.f;
#pragma unroll
      for (int k = 0; k < kTcWarps; ++k) mx = fmaxf(mx, misc[8 + k]);
      int ex = 0;
      if (mx > 0.f && mx < 3.0e38f) frexpf(mx, &ex);       // mx < 2^ex
      const float scale = ldexpf(1.f, min(max(14 - ex, -100), 100));
      uint4* wq = reinterpret_cast<uint4*>(w);
      fill_frags<4, 4>(wq + kT0 * 32, src, scale, [](int in, int i, int g) { return in < 63 ? kOffL0 + 32 + in * 32 + 8 * i + g : -1; });
      fill_frags<2, 4>(wq + kT1 * 32, src, scale, [](int in, int i, int g) { return kOffL1 + 32 + in * 32 + 8 * i + g; });
      fill_frags<2, 5>(wq + kT2 * 32, src, scale, [](int in, int i, int g) { return i < 4 ? kOffL2 + 33 + in * 33 + 1 + 8 * i + g : (g == 0 ? kOffL2 + 33 + in * 33 : -1); });
      fill_frags<4, 4>(wq + kT3 * 32, src, scale, [](int in, int i, int g) { return in < 59 ? kOffL3 + 32 + in * 32 + 8 * i + g : -1; });
      fill_frags<2, 1>(wq + kT4 * 32, src, scale, [](int in, int i, int g) { return g < 3 ? kOffL4 + 3 + in * 3 + g : -1; });
      for (int c = threadIdx.x; c < kTcWFloats - kTb0; c += kTcThreads) {
        int p = -1;
        if (c < 32) p = kOffL0 + c;
        else if (c < 64) p = kOffL1 + c - 32;
        else if (c < 104) { const int k = c - 64; p = k < 32 ? kOffL2 + 1 + k : (k == 32 ? kOffL2 : -1); }
        else if (c < 136) p = kOffL3 + c - 104;
        else if (c < 139) p = kOffL4 + c - 136;
        w[kTb0 + c] = p >= 0 ? __ldg(src + p) : 0.f;
      }
      if (threadIdx.x < 3) misc[threadIdx.x] = domain_mins[net * 3 + threadIdx.x];
      else if (threadIdx.x < 6) misc[threadIdx.x] = domain_maxs[net * 3 + threadIdx.x - 3];
      else if (threadIdx.x == 6) misc[6] = scale;
      else if (threadIdx.x == 7) misc[7] = 1.f / scale;
      __syncthreads();
      cached_net = net;
    }
    const float inv_s = misc[7];
    const int first = start[net] + (item - item_start[net]) * kChunk;
    const int end = start[net + 1];
#pragma unroll 1
    for (int sub = warp; sub < kChunk / 32; sub += kTcWarps) {
      if (first + sub * 32 >= end) break;   // warp-uniform
      const int idx = first + sub * 32 + lane;
      const int slot = idx < end ? sorted[idx] : -1;
      float pe[3], dir[3];
      {
        int q = slot >= 0 ? query[slot] : 0;
        const int depth = q % cam.max_depth;
        q /= cam.max_depth;
        const int x = q % cam.W, y = q / cam.W;
        const float v[3] = {__fdiv_rn((float)x - cam.cx, cam.fx), -__fdiv_rn((float)y - cam.cy, cam.fy), -1.f};
        float d[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
          for (int k = 0; k < 3; ++k) d[k] = __fadd_rn(d[k], __fmul_rn(v[j], cam.c2w[k * 3 + j]));
        const float dist = __fadd_rn(cam.min_distance, __fmul_rn((float)depth, cam.dbp));
        const float norm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float p = __fadd_rn(cam.origin[k], __fmul_rn(dist, d[k]));
          pe[k] = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, __fsub_rn(p, dom[k])), __fsub_rn(dom[3 + k], dom[k])), 1.f);
          dir[k] = __fdiv_rn(d[k], norm);
        }
      }
      // position embedding of sample `lane` -> staging tile (feature-major)
      __syncwarp();
#pragma unroll 1
      for (int j = 0; j < 3; ++j) {
        float e[21];
        fourier<10>(pe[j], e);
#pragma unroll
        for (int k = 0; k < 21; ++k) stg[(j * 21 + k) * kStageStride + lane] = e[k];
      }
      stg[63 * kStageStride + lane] = 0.f;
      __syncwarp();
      // layer 0: 63 -> 32
      float h0[2][4][4];
      zero_acc<4>(h0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a[2][8];
        frag_stage_to_a(a, stg, j, g, t);
        mma_kstep<4>(h0, a, wf + (kT0 + j * 4) * 32, lane);
      }
      finish_layer<4, true>(h0, w + kTb0, inv_s, t);
      // direction embedding replaces the (consumed) position embedding in the staging tile
      __syncwarp();
#pragma unroll 1
      for (int j = 0; j < 3; ++j) {
        float e[9];
        fourier<4>(dir[j], e);
#pragma unroll
        for (int k = 0; k < 9; ++k) stg[(j * 9 + k) * kStageStride + lane] = e[k];
      }
#pragma unroll
      for (int k = 27; k < 32; ++k) stg[k * kStageStride + lane] = 0.f;
      __syncwarp();
      // layer 1: relu(h0) 32 -> 32
      float h1[2][4][4];
      zero_acc<4>(h1);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float a[2][8];
        frag_c_to_a<4>(a, h0, j);
        mma_kstep<4>(h1, a, wf + (kT1 + j * 4) * 32, lane);
      }
      finish_layer<4, true>(h1, w + kTb1, inv_s, t);
      // layer 2: relu(h1) 32 -> feature(32) | density (no activation on the feature)
      float h2[2][5][4];
      zero_acc<5>(h2);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float a[2][8];
        frag_c_to_a<4>(a, h1, j);
        mma_kstep<5>(h2, a, wf + (kT2 + j * 5) * 32, lane);
      }
      finish_layer<5, false>(h2, w + kTb2, inv_s, t);
      // layer 3: feature(32) | direction embedding(27) -> 32
      float (&h3)[2][4][4] = h0;
      zero_acc<4>(h3);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float a[2][8];
        frag_c_to_a<5>(a, h2, j);
        mma_kstep<4>(h3, a, wf + (kT3 + j * 4) * 32, lane);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float a[2][8];
        frag_stage_to_a(a, stg, j, g, t);
        mma_kstep<4>(h3, a, wf + (kT3 + (2 + j) * 4) * 32, lane);
      }
      finish_layer<4, true>(h3, w + kTb3, inv_s, t);
      // layer 4: relu(h3) 32 -> 3
      float o[2][1][4];
      zero_acc<1>(o);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float a[2][8];
        frag_c_to_a<4>(a, h3, j);
        mma_kstep<1>(o, a, wf + (kT4 + j) * 32, lane);
      }
      finish_layer<1, false>(o, w + kTb4, inv_s, t);
      // outputs: thread t = 0 of a quad holds (r, g) of rows g / g+8 and the density (column 0 of tile 4 of layer 2),
      // t = 1 holds b.  Lane t of the quad finishes and stores component t of the row's float4 (one sigmoid per lane and
      // row instead of three on a quarter of the lanes; the quad's four 4-byte stores fall into one 16-byte segment)
      const int q0 = lane & ~3;
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
          const float c_r = o[m][0][2 * hrow], c_g = o[m][0][2 * hrow + 1], c_s = h2[m][4][2 * hrow];
          const float x_g = __shfl_sync(0xffffffffu, c_g, q0);
          const float x_b = __shfl_sync(0xffffffffu, c_r, q0 + 1);
          const float x_s = __shfl_sync(0xffffffffu, c_s, q0);
          const int s_row = __shfl_sync(0xffffffffu, slot, m * 16 + g + 8 * hrow);
          const float x = t == 0 ? c_r : (t == 1 ? x_g : (t == 2 ? x_b : x_s));
          const float y = t == 3 ? fmaxf(x, 0.f) : 1.f / (1.f + expf(-x));
          if (s_row >= 0) reinterpret_cast<float*>(rgb_sigma + s_row)[t] = y;
        }
    }
  }
}

// integrate.cu:9-57, one thread per ray over the filled slots of this pass
__global__ void integrate_kernel(const float4* __restrict__ rgb_sigma, const int16_t* __restrict__ assigned,
                                 const float* __restrict__ dists, float* __restrict__ rgb_map, float* __restrict__ acc_map,
                                 float* __restrict__ T, uint8_t* __restrict__ active, int n_rays, int spp, float thr, int initial,
                                 int32_t* __restrict__ counters) {
  if (!initial && counters && counters[C_ACTIVE] == 0) return;
  int still = 0;
  for (int ray = blockIdx.x * blockDim.x + threadIdx.x; ray < n_rays; ray += gridDim.x * blockDim.x) {
    float t = initial ? 1.f : T[ray];
    const bool act = t > thr;
    float r = 0.f, g = 0.f, b = 0.f, acc = 0.f;
    if (act) {
      if (!initial) { r = rgb_map[ray * 3]; g = rgb_map[ray * 3 + 1]; b = rgb_map[ray * 3 + 2]; acc = acc_map[ray]; }
      const float dist = dists[ray];
      const size_t s0 = (size_t)ray * spp;
      for (int s = 0; s < spp && assigned[s0 + s] >= 0; ++s) {
        const float4 v = rgb_sigma[s0 + s];
        const float alpha = __fsub_rn(1.f, expf(-__fmul_rn(v.w, dist)));
        const float wgt = __fmul_rn(alpha, t);
        t = __fmul_rn(t, __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f));
        r = __fadd_rn(r, __fmul_rn(v.x, wgt));
        g = __fadd_rn(g, __fmul_rn(v.y, wgt));
        b = __fadd_rn(b, __fmul_rn(v.z, wgt));
        acc = __fadd_rn(acc, wgt);
      }
      T[ray] = t;
      if (t <= thr) active[ray] = 0;
    } else if (initial) {
      T[ray] = t;
    }
    if (act || initial) {
      rgb_map[ray * 3] = r; rgb_map[ray * 3 + 1] = g; rgb_map[ray * 3 + 2] = b;
      acc_map[ray] = acc;
    }
    still += active[ray] ? 1 : 0;
  }
  if (counters) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) still += __shfl_xor_sync(0xffffffffu, still, s);
    if ((threadIdx.x & 31) == 0 && still) atomicAdd(&counters[C_ACTIVE_NEXT], still);
  }
}

// integrate.cu:84-97
__global__ void background_kernel(float* __restrict__ rgb_map, const float* __restrict__ acc_map, int n, float br, float bg, float bb) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float t = __fsub_rn(1.f, acc_map[i]);
    rgb_map[i * 3] = __fadd_rn(rgb_map[i * 3], __fmul_rn(br, t));
    rgb_map[i * 3 + 1] = __fadd_rn(rgb_map[i * 3 + 1], __fmul_rn(bg, t));
    rgb_map[i * 3 + 2] = __fadd_rn(rgb_map[i * 3 + 2], __fmul_rn(bb, t));
  }
}

// between passes: stats += queries; ACTIVE <- ACTIVE_NEXT; clear per-pass counters and the histogram
__global__ void next_pass_kernel(int32_t* __restrict__ counters, int32_t* __restrict__ count, int num_networks, int64_t* __restrict__ stats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < num_networks) count[i] = 0;
  if (i == 0) {
    if (stats) { stats[0] += counters[C_QUERIES]; stats[1] += counters[C_QUERIES] > 0 ? 1 : 0; }
    counters[C_ACTIVE] = counters[C_ACTIVE_NEXT];
    counters[C_ACTIVE_NEXT] = 0;
    counters[C_QUERIES] = 0;
    counters[C_ITEMS] = 0;
  }
}

struct Work {
  float* dirs; float* dists; float* T; float4* rgb_sigma; int32_t* query; int32_t* sorted; int32_t* depth_idx;
  int32_t* count; int32_t* start; int32_t* cursor; int32_t* item_start; int32_t* counters; int16_t* assigned; uint8_t* active;
  float* origin; size_t bytes;
};
static size_t al(size_t x) { return (x + 255) & ~(size_t)255; }
static Work carve(void* base, int n_rays, int spp, int num_networks) {
  Work w;
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? (char*)base + off : nullptr; off += al(bytes); return p; };
  const size_t slots = (size_t)n_rays * spp;
  w.dirs = (float*)take((size_t)n_rays * 12);
  w.dists = (float*)take((size_t)n_rays * 4);
  w.T = (float*)take((size_t)n_rays * 4);
  w.rgb_sigma = (float4*)take(slots * 16);
  w.query = (int32_t*)take(slots * 4);
  w.sorted = (int32_t*)take(slots * 4);
  w.depth_idx = (int32_t*)take((size_t)n_rays * 4);
  w.count = (int32_t*)take((size_t)(num_networks + 1) * 4);
  w.start = (int32_t*)take((size_t)(num_networks + 1) * 4);
  w.cursor = (int32_t*)take((size_t)(num_networks + 1) * 4);
  w.item_start = (int32_t*)take((size_t)(num_networks + 1) * 4);
  w.counters = (int32_t*)take(C_COUNT * 4);
  w.assigned = (int16_t*)take(slots * 2);
  w.active = (uint8_t*)take((size_t)n_rays);
  w.origin = (float*)take(16);
  w.bytes = off;
  return w;
}

static int check_grid(const nerfb200_kilo_grid* g, const char* who) {
  NB_CHECK_ARG(g, "%s: null grid description", who);
  for (int c = 0; c < 3; ++c) NB_CHECK_ARG(g->res[c] >= 1 && g->res[c] <= 1024 && g->gmax[c] > g->gmin[c], "%s: bad grid axis %d", who, c);
  return 0;
}

}  // namespace kilo
}  // namespace nb

using namespace nb;
using namespace nb::kilo;

static int grid_for(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  return (int)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

extern "C" int nerfb200_kilo_param_size(void) { return kParamSize; }

extern "C" int nerfb200_kilo_rays_d(const nerfb200_kilo_camera* cam, float* dirs, void* stream) {
  NB_CHECK_ARG(cam && dirs && cam->H > 0 && cam->W > 0, "kilo_rays_d: bad arguments");
  rays_d_kernel<<<grid_for((long long)cam->H * cam->W, 256), 256, 0, (cudaStream_t)stream>>>(*cam, dirs, nullptr, 0.f);
  NB_LAUNCH_OK("kilo::rays_d_kernel");
  return 0;
}

extern "C" int nerfb200_kilo_march(const nerfb200_kilo_grid* g, const float* origin, const float* dirs, const int16_t* grid,
                                   int n_rays, float distance_between_points, int max_samples_per_ray, int max_depth_index,
                                   float min_distance, int is_initial_query, int32_t* query_indices, int16_t* assigned_networks,
                                   uint8_t* active_ray_mask, int32_t* depth_indices, void* stream) {
  if (check_grid(g, "kilo_march")) return 1;
  NB_CHECK_ARG(n_rays >= 0 && max_samples_per_ray >= 1 && max_depth_index >= 1, "kilo_march: bad sizes");
  NB_CHECK_ARG((long long)n_rays * max_depth_index < (1LL << 31), "kilo_march: n_rays * max_depth_index overflows the int32 query index");
  if (n_rays == 0) return 0;
  NB_CHECK_ARG(origin && dirs && grid && query_indices && assigned_networks && active_ray_mask && depth_indices, "kilo_march: null pointer");
  march_kernel<<<grid_for(n_rays, 128), 128, 0, (cudaStream_t)stream>>>(*g, origin, dirs, grid, query_indices, assigned_networks,
                                                                       active_ray_mask, depth_indices, n_rays, distance_between_points,
                                                                       max_samples_per_ray, max_depth_index, min_distance,
                                                                       is_initial_query, nullptr);
  NB_LAUNCH_OK("kilo::march_kernel");
  return 0;
}

extern "C" int nerfb200_kilo_integrate(const float* rgb_sigma, const int16_t* assigned_networks, const float* dists, int n_rays,
                                       int samples_per_ray, float transmittance_threshold, int is_initial_query, float* rgb_map,
                                       float* acc_map, float* transmittance, uint8_t* active_ray_mask, void* stream) {
  NB_CHECK_ARG(n_rays >= 0 && samples_per_ray >= 1, "kilo_integrate: bad sizes");
  if (n_rays == 0) return 0;
  NB_CHECK_ARG(rgb_sigma && assigned_networks && dists && rgb_map && acc_map && transmittance && active_ray_mask, "kilo_integrate: null pointer");
  NB_CHECK_ARG(((uintptr_t)rgb_sigma & 15) == 0, "kilo_integrate: rgb_sigma must be 16-byte aligned");
  integrate_kernel<<<grid_for(n_rays, 128), 128, 0, (cudaStream_t)stream>>>((const float4*)rgb_sigma, assigned_networks, dists, rgb_map,
                                                                           acc_map, transmittance, active_ray_mask, n_rays,
                                                                           samples_per_ray, transmittance_threshold, is_initial_query, nullptr);
  NB_LAUNCH_OK("kilo::integrate_kernel");
  return 0;
}

extern "C" size_t nerfb200_kilo_workspace_bytes(int n_rays, int max_samples_per_ray, int num_networks) {
  if (n_rays <= 0 || max_samples_per_ray <= 0 || num_networks <= 0) return 0;
  return carve(nullptr, n_rays, max_samples_per_ray, num_networks).bytes;
}

// sort + evaluate the queries sitting in (query, assigned): shared by the kernel-level entry and the renderer
static int eval_pass(const Work& w, const nerfb200_kilo_camera* cam, const nerfb200_kilo_march_params* mp, int n_rays,
                     int num_networks, const float* params, const float* domain_mins, const float* domain_maxs, cudaStream_t st) {
  const size_t slots = (size_t)n_rays * mp->max_samples_per_ray;
  const int sort_blocks = (int)(slots / 8192 < 1 ? 1 : (slots / 8192 > 592 ? 592 : slots / 8192));
  const size_t sh = (size_t)num_networks * 4;
  hist_kernel<<<sort_blocks, 256, sh, st>>>(w.assigned, slots, num_networks, w.count, w.counters);
  NB_LAUNCH_OK("kilo::hist_kernel");
  scan_kernel<<<1, 1024, 0, st>>>(w.count, num_networks, w.start, w.cursor, w.item_start, w.counters);
  NB_LAUNCH_OK("kilo::scan_kernel");
  scatter_kernel<<<sort_blocks, 256, sh, st>>>(w.assigned, slots, num_networks, w.cursor, w.sorted, w.counters);
  NB_LAUNCH_OK("kilo::scatter_kernel");
  EvalCam ec;
  for (int i = 0; i < 9; ++i) ec.c2w[i] = cam->c2w[i];
  for (int i = 0; i < 3; ++i) ec.origin[i] = cam->origin[i];
  ec.cx = cam->cx; ec.cy = cam->cy; ec.fx = cam->fx; ec.fy = cam->fy;
  ec.W = cam->W; ec.max_depth = mp->max_depth_index; ec.min_distance = mp->min_distance; ec.dbp = mp->distance_between_points;
  int dev = 0, sms = 0;
  NB_CUDA(cudaGetDevice(&dev));
  NB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  NB_CHECK_ARG(dev >= 0 && dev < 64, "kilo: device ordinal %d out of range", dev);
  static bool attr_set[64] = {};   // opt-in to > 48 KB of dynamic shared memory: per device, sticky
  if (!attr_set[dev]) {
    NB_CUDA(cudaFuncSetAttribute(eval_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
    attr_set[dev] = true;
  }
  eval_tc_kernel<<<sms * 2, kTcThreads, kTcSmemBytes, st>>>(w.query, w.sorted, w.start, w.item_start, num_networks, params, domain_mins,
                                                            domain_maxs, ec, w.rgb_sigma, w.counters);
  NB_LAUNCH_OK("kilo::eval_tc_kernel");
  return 0;
}

static int check_common(const nerfb200_kilo_camera* cam, const nerfb200_kilo_grid* g, const nerfb200_kilo_march_params* mp,
                        int num_networks, const char* who) {
  NB_CHECK_ARG(cam && mp, "%s: null parameter block", who);
  if (check_grid(g, who)) return 1;
  NB_CHECK_ARG(cam->H > 0 && cam->W > 0 && cam->fx != 0.f && cam->fy != 0.f, "%s: bad camera", who);
  NB_CHECK_ARG(mp->max_samples_per_ray >= 1 && mp->max_depth_index >= 1 && mp->distance_between_points > 0.f, "%s: bad march parameters", who);
  NB_CHECK_ARG(num_networks >= 1 && num_networks <= 8192, "%s: num_networks=%d out of range [1, 8192] (shared-memory histograms)", who, num_networks);
  NB_CHECK_ARG((long long)cam->H * cam->W * mp->max_depth_index < (1LL << 31), "%s: H*W*max_depth_index overflows the int32 query index", who);
  NB_CHECK_ARG((long long)cam->H * cam->W * mp->max_samples_per_ray < (1LL << 31), "%s: H*W*max_samples_per_ray overflows int32 slots", who);
  return 0;
}

extern "C" int nerfb200_kilo_network_eval(const nerfb200_kilo_camera* cam, const nerfb200_kilo_march_params* mp,
                                          const int32_t* query_indices, const int16_t* assigned_networks, int n_rays,
                                          const float* params, const float* domain_mins, const float* domain_maxs, int num_networks,
                                          void* workspace, size_t workspace_bytes, float* rgb_sigma, void* stream) {
  nerfb200_kilo_grid dummy = {{1, 1, 1}, {0.f, 0.f, 0.f}, {1.f, 1.f, 1.f}};
  if (check_common(cam, &dummy, mp, num_networks, "kilo_network_eval")) return 1;
  if (n_rays <= 0) return 0;
  NB_CHECK_ARG(query_indices && assigned_networks && params && domain_mins && domain_maxs && workspace && rgb_sigma, "kilo_network_eval: null pointer");
  NB_CHECK_ARG(((uintptr_t)params & 15) == 0 && ((uintptr_t)rgb_sigma & 15) == 0 && ((uintptr_t)workspace & 255) == 0, "kilo_network_eval: misaligned buffer");
  Work w = carve(workspace, n_rays, mp->max_samples_per_ray, num_networks);
  NB_CHECK_ARG(workspace_bytes >= w.bytes, "kilo_network_eval: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t slots = (size_t)n_rays * mp->max_samples_per_ray;
  w.query = const_cast<int32_t*>(query_indices);
  w.assigned = const_cast<int16_t*>(assigned_networks);
  w.rgb_sigma = reinterpret_cast<float4*>(rgb_sigma);
  NB_CUDA(cudaMemsetAsync(w.count, 0, (size_t)(num_networks + 1) * 4, st));
  NB_CUDA(cudaMemsetAsync(w.counters, 0, C_COUNT * 4, st));
  NB_CUDA(cudaMemsetAsync(rgb_sigma, 0, slots * 16, st));
  const int32_t one = 1;   // "there may be queries": the histogram decides
  NB_CUDA(cudaMemcpyAsync(w.counters + C_QUERIES, &one, 4, cudaMemcpyHostToDevice, st));
  return eval_pass(w, cam, mp, n_rays, num_networks, params, domain_mins, domain_maxs, st);
}

extern "C" int nerfb200_kilo_render(const nerfb200_kilo_camera* cam, const nerfb200_kilo_grid* g, const nerfb200_kilo_march_params* mp,
                                    const int16_t* occupancy_grid, const float* params, const float* domain_mins,
                                    const float* domain_maxs, int num_networks, void* workspace, size_t workspace_bytes,
                                    float* rgb_map, float* acc_map, int64_t* stats, void* stream) {
  if (check_common(cam, g, mp, num_networks, "kilo_render")) return 1;
  NB_CHECK_ARG(mp->max_passes >= 1, "kilo_render: max_passes must be >= 1");
  NB_CHECK_ARG(occupancy_grid && params && domain_mins && domain_maxs && workspace && rgb_map && acc_map, "kilo_render: null pointer");
  NB_CHECK_ARG(((uintptr_t)params & 15) == 0 && ((uintptr_t)workspace & 255) == 0, "kilo_render: misaligned buffer");
  const int n_rays = cam->H * cam->W, spp = mp->max_samples_per_ray;
  Work w = carve(workspace, n_rays, spp, num_networks);
  NB_CHECK_ARG(workspace_bytes >= w.bytes, "kilo_render: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  NB_CUDA(cudaMemsetAsync(w.count, 0, (size_t)(num_networks + 1) * 4, st));
  NB_CUDA(cudaMemsetAsync(w.counters, 0, C_COUNT * 4, st));
  NB_CUDA(cudaMemcpyAsync(w.origin, cam->origin, 12, cudaMemcpyHostToDevice, st));
  rays_d_kernel<<<grid_for(n_rays, 256), 256, 0, st>>>(*cam, w.dirs, w.dists, mp->distance_between_points);
  NB_LAUNCH_OK("kilo::rays_d_kernel");
  for (int p = 0; p < mp->max_passes; ++p) {
    march_kernel<<<grid_for(n_rays, 128), 128, 0, st>>>(*g, w.origin, w.dirs, occupancy_grid, w.query, w.assigned, w.active, w.depth_idx,
                                                        n_rays, mp->distance_between_points, spp, mp->max_depth_index, mp->min_distance,
                                                        p == 0, w.counters);
    NB_LAUNCH_OK("kilo::march_kernel");
    int rc = eval_pass(w, cam, mp, n_rays, num_networks, params, domain_mins, domain_maxs, st);
    if (rc) return rc;
    integrate_kernel<<<grid_for(n_rays, 128), 128, 0, st>>>(w.rgb_sigma, w.assigned, w.dists, rgb_map, acc_map, w.T, w.active, n_rays, spp,
                                                            mp->transmittance_threshold, p == 0, w.counters);
    NB_LAUNCH_OK("kilo::integrate_kernel");
    next_pass_kernel<<<(num_networks + 255) / 256, 256, 0, st>>>(w.counters, w.count, num_networks, stats);
    NB_LAUNCH_OK("kilo::next_pass_kernel");
  }
  if (mp->white_bkgd) {
    background_kernel<<<grid_for(n_rays, 256), 256, 0, st>>>(rgb_map, acc_map, n_rays, 1.f, 1.f, 1.f);
    NB_LAUNCH_OK("kilo::background_kernel");
  }
  return 0;
}
