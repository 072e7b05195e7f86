// a1 ray generation, a2 stratified sampling, a4 sample_pdf + merge.
// Reference: src/models/nerf/renderer/volume_renderer.py:115-147, :218-237, :239-268, :181-183.
//
// All of these are HBM-bound streaming kernels (a few hundred bytes per ray); they are
// written warp-per-ray with coalesced row accesses and no shared-memory staging beyond the
// per-warp cdf/bins rows.  Arithmetic that feeds the positional encoding or the bin search
// uses explicit __fmul_rn/__fadd_rn so nvcc cannot contract it into FMAs: the reference
// (torch CPU, one op per kernel) rounds after every op, and PE amplifies 1 ulp of a
// coordinate by 2^9.
#include "common.cuh"

namespace nb {

// ------------------------------------------------------------------------------------------
// a1: rays from pose / intrinsics.  Bit-exact restatement of the torch CPU arithmetic
// (probed): dirs = ((x-cx)/fx, -(y-cy)/fy, -1); d_c = (dirs0*R[c][0] + dirs1*R[c][1]) +
// dirs2*R[c][2]; norm = sqrt(fma(d2,d2,fma(d1,d1,d0*d0))) (ATen's norm kernel accumulates
// with FMAs); d /= norm.
// ------------------------------------------------------------------------------------------
__global__ void raygen_kernel(const float* __restrict__ pose, const float* __restrict__ K, int H,
                              int W, float* __restrict__ rays_o, float* __restrict__ rays_d) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * W) return;
  int y = idx / W, x = idx - y * W;
  float fx = K[0], cx = K[2], fy = K[4], cy = K[5];
  float d0 = __fdiv_rn(__fsub_rn((float)x, cx), fx);
  float d1 = -__fdiv_rn(__fsub_rn((float)y, cy), fy);
  float d2 = -1.0f;
  float r[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float p0 = __fmul_rn(d0, pose[c * 4 + 0]);
    float p1 = __fmul_rn(d1, pose[c * 4 + 1]);
    float p2 = __fmul_rn(d2, pose[c * 4 + 2]);
    r[c] = __fadd_rn(__fadd_rn(p0, p1), p2);
  }
  float n2 = __fmaf_rn(r[2], r[2], __fmaf_rn(r[1], r[1], __fmul_rn(r[0], r[0])));
  float n = __fsqrt_rn(n2);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    rays_d[idx * 3 + c] = __fdiv_rn(r[c], n);
    rays_o[idx * 3 + c] = pose[c * 4 + 3];
  }
}

// ------------------------------------------------------------------------------------------
// a2: coarse z.  perturb==0 broadcasts the caller's table (bit-identical to the reference's
// linspace arithmetic).  perturb!=0: mids/upper/lower as :228-235 with a counter-based hash
// RNG (the reference's torch.rand stream cannot be reproduced by a kernel).
// ------------------------------------------------------------------------------------------
__global__ void sample_coarse_kernel(const float* __restrict__ z_table, long long total, int S,
                                     int perturb, uint64_t seed, float* __restrict__ z_vals) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int s = (int)(idx % S);
  float z = z_table[s];
  if (perturb) {
    long long ray = idx / S;
    float zl = z_table[max(s - 1, 0)], zu = z_table[min(s + 1, S - 1)];
    float lower = (s == 0) ? z : __fmul_rn(0.5f, __fadd_rn(z, zl));
    float upper = (s == S - 1) ? z : __fmul_rn(0.5f, __fadd_rn(zu, z));
    float t = uniform01(seed, (uint32_t)ray, (uint32_t)s);
    z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t));
  }
  z_vals[idx] = z;
}

// ------------------------------------------------------------------------------------------
// a4 core: inverse-CDF lookup for one u, given a warp-private cdf/bins row in shared memory.
// searchsorted(cdf, u, right=True) = first index with cdf[idx] > u, as a branch-free binary
// search over nb entries.  Then :255-266 literally.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int upper_bound_row(const float* cdf, int nbins, float u) {
  int lo = 0, hi = nbins;  // answer in [0, nbins]
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (cdf[mid] > u) hi = mid; else lo = mid + 1;
  }
  return lo;
}

__device__ __forceinline__ float invert_cdf(const float* cdf, const float* bins, int nbins, float u,
                                            int* ind_out) {
  int ind = upper_bound_row(cdf, nbins, u);
  int below = max(ind - 1, 0);
  int above = min(nbins - 1, ind);
  float c0 = cdf[below], c1 = cdf[above];
  float b0 = bins[below], b1 = bins[above];
  float denom = __fsub_rn(c1, c0);
  if (denom < 1e-5f) denom = 1.0f;
  float t = __fdiv_rn(__fsub_rn(u, c0), denom);
  *ind_out = ind;
  return __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
}

constexpr int kWarpsPerBlock = 8;
constexpr int kMaxBins = 256;  // S-1 <= 256

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
sample_from_cdf_kernel(const float* __restrict__ cdf_g, const float* __restrict__ bins_g,
                       const float* __restrict__ u_g, int u_per_ray, int n_rays, int nbins, int n_u,
                       float* __restrict__ samples, int32_t* __restrict__ inds) {
  __shared__ float s_cdf[kWarpsPerBlock][kMaxBins];
  __shared__ float s_bins[kWarpsPerBlock][kMaxBins];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int ray = blockIdx.x * kWarpsPerBlock + warp;
  if (ray >= n_rays) return;
  for (int i = lane; i < nbins; i += 32) {
    s_cdf[warp][i] = cdf_g[(size_t)ray * nbins + i];
    s_bins[warp][i] = bins_g[(size_t)ray * nbins + i];
  }
  __syncwarp();
  const float* u_row = u_per_ray ? u_g + (size_t)ray * n_u : u_g;
  for (int i = lane; i < n_u; i += 32) {
    int ind;
    float v = invert_cdf(s_cdf[warp], s_bins[warp], nbins, u_row[i], &ind);
    samples[(size_t)ray * n_u + i] = v;
    if (inds) inds[(size_t)ray * n_u + i] = ind;
  }
}

// ------------------------------------------------------------------------------------------
// a4 whole: weights[1:-1] + 1e-5 -> pdf -> cdf -> samples -> sorted merge with the coarse z.
//  * normaliser: summed in torch-CPU's own fp32 order (pdf_normaliser below), so the cdf is bit-identical to
//    the reference's whenever the coarse weights are.
//  * cdf: torch CPU cumsum accumulates float in DOUBLE and rounds each output (probed, 0
//    mismatches) -- reproduced exactly with a warp scan in fp64.
//  * merge: rank = own index + count of elements of the other list that sort before it
//    (binary search); eval-mode u is monotone so the samples are already sorted.  With
//    per-ray random u the samples are ranked among themselves by counting.
// One warp per ray; rows staged in shared memory.
// ------------------------------------------------------------------------------------------
constexpr int kMaxS = 256;
constexpr int kMaxU = 256;

__device__ __forceinline__ double warp_incl_scan(double v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    double o = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += o;
  }
  return v;
}

// sum_i (w[i] + 1e-5) over the n interior weights of one ray, in EXACTLY the order torch's CPU kernel uses for
// `torch.sum(weights, -1)` on a contiguous fp32 row (volume_renderer.py:242; ATen SumKernel.cpp vectorized_inner_sum,
// probed on torch 2.11 at the DEFAULT, AVX2 and AVX512 dispatch levels -- all three use 8-lane vectors -- for
// n = 8..254): the row is cut into vectors of 8; vector 4i+k (i < nv/4) goes to accumulator k of four, the remaining
// vectors to accumulator 0; the accumulators are added ((a0+a1)+a2)+a3 lane-wise; the result starts from the scalar
// tail x[8 nv ..] summed left to right and then adds lanes 0..7 in order.  With it the cdf -- hence every bin index of
// sample_pdf -- is bit-identical to the reference's whenever the coarse weights are (before: fp64 sum, 1 ulp off on 42 %
// of the rows).  The 32 (accumulator, lane) chains map onto the 32 lanes of the warp.  n < 8 (n_samples < 10) takes
// another path inside torch; the correctly rounded fp64 sum is used there.
template <int Q>   // Q = ceil(n / 32) at most: the per-lane loads are unrolled (2 for the 64-sample configuration, 8 in general)
__device__ __forceinline__ float pdf_normaliser(const float* __restrict__ w, int n, int lane) {
#ifdef NERFB200_PDF_FP64_SUM   // A/B experiments only: the round-1 normaliser (correctly rounded fp64 sum, 1 ulp from torch on 42 % of the rows)
  const bool fp64_sum = true;
#else
  const bool fp64_sum = false;
#endif
  if (n < 8 || fp64_sum) {
    double part = 0.0;
    for (int i = lane; i < n; i += 32) part += (double)__fadd_rn(w[i], 1e-5f);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    return (float)part;
  }
  // Lane l holds x_q = w[l + 32 q] + 1e-5: element e = 8 vec + v sits in lane e % 32 = 8 (vec % 4) + v at q = vec / 4,
  // i.e. exactly in the (accumulator k = lane / 8, vector lane v = lane % 8) chain torch assigns it to.  All loads are
  // issued up front (independent, coalesced); what follows is a chain of register adds fed by shuffles whose sources
  // are ready early (no loop with a data-dependent trip count: measured 1.61 -> ~1.1 ms per 640 000 rays).
  const int v = lane & 7;
  const int nv = n >> 3, size_ilp = nv >> 2;       // full vectors; full rows of four vectors
  const int rem = nv - (size_ilp << 2);            // full vectors in the last, partial row (0..3)
  const int ntail = n - (nv << 3);                 // scalar tail (0..7)
  float x[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int e = lane + 32 * q;
    x[q] = e < n ? __fadd_rn(w[e], 1e-5f) : 0.f;
  }
  float acc = 0.f, xr = 0.f;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    if (q < size_ilp) acc = __fadd_rn(acc, x[q]);
    if (q == size_ilp) xr = x[q];
  }
  // remaining full vectors of the partial row go to accumulator 0 (lanes 0..7), in order
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float t = __shfl_sync(0xffffffffu, xr, 8 * r + v);
    if (r < rem && lane < 8) acc = __fadd_rn(acc, t);
  }
  const float a1 = __shfl_sync(0xffffffffu, acc, v + 8), a2 = __shfl_sync(0xffffffffu, acc, v + 16),
              a3 = __shfl_sync(0xffffffffu, acc, v + 24);
  acc = __fadd_rn(__fadd_rn(__fadd_rn(acc, a1), a2), a3);   // meaningful on lanes 0..7
  float fin = 0.f;
#pragma unroll
  for (int t = 0; t < 7; ++t) {
    const float y = __shfl_sync(0xffffffffu, xr, (8 * rem + t) & 31);
    if (t < ntail) fin = __fadd_rn(fin, y);
  }
#pragma unroll
  for (int l = 0; l < 8; ++l) fin = __fadd_rn(fin, __shfl_sync(0xffffffffu, acc, l));
  return fin;
}

// number of entries of sorted row a[0..n) that are < x (strict) or <= x
__device__ __forceinline__ int count_less(const float* a, int n, float x, bool or_equal) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    bool before = or_equal ? (a[mid] <= x) : (a[mid] < x);
    if (before) lo = mid + 1; else hi = mid;
  }
  return lo;
}

template <int Q>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 6)
sample_pdf_merge_kernel(const float* __restrict__ z_coarse, const float* __restrict__ weights,
                        const float* __restrict__ u_g, int u_per_ray, int n_rays, int S, int n_u,
                        float* __restrict__ z_all, float* __restrict__ z_samples,
                        int32_t* __restrict__ inds, float* __restrict__ cdf_out,
                        const int32_t* __restrict__ ray_list, const int32_t* __restrict__ n_list) {
  __shared__ float s_z[kWarpsPerBlock][kMaxS];
  __shared__ float s_cdf[kWarpsPerBlock][kMaxS];
  __shared__ float s_bins[kWarpsPerBlock][kMaxS];
  __shared__ float s_smp[kWarpsPerBlock][kMaxU];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ray_list != NULL: persistent grid over the rays that survived the culling (common.cuh: RayList)
  const int n_eff = ray_list != nullptr ? *n_list : n_rays;
  for (int it = blockIdx.x * kWarpsPerBlock + warp; it < n_eff; it += gridDim.x * kWarpsPerBlock) {
  const int ray = ray_list != nullptr ? ray_list[it] : it;
  const int nbins = S - 1;   // cdf / t_mid entries (63)
  const int nw = S - 2;      // interior weights (62)
  float* zr = s_z[warp];
  float* cdf = s_cdf[warp];
  float* bins = s_bins[warp];
  float* smp = s_smp[warp];
  for (int i = lane; i < S; i += 32) zr[i] = z_coarse[(size_t)ray * S + i];
  __syncwarp();
  for (int i = lane; i < nbins; i += 32) bins[i] = __fmul_rn(0.5f, __fadd_rn(zr[i + 1], zr[i]));
  // pdf normaliser, in torch-CPU's summation order (see pdf_normaliser)
  const float wsum = pdf_normaliser<Q>(weights + (size_t)ray * S + 1, nw, lane);
  // cdf = [0, cumsum(pdf)] with fp64 accumulation, each lane owns a contiguous segment
  const int per = (nw + 31) / 32;
  double local = 0.0;
  float pdf_loc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int i = lane * per + j;
    float p = 0.f;
    if (j < per && i < nw) p = __fdiv_rn(__fadd_rn(weights[(size_t)ray * S + 1 + i], 1e-5f), wsum);
    pdf_loc[j] = p;
    local += (double)p;
  }
  double incl = warp_incl_scan(local, lane);
  double run = incl - local;
  if (lane == 0) cdf[0] = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int i = lane * per + j;
    if (j < per && i < nw) {
      run += (double)pdf_loc[j];
      cdf[i + 1] = (float)run;
    }
  }
  __syncwarp();
  if (cdf_out)
    for (int i = lane; i < nbins; i += 32) cdf_out[(size_t)ray * nbins + i] = cdf[i];
  // inverse CDF
  const float* u_row = u_per_ray ? u_g + (size_t)ray * n_u : u_g;
  for (int i = lane; i < n_u; i += 32) {
    int ind;
    float v = invert_cdf(cdf, bins, nbins, u_row[i], &ind);
    smp[i] = v;
    if (z_samples) z_samples[(size_t)ray * n_u + i] = v;
    if (inds) inds[(size_t)ray * n_u + i] = ind;
  }
  __syncwarp();
  // merge (values only are compared downstream, so tie order is free; ties are broken
  // coarse-first to make ranks a permutation)
  float* out = z_all + (size_t)ray * (S + n_u);
  // monotone u gives sorted samples except for rare 1-ulp inversions at bin edges
  // (b0 + fl(b1-b0) may exceed b1): check, and fall back to counting ranks if unsorted.
  bool unsorted = false;
  for (int i = lane + 1; i < n_u; i += 32) unsorted |= (smp[i] < smp[i - 1]);
  unsorted = __any_sync(0xffffffffu, unsorted);
  if (!unsorted) {
    for (int i = lane; i < S; i += 32) out[i + count_less(smp, n_u, zr[i], false)] = zr[i];
    for (int i = lane; i < n_u; i += 32) out[i + count_less(zr, S, smp[i], true)] = smp[i];
  } else {
    for (int i = lane; i < S; i += 32) {
      float x = zr[i];
      int c = 0;
      for (int j = 0; j < n_u; ++j) c += (smp[j] < x);
      out[i + c] = x;
    }
    for (int i = lane; i < n_u; i += 32) {
      float x = smp[i];
      int c = count_less(zr, S, x, true);
      for (int j = 0; j < n_u; ++j) c += (smp[j] < x) || (smp[j] == x && j < i);
      out[c] = x;
    }
  }
  __syncwarp();   // the per-warp rows are reused by the next ray of the list
  }
}

// ------------------------------------------------------------------------------------------
// a7 with the reference's NON-detached sampler (volume_renderer.py:181-183 feeds sample_pdf's output, :239-268,
// into the fine pass under autograd): gradient of the merged depths z_all with respect to the coarse weights.
//   samples_j = b0 + t (b1 - b0),  t = (u_j - c0) / denom,  denom = c1 - c0 (replaced by the constant 1 below 1e-5)
//   cdf = [0, cumsum(pdf)],  pdf = (w + 1e-5) / sum(w + 1e-5)
// The kernel recomputes the forward (cdf, samples, merge positions) exactly as sample_pdf_merge_kernel does, so
// nothing has to be saved, then:  g_cdf (shared-memory fp64 atomics: several u fall into one bin) -> reverse
// cumulative sum -> g_pdf -> quotient rule -> g_weights[:, 1:-1]; the first and last coarse weight get 0.
// The coarse depths inside z_all are constants (no parameter reaches them) and receive no gradient.
// ------------------------------------------------------------------------------------------
template <int Q>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
sample_pdf_backward_kernel(const float* __restrict__ z_coarse, const float* __restrict__ weights,
                           const float* __restrict__ u_g, int u_per_ray, int n_rays, int S, int n_u,
                           const float* __restrict__ g_z_all, float* __restrict__ g_weights) {
  __shared__ float s_z[kWarpsPerBlock][kMaxS];
  __shared__ float s_cdf[kWarpsPerBlock][kMaxS];
  __shared__ float s_bins[kWarpsPerBlock][kMaxS];
  __shared__ float s_smp[kWarpsPerBlock][kMaxU];
  __shared__ double s_gcdf[kWarpsPerBlock][kMaxS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kWarpsPerBlock + warp;
  if (ray >= n_rays) return;
  const int nbins = S - 1, nw = S - 2;
  float* zr = s_z[warp];
  float* cdf = s_cdf[warp];
  float* bins = s_bins[warp];
  float* smp = s_smp[warp];
  double* gcdf = s_gcdf[warp];
  for (int i = lane; i < S; i += 32) { zr[i] = z_coarse[(size_t)ray * S + i]; gcdf[i] = 0.0; }
  __syncwarp();
  for (int i = lane; i < nbins; i += 32) bins[i] = __fmul_rn(0.5f, __fadd_rn(zr[i + 1], zr[i]));
  // ---- forward recompute (same arithmetic as sample_pdf_merge_kernel)
  const float wsum = pdf_normaliser<Q>(weights + (size_t)ray * S + 1, nw, lane);
  const int per = (nw + 31) / 32;
  double local = 0.0;
  float pdf_loc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int i = lane * per + j;
    float p = 0.f;
    if (j < per && i < nw) p = __fdiv_rn(__fadd_rn(weights[(size_t)ray * S + 1 + i], 1e-5f), wsum);
    pdf_loc[j] = p;
    local += (double)p;
  }
  double incl = warp_incl_scan(local, lane);
  double run = incl - local;
  if (lane == 0) cdf[0] = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int i = lane * per + j;
    if (j < per && i < nw) {
      run += (double)pdf_loc[j];
      cdf[i + 1] = (float)run;
    }
  }
  __syncwarp();
  const float* u_row = u_per_ray ? u_g + (size_t)ray * n_u : u_g;
  for (int i = lane; i < n_u; i += 32) {
    int ind;
    smp[i] = invert_cdf(cdf, bins, nbins, u_row[i], &ind);
  }
  __syncwarp();
  bool unsorted = false;
  for (int i = lane + 1; i < n_u; i += 32) unsorted |= (smp[i] < smp[i - 1]);
  unsorted = __any_sync(0xffffffffu, unsorted);
  // ---- d samples / d cdf
  const float* gz_row = g_z_all + (size_t)ray * (S + n_u);
  for (int i = lane; i < n_u; i += 32) {
    const float x = smp[i];
    int pos = count_less(zr, S, x, true);
    if (!unsorted) pos += i;
    else
      for (int j = 0; j < n_u; ++j) pos += (smp[j] < x) || (smp[j] == x && j < i);
    const float gz = gz_row[pos];
    const float u = u_row[i];
    const int ind = upper_bound_row(cdf, nbins, u);
    const int below = max(ind - 1, 0), above = min(nbins - 1, ind);
    const float c0 = cdf[below], c1 = cdf[above];
    const float db = __fsub_rn(bins[above], bins[below]);
    float denom = __fsub_rn(c1, c0);
    const bool repl = denom < 1e-5f;
    if (repl) denom = 1.0f;
    const float t = __fdiv_rn(__fsub_rn(u, c0), denom);
    const double gt = (double)gz * (double)db;           // dL/dt
    if (repl) {
      atomicAdd(&gcdf[below], -gt);
    } else {
      atomicAdd(&gcdf[below], gt * ((double)t - 1.0) / (double)denom);
      atomicAdd(&gcdf[above], -gt * (double)t / (double)denom);
    }
  }
  __syncwarp();
  // ---- g_pdf[m] = sum_{k > m} g_cdf[k]  (cdf[k] = sum_{m < k} pdf[m]);  own segment m = lane*per + j
  double seg = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int m = lane * per + j;
    if (j < per && m < nw) seg += gcdf[m + 1];
  }
  double seg_incl = warp_incl_scan(seg, lane);
  double total = __shfl_sync(0xffffffffu, seg_incl, 31);
  double suffix = total - seg_incl;     // g_cdf of the entries owned by later lanes
  double gpdf[8];
  double dot = 0.0;
#pragma unroll
  for (int j = 7; j >= 0; --j) {
    int m = lane * per + j;
    gpdf[j] = 0.0;
    if (j < per && m < nw) {
      suffix += gcdf[m + 1];
      gpdf[j] = suffix;
      dot += suffix * (double)pdf_loc[j];
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, d);
  float* gw = g_weights + (size_t)ray * S;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int m = lane * per + j;
    if (j < per && m < nw) gw[1 + m] = (float)((gpdf[j] - dot) / (double)wsum);
  }
  if (lane == 0) { gw[0] = 0.f; gw[S - 1] = 0.f; }
}

}  // namespace nb

using namespace nb;

static int sample_pdf_merge_impl(const float* z_coarse, const float* weights, const float* u, int u_per_ray,
                                 RayList rl, int n_rays, int n_samples, int n_u, float* z_all,
                                 float* z_samples, int32_t* inds, float* cdf, void* stream);

extern "C" int nerfb200_raygen(const float* pose, const float* intrinsics, int H, int W,
                               float* rays_o, float* rays_d, void* stream) {
  NB_CHECK_ARG(pose && intrinsics && rays_o && rays_d, "raygen: null pointer");
  NB_CHECK_ARG(H > 0 && W > 0 && (long long)H * W < (1LL << 30), "raygen: bad H=%d W=%d", H, W);
  int n = H * W;
  raygen_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(pose, intrinsics, H, W, rays_o, rays_d);
  NB_LAUNCH_OK("raygen_kernel");
  return 0;
}

extern "C" int nerfb200_sample_coarse(const float* z_table, int n_rays, int n_samples, int perturb,
                                      uint64_t seed, float* z_vals, void* stream) {
  NB_CHECK_ARG(n_rays <= 0 || (z_table && z_vals), "sample_coarse: null pointer");
  NB_CHECK_ARG(n_rays >= 0 && n_samples >= 2, "sample_coarse: bad sizes n_rays=%d S=%d", n_rays, n_samples);
  if (n_rays == 0) return 0;
  long long total = (long long)n_rays * n_samples;
  sample_coarse_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(z_table, total, n_samples, perturb, seed, z_vals);
  NB_LAUNCH_OK("sample_coarse_kernel");
  return 0;
}

extern "C" int nerfb200_sample_from_cdf(const float* cdf, const float* bins, const float* u,
                                        int u_per_ray, int n_rays, int n_bins, int n_u,
                                        float* samples, int32_t* inds, void* stream) {
  NB_CHECK_ARG(n_rays <= 0 || n_u <= 0 || (cdf && bins && u && samples), "sample_from_cdf: null pointer");
  NB_CHECK_ARG(n_bins >= 1 && n_bins <= kMaxBins, "sample_from_cdf: n_bins=%d out of range [1,%d]", n_bins, kMaxBins);
  NB_CHECK_ARG(n_rays >= 0 && n_u >= 0, "sample_from_cdf: negative size");
  if (n_rays == 0 || n_u == 0) return 0;
  sample_from_cdf_kernel<<<ceil_div(n_rays, kWarpsPerBlock), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
      cdf, bins, u, u_per_ray, n_rays, n_bins, n_u, samples, inds);
  NB_LAUNCH_OK("sample_from_cdf_kernel");
  return 0;
}

extern "C" int nerfb200_sample_pdf_merge(const float* z_coarse, const float* weights, const float* u,
                                         int u_per_ray, int n_rays, int n_samples, int n_u,
                                         float* z_all, float* z_samples, int32_t* inds, float* cdf,
                                         void* stream) {
  return sample_pdf_merge_impl(z_coarse, weights, u, u_per_ray, RayList{nullptr, nullptr}, n_rays, n_samples, n_u, z_all,
                               z_samples, inds, cdf, stream);
}

int nb::sample_pdf_merge_culled(const float* z_coarse, const float* weights, const float* u, int u_per_ray,
                                RayList rl, int n_rays, int n_samples, int n_u, float* z_all, void* stream) {
  return sample_pdf_merge_impl(z_coarse, weights, u, u_per_ray, rl, n_rays, n_samples, n_u, z_all, nullptr,
                               nullptr, nullptr, stream);
}

static int sample_pdf_merge_impl(const float* z_coarse, const float* weights, const float* u, int u_per_ray,
                                 RayList rl, int n_rays, int n_samples, int n_u, float* z_all,
                                 float* z_samples, int32_t* inds, float* cdf, void* stream) {
  NB_CHECK_ARG(n_rays <= 0 || (z_coarse && weights && u && z_all), "sample_pdf_merge: null pointer");
  NB_CHECK_ARG(n_samples >= 3 && n_samples <= kMaxS, "sample_pdf_merge: n_samples=%d out of range [3,%d]", n_samples, kMaxS);
  NB_CHECK_ARG(n_u >= 1 && n_u <= kMaxU, "sample_pdf_merge: n_u=%d out of range [1,%d]", n_u, kMaxU);
  NB_CHECK_ARG(n_rays >= 0, "sample_pdf_merge: negative n_rays");
  if (n_rays == 0) return 0;
  int blocks = ceil_div(n_rays, kWarpsPerBlock);
  if (rl.rays != nullptr && blocks > kPersistentBlocks) blocks = kPersistentBlocks;
  if (n_samples - 2 <= 64)
    sample_pdf_merge_kernel<2><<<blocks, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        z_coarse, weights, u, u_per_ray, n_rays, n_samples, n_u, z_all, z_samples, inds, cdf, rl.rays, rl.count);
  else
    sample_pdf_merge_kernel<8><<<blocks, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        z_coarse, weights, u, u_per_ray, n_rays, n_samples, n_u, z_all, z_samples, inds, cdf, rl.rays, rl.count);
  NB_LAUNCH_OK("sample_pdf_merge_kernel");
  return 0;
}

extern "C" int nerfb200_sample_pdf_backward(const float* z_coarse, const float* weights, const float* u, int u_per_ray,
                                            int n_rays, int n_samples, int n_u, const float* g_z_all,
                                            float* g_weights, void* stream) {
  NB_CHECK_ARG(n_rays <= 0 || (z_coarse && weights && u && g_z_all && g_weights), "sample_pdf_backward: null pointer");
  NB_CHECK_ARG(n_samples >= 3 && n_samples <= kMaxS, "sample_pdf_backward: n_samples=%d out of range [3,%d]", n_samples, kMaxS);
  NB_CHECK_ARG(n_u >= 1 && n_u <= kMaxU, "sample_pdf_backward: n_u=%d out of range [1,%d]", n_u, kMaxU);
  NB_CHECK_ARG(n_rays >= 0, "sample_pdf_backward: negative n_rays");
  if (n_rays == 0) return 0;
  if (n_samples - 2 <= 64)
    sample_pdf_backward_kernel<2><<<ceil_div(n_rays, kWarpsPerBlock), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        z_coarse, weights, u, u_per_ray, n_rays, n_samples, n_u, g_z_all, g_weights);
  else
    sample_pdf_backward_kernel<8><<<ceil_div(n_rays, kWarpsPerBlock), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        z_coarse, weights, u, u_per_ray, n_rays, n_samples, n_u, g_z_all, g_weights);
  NB_LAUNCH_OK("sample_pdf_backward_kernel");
  return 0;
}
