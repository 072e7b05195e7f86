// a5/a6 alpha compositing forward (+ERT variants) and a7 analytic backward.
// Reference: volume_renderer.py:286-357 (_raw2outputs), :1089-1157 (_raw2outputs_with_ert).
//
// One warp per ray; every lane owns a contiguous run of samples, the transmittance is an
// exclusive prefix product done as lane-local products + a warp shuffle scan.  The scan and
// the map reductions run in fp64: torch CPU's cumprod accumulates float in double and rounds
// every output (probed: 0 mismatches), so T_i here is bit-identical to the reference's given
// the same alphas.  HBM-bound: reads 20 B/sample (raw float4 + z), writes 4 B/sample
// (weights) + 24 B/ray.
#include "common.cuh"

namespace nb {

constexpr int kCompWarps = 8;
// Minimum resident blocks per SM the compositors are compiled for, i.e. a register bound: these kernels are latency
// bound (strided row reads, shuffle scans), so warps in flight matter more than registers per thread.  Measured on
// B200 (scripts/stream_kernels.py, 640 000 rays; unbounded = 63-80 registers, 3 blocks per SM):
//   fast-math variants (MUFU, fp32 scans):   6 blocks (40 registers)  S=64 0.49 -> 0.38 ms,  S=192 0.76 -> 0.63 ms
//   exact variants (fp64 exp / prefix):      4 blocks (64 registers)  S=64 0.77,  S=192 1.63 ms  (6 blocks: 0.78 / 1.83, spills)
//   backward:                                4 blocks                 2.63 ms  (6 blocks: 3.09)
constexpr int kMinBlocksFast = 6, kMinBlocksExact = 4;
constexpr int kMaxPer = 8;  // ceil(S/32) <= 8  => S <= 256

// exp / sigmoid as the reference's CPU sees them: torch CPU evaluates exp with SLEEF (<= 1 ulp);
// rounding the fp64 exp matches it on 96.5-99% of arguments while CUDA's expf matches only
// 60-70% (scripts/probe_exp.py on the B200 box).  alpha = 1 - exp(-x) cancels, so each mismatch
// costs ulp(1) absolute on alpha.  sigmoid on CPU is 1/(1+exp(-x)) in fp32 steps.
// kFast (NERFB200_COMPOSITE_FAST_MATH, used by the bf16 mode whose MLP outputs carry 1e-3 already): MUFU-based
// __expf / __fdividef instead of the fp64 exp -- the exact variant spends 4 fp64 exps per sample and reaches
// only 0.22-0.26 of the HBM roofline at full-frame size (scripts/stream_kernels.py).
template <bool kFast>
__device__ __forceinline__ float exp_cr(float x) { return kFast ? __expf(x) : (float)exp((double)x); }
template <bool kFast>
__device__ __forceinline__ float sigmoid_ref(float x) {
  return kFast ? __fdividef(1.f, 1.f + __expf(-x)) : __fdiv_rn(1.f, __fadd_rn(1.f, exp_cr<false>(-x)));
}

__device__ __forceinline__ float ray_norm(const float* __restrict__ d) {
  // torch.norm(rays_d[..., None, :], dim=-1): FMA-chain accumulation (probed, bit-exact)
  return __fsqrt_rn(__fmaf_rn(d[2], d[2], __fmaf_rn(d[1], d[1], __fmul_rn(d[0], d[0]))));
}

// A = double in the exact variants (torch CPU's cumprod / sum accumulate in double), float in the fast-math
// variant: there the fp64 path is pure overhead -- two SHFLs per scan step, eight F2F conversions per sample
// (quarter-rate pipe) -- and made the kernel instruction-bound at the same 1.2 ms per frame whether or not the
// samples were skipped (ncu launch lists of the dense and the sparse frame).
template <typename A>
__device__ __forceinline__ A warp_excl_prod(A local, int lane) {
  A v = local;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    A o = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v *= o;
  }
  A ex = __shfl_up_sync(0xffffffffu, v, 1);
  return lane == 0 ? (A)1 : ex;
}

// Sums of eight per-lane values over the warp with 9 shuffles instead of 8 x 5: each butterfly step keeps half of
// the slots and sends the other half.  Returns the total of slot (lane >> 2), replicated over the slot's 4 lanes.
__device__ __forceinline__ float warp_reduce8(float (&v)[8], int lane) {
  bool up = (lane & 16) != 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float send = up ? v[j] : v[4 + j], keep = up ? v[4 + j] : v[j];
    v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  up = (lane & 8) != 0;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    float send = up ? v[j] : v[2 + j], keep = up ? v[2 + j] : v[j];
    v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  up = (lane & 4) != 0;
  float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
  float t = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  t += __shfl_xor_sync(0xffffffffu, t, 2);
  t += __shfl_xor_sync(0xffffffffu, t, 1);
  return t;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

template <bool kFast> struct AccT { typedef double type; };
template <> struct AccT<true> { typedef float type; };

struct RaySamples {
  float alpha[kMaxPer];
  float T[kMaxPer];     // transmittance before sample (fp32-rounded, as the reference sees it)
  bool low_any;         // (ERT) any T < threshold on this ray
  int first_low;        // (ERT) first sample index with T < threshold, S if none
};

// alpha + transmittance for the lane's samples.  eps = 1e-10 (PLAIN) or 0 (ERT variants).
// keep_bits (may be NULL): bit m of the array says whether row m = ray*S+i went through the MLP (sparse
// empty-space-skipping launch, nerfb200_ess_compact).  A cleared bit means zero density, and the row's raw
// entry is never read -- so the sparse path neither zero-fills raw nor streams the skipped rows (16 B each).
template <bool kErt, bool kFast = false>
__device__ __forceinline__ void ray_alpha_T(const float* __restrict__ raw_row,
                                            const float* __restrict__ z_row, float dnorm, int S,
                                            int per, int lane, float thr, RaySamples& rs,
                                            const uint32_t* __restrict__ keep_bits = nullptr, size_t bit_base = 0) {
  using A = typename AccT<kFast>::type;
  A local = (A)1;
  float fac[kMaxPer];
#pragma unroll
  for (int j = 0; j < kMaxPer; ++j) {
    int i = lane * per + j;
    float a = 0.f, f = 1.f;
    if (j < per && i < S) {
      float z0 = z_row[i];
      float dist = (i + 1 < S) ? __fsub_rn(z_row[i + 1], z0) : 1e10f;
      dist = __fmul_rn(dist, dnorm);
      bool kept = true;
      if (keep_bits != nullptr) {
        const size_t b = bit_base + (size_t)i;
        kept = (keep_bits[b >> 5] >> (b & 31)) & 1u;
      }
      float sig = kept ? fmaxf(raw_row[(size_t)i * 4 + 3], 0.f) : 0.f;
      // sigma == 0 (empty / skipped sample): exp(-0) = 1 and alpha = 0 exactly -- no fp64 exp needed
      a = sig == 0.f ? 0.f : __fsub_rn(1.f, exp_cr<kFast>(__fmul_rn(-sig, dist)));
      f = kErt ? __fsub_rn(1.f, a) : __fadd_rn(__fsub_rn(1.f, a), 1e-10f);
      local *= (A)f;
    }
    rs.alpha[j] = a;
    fac[j] = f;
  }
  A run = warp_excl_prod<A>(local, lane);
  int first = 1 << 30;
#pragma unroll
  for (int j = 0; j < kMaxPer; ++j) {
    int i = lane * per + j;
    rs.T[j] = (float)run;
    if (kErt && j < per && i < S && rs.T[j] < thr) first = min(first, i);
    run *= (A)fac[j];
  }
  if (kErt) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, d));
    rs.low_any = first < S;
    rs.first_low = rs.low_any ? first : S;
  } else {
    rs.low_any = false;
    rs.first_low = S;
  }
}

// cut = first sample index whose weight is forced to zero (S = none)
template <bool kFast = false>
__device__ __forceinline__ void ray_outputs(const RaySamples& rs, const float* __restrict__ raw_row,
                                            const float* __restrict__ z_row, int S, int per, int lane,
                                            int cut, int white_bkgd, size_t ray,
                                            float* __restrict__ rgb_map, float* __restrict__ disp_map,
                                            float* __restrict__ acc_map, float* __restrict__ depth_map,
                                            float* __restrict__ weights) {
  using A = typename AccT<kFast>::type;
  A sr = 0, sg = 0, sb = 0, sd = 0, sa = 0;
#pragma unroll
  for (int j = 0; j < kMaxPer; ++j) {
    int i = lane * per + j;
    if (j < per && i < S) {
      float w = __fmul_rn(rs.alpha[j], rs.T[j]);
      if (i >= cut) w = __fmul_rn(w, 0.f);  // weights * (~mask).float()
      if (w != 0.f) {   // w == +0: every product below is +0 exactly (sigmoid and z are finite), skip the three exps
        float4 r4 = *reinterpret_cast<const float4*>(raw_row + (size_t)i * 4);
        float cr = sigmoid_ref<kFast>(r4.x), cg = sigmoid_ref<kFast>(r4.y), cb = sigmoid_ref<kFast>(r4.z);
        sr += (A)__fmul_rn(w, cr);
        sg += (A)__fmul_rn(w, cg);
        sb += (A)__fmul_rn(w, cb);
        sd += (A)__fmul_rn(w, z_row[i]);
        sa += (A)w;
      }
      if (weights) weights[ray * S + i] = w;
    }
  }
  if constexpr (kFast) {
    // slots 0..4 = r, g, b, depth, acc; slot k's total ends up in lanes 4k..4k+3, which write their own output
    float v[8] = {(float)sr, (float)sg, (float)sb, (float)sd, (float)sa, 0.f, 0.f, 0.f};
    const float tot = warp_reduce8(v, lane);
    const float acc = __shfl_sync(0xffffffffu, tot, 16);
    if ((lane & 3) == 0) {
      const int slot = lane >> 2;
      if (slot < 3) {
        rgb_map[ray * 3 + slot] = white_bkgd ? tot + (1.f - acc) : tot;
      } else if (slot == 3) {
        float q = __fdiv_rn(tot, acc);
        float m = (q != q) ? q : fmaxf(1e-10f, q);  // torch.max propagates NaN (acc == 0)
        depth_map[ray] = tot;
        disp_map[ray] = __fdiv_rn(1.f, m);
      } else if (slot == 4) {
        acc_map[ray] = tot;
      }
    }
    return;
  }
  sr = warp_sum(sr); sg = warp_sum(sg); sb = warp_sum(sb); sd = warp_sum(sd); sa = warp_sum(sa);
  if (lane == 0) {
    float acc = (float)sa, depth = (float)sd;
    float r = (float)sr, g = (float)sg, b = (float)sb;
    float q = __fdiv_rn(depth, acc);
    float m = (q != q) ? q : fmaxf(1e-10f, q);  // torch.max propagates NaN (acc == 0)
    if (white_bkgd) {
      float bg = __fsub_rn(1.f, acc);
      r = __fadd_rn(r, bg); g = __fadd_rn(g, bg); b = __fadd_rn(b, bg);
    }
    rgb_map[ray * 3 + 0] = r; rgb_map[ray * 3 + 1] = g; rgb_map[ray * 3 + 2] = b;
    disp_map[ray] = __fdiv_rn(1.f, m);
    acc_map[ray] = acc;
    depth_map[ray] = depth;
  }
}

// PLAIN and ERT: one warp per ray
template <bool kErt, bool kFast>
__global__ void __launch_bounds__(kCompWarps * 32, kFast ? kMinBlocksFast : kMinBlocksExact)
composite_kernel(const float* __restrict__ raw, const float* __restrict__ z_vals,
                 const float* __restrict__ rays_d, int n_rays, int S, float thr, int white_bkgd,
                 float* __restrict__ rgb_map, float* __restrict__ disp_map,
                 float* __restrict__ acc_map, float* __restrict__ depth_map,
                 float* __restrict__ weights, const uint32_t* __restrict__ keep_bits,
                 const int32_t* __restrict__ ray_list, const int32_t* __restrict__ n_list,
                 uint8_t* __restrict__ low_flag = nullptr, int32_t* __restrict__ chunk_any = nullptr, int compat_chunk = 1) {
  const int lane = threadIdx.x & 31;
  const int per = (S + 31) / 32;
  // ray_list != NULL: persistent grid over the rays that survived the culling (common.cuh: RayList)
  const size_t n_eff = ray_list != nullptr ? (size_t)*n_list : (size_t)n_rays;
  for (size_t i = (size_t)blockIdx.x * kCompWarps + (threadIdx.x >> 5); i < n_eff; i += (size_t)gridDim.x * kCompWarps) {
    const size_t ray = ray_list != nullptr ? (size_t)ray_list[i] : i;
    const float* raw_row = raw + ray * S * 4;
    const float* z_row = z_vals + ray * S;
    RaySamples rs;
    ray_alpha_T<kErt, kFast>(raw_row, z_row, ray_norm(rays_d + ray * 3), S, per, lane, thr, rs, keep_bits, ray * S);
    if (kErt && low_flag != nullptr && lane == 0) {   // first pass of the two-kernel ERT_COMPAT path below
      low_flag[ray] = rs.low_any ? 1 : 0;
      if (rs.low_any) chunk_any[ray / (size_t)compat_chunk] = 1;   // same value from every writer
    }
    ray_outputs<kFast>(rs, raw_row, z_row, S, per, lane, rs.first_low, white_bkgd, ray, rgb_map, disp_map,
                       acc_map, depth_map, weights);
  }
}

// ERT_COMPAT in two launches (whole-pass driver; the single-kernel version below serialises a 2048-ray chunk in one
// thread block and made the reference's DEFAULT configuration -- lego.yaml: enable_ert = True -- 19 ms per frame
// slower than the plain compositor).  The quirk (:1115-1123): `if low.any()` is evaluated over the chunk; when it
// fires, first = argmax(low) is 0 for the rays that never go low, so those rays lose all their weights.  Pass 1 =
// composite_kernel<ERT> (every ray truncated at its own first low sample -- already the final answer for the rays that
// go low and for chunks where nobody does) + one flag per ray and per chunk; pass 2 redoes, with cut = 0, only the
// rays that never went low inside a chunk where somebody did.
template <bool kFast>
__global__ void __launch_bounds__(kCompWarps * 32)
composite_compat_fixup_kernel(const float* __restrict__ raw, const float* __restrict__ z_vals,
                              const float* __restrict__ rays_d, int n_rays, int S, float thr, int white_bkgd,
                              int compat_chunk, float* __restrict__ rgb_map, float* __restrict__ disp_map,
                              float* __restrict__ acc_map, float* __restrict__ depth_map,
                              float* __restrict__ weights, const uint8_t* __restrict__ low_flag,
                              const int32_t* __restrict__ chunk_any) {
  const int lane = threadIdx.x & 31;
  const size_t ray = (size_t)blockIdx.x * kCompWarps + (threadIdx.x >> 5);
  if (ray >= (size_t)n_rays) return;
  if (low_flag[ray] || !chunk_any[ray / (size_t)compat_chunk]) return;
  const int per = (S + 31) / 32;
  const float* raw_row = raw + ray * S * 4;
  const float* z_row = z_vals + ray * S;
  RaySamples rs;
  ray_alpha_T<true, kFast>(raw_row, z_row, ray_norm(rays_d + ray * 3), S, per, lane, thr, rs);
  ray_outputs<kFast>(rs, raw_row, z_row, S, per, lane, /*cut=*/0, white_bkgd, ray, rgb_map, disp_map, acc_map,
                     depth_map, weights);
}

// ERT_COMPAT: literal :1115-1123.  `if low.any()` is evaluated over the whole call (a
// 2048-ray chunk in the reference); when it fires, first = argmax(low) is 0 for rays that never
// go low, so those rays lose ALL weights.  One block per chunk, two passes.
__global__ void __launch_bounds__(1024)
composite_ert_compat_kernel(const float* __restrict__ raw, const float* __restrict__ z_vals,
                            const float* __restrict__ rays_d, int n_rays, int S, float thr,
                            int white_bkgd, int chunk, float* __restrict__ rgb_map,
                            float* __restrict__ disp_map, float* __restrict__ acc_map,
                            float* __restrict__ depth_map, float* __restrict__ weights) {
  __shared__ int s_any;
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  size_t begin = (size_t)blockIdx.x * chunk;
  size_t end = min(begin + (size_t)chunk, (size_t)n_rays);
  int per = (S + 31) / 32;
  if (threadIdx.x == 0) s_any = 0;
  __syncthreads();
  bool any = false;
  for (size_t ray = begin + warp; ray < end; ray += nwarps) {
    RaySamples rs;
    ray_alpha_T<true>(raw + ray * S * 4, z_vals + ray * S, ray_norm(rays_d + ray * 3), S, per, lane, thr, rs);
    any |= rs.low_any;
  }
  if (any && lane == 0) atomicOr(&s_any, 1);
  __syncthreads();
  bool chunk_any = s_any != 0;
  for (size_t ray = begin + warp; ray < end; ray += nwarps) {
    const float* raw_row = raw + ray * S * 4;
    const float* z_row = z_vals + ray * S;
    RaySamples rs;
    ray_alpha_T<true>(raw_row, z_row, ray_norm(rays_d + ray * 3), S, per, lane, thr, rs);
    int cut = chunk_any ? (rs.low_any ? rs.first_low : 0) : S;
    ray_outputs(rs, raw_row, z_row, S, per, lane, cut, white_bkgd, ray, rgb_map, disp_map, acc_map,
                depth_map, weights);
  }
}

// ------------------------------------------------------------------------------------------
// a7: backward of the PLAIN variant.  With G_i = dL/dw_i = g_rgb.c_i + g_acc' + g_depth z_i +
// g_w_i (g_acc' = g_acc - sum(g_rgb) under white_bkgd), f_i = 1-alpha_i+1e-10:
//   dL/dalpha_i = G_i T_i - (sum_{j>i} G_j w_j) / f_i
//   dalpha/dsigma_raw = dist * exp(-sigma*dist) * [sigma_raw > 0]     (not dist*(1-alpha): inf*0)
//   dL/drgb_raw_c = w_i g_rgb_c c (1-c)
// The suffix sum is a reverse warp scan in fp64.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCompWarps * 32, kMinBlocksExact)
composite_backward_kernel(const float* __restrict__ raw, const float* __restrict__ z_vals,
                          const float* __restrict__ rays_d, int n_rays, int S, int white_bkgd,
                          const float* __restrict__ g_rgb_map, const float* __restrict__ g_acc_map,
                          const float* __restrict__ g_depth_map, const float* __restrict__ g_weights,
                          float* __restrict__ g_raw, float* __restrict__ g_z) {
  int lane = threadIdx.x & 31;
  size_t ray = (size_t)blockIdx.x * kCompWarps + (threadIdx.x >> 5);
  if (ray >= (size_t)n_rays) return;
  int per = (S + 31) / 32;
  const float* raw_row = raw + ray * S * 4;
  const float* z_row = z_vals + ray * S;
  float dnorm = ray_norm(rays_d + ray * 3);
  RaySamples rs;
  ray_alpha_T<false>(raw_row, z_row, dnorm, S, per, lane, 0.f, rs);
  float gr = g_rgb_map ? g_rgb_map[ray * 3 + 0] : 0.f;
  float gg = g_rgb_map ? g_rgb_map[ray * 3 + 1] : 0.f;
  float gb = g_rgb_map ? g_rgb_map[ray * 3 + 2] : 0.f;
  float ga = g_acc_map ? g_acc_map[ray] : 0.f;
  float gd = g_depth_map ? g_depth_map[ray] : 0.f;
  if (white_bkgd) ga -= (gr + gg + gb);
  float G[kMaxPer], w[kMaxPer], c[kMaxPer][3];
  double local = 0.0;
#pragma unroll
  for (int j = 0; j < kMaxPer; ++j) {
    int i = lane * per + j;
    G[j] = 0.f; w[j] = 0.f;
    if (j < per && i < S) {
      float4 r4 = *reinterpret_cast<const float4*>(raw_row + (size_t)i * 4);
      c[j][0] = sigmoid_ref<false>(r4.x);
      c[j][1] = sigmoid_ref<false>(r4.y);
      c[j][2] = sigmoid_ref<false>(r4.z);
      w[j] = rs.alpha[j] * rs.T[j];
      G[j] = gr * c[j][0] + gg * c[j][1] + gb * c[j][2] + ga + gd * z_row[i] +
             (g_weights ? g_weights[ray * S + i] : 0.f);
      local += (double)G[j] * (double)w[j];
    }
  }
  // suffix sums: total - inclusive prefix
  double incl = local;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    double o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  double total = __shfl_sync(0xffffffffu, incl, 31);
  double after = total - incl;  // sum over samples owned by later lanes
  // walk own samples from last to first
  double suffix = after;
  float gdist[kMaxPer];   // dL/d(z[i+1] - z[i]) of the interval that starts at own sample j (g_z only)
  float gdist_last = 0.f; // ... of this lane's last interval, handed to the next lane
#pragma unroll
  for (int j = kMaxPer - 1; j >= 0; --j) {
    int i = lane * per + j;
    gdist[j] = 0.f;
    if (j < per && i < S) {
      float z0 = z_row[i];
      float dist = ((i + 1 < S) ? (z_row[i + 1] - z0) : 1e10f) * dnorm;
      float sraw = raw_row[(size_t)i * 4 + 3];
      float sig = fmaxf(sraw, 0.f);
      float f = (1.f - rs.alpha[j]) + 1e-10f;
      double dalpha = (double)G[j] * (double)rs.T[j] - suffix / (double)f;
      float e = (sraw > 0.f) ? expf(-sig * dist) : 0.f;
      float dsig = dist * e;
      float4 o;
      o.x = w[j] * gr * c[j][0] * (1.f - c[j][0]);
      o.y = w[j] * gg * c[j][1] * (1.f - c[j][1]);
      o.z = w[j] * gb * c[j][2] * (1.f - c[j][2]);
      o.w = (float)(dalpha * (double)dsig);
      *reinterpret_cast<float4*>(g_raw + (ray * S + i) * 4) = o;
      // dalpha/ddist = sigma exp(-sigma dist); the last interval (1e10) does not depend on z
      if (i + 1 < S) gdist[j] = (float)(dalpha * (double)(sig * e)) * dnorm;
      if (j == per - 1 || i == S - 1) gdist_last = gdist[j];
      suffix += (double)G[j] * (double)w[j];
    }
  }
  if (g_z != nullptr) {
    // z_i enters dist_i with -1 and dist_{i-1} with +1 (volume_renderer.py:295-297), and depth_map = sum w z (:339)
    float prev = __shfl_up_sync(0xffffffffu, gdist_last, 1);
    if (lane == 0) prev = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxPer; ++j) {
      int i = lane * per + j;
      if (j < per && i < S) {
        g_z[ray * S + i] = gd * w[j] - gdist[j] + prev;
        prev = gdist[j];
      }
    }
  }
}

// raw_noise_std (volume_renderer.py:310-314, :1099-1103): sigma_raw += randn * std before the relu.  Applied in
// place to the MLP output, so the compositing forward AND backward see the noisy density (d noisy / d raw = 1).
// Box-Muller on two counter-based uniforms keyed on (seed, row).
__global__ void sigma_noise_kernel(float* __restrict__ raw, long long n_rows, float std, uint64_t seed) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rows) return;
  const uint32_t lo = (uint32_t)idx, hi = (uint32_t)(idx >> 32);
  float u1 = 1.0f - uniform01(seed, lo, 2u * hi);         // (0,1]
  float u2 = uniform01(seed ^ 0xD1B54A32D192ED03ull, lo, 2u * hi + 1u);
  float g = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
  raw[idx * 4 + 3] += std * g;
}

}  // namespace nb

using namespace nb;

extern "C" int nerfb200_sigma_noise(float* raw, long long n_rows, float std, uint64_t seed, void* stream) {
  NB_CHECK_ARG(n_rows <= 0 || raw, "sigma_noise: null pointer");
  NB_CHECK_ARG(n_rows >= 0 && std >= 0.f, "sigma_noise: bad arguments (n_rows=%lld, std=%g)", n_rows, (double)std);
  if (n_rows == 0 || std == 0.f) return 0;
  sigma_noise_kernel<<<ceil_div(n_rows, 256), 256, 0, (cudaStream_t)stream>>>(raw, n_rows, std, seed);
  NB_LAUNCH_OK("sigma_noise_kernel");
  return 0;
}

extern "C" int nerfb200_composite_forward(const float* raw, const float* z_vals, const float* rays_d,
                                          int n_rays, int n_samples, int variant, float ert_threshold,
                                          int white_bkgd, int compat_chunk, float* rgb_map,
                                          float* disp_map, float* acc_map, float* depth_map,
                                          float* weights, void* stream) {
  return nerfb200_composite_forward_masked(raw, z_vals, rays_d, nullptr, n_rays, n_samples, variant, ert_threshold,
                                           white_bkgd, compat_chunk, rgb_map, disp_map, acc_map, depth_map, weights,
                                           stream);
}

extern "C" int nerfb200_composite_forward_masked(const float* raw, const float* z_vals, const float* rays_d,
                                                 const uint32_t* keep_bits, int n_rays, int n_samples, int variant,
                                                 float ert_threshold, int white_bkgd, int compat_chunk,
                                                 float* rgb_map, float* disp_map, float* acc_map, float* depth_map,
                                                 float* weights, void* stream) {
  return nb::composite_forward_culled(raw, z_vals, rays_d, keep_bits, RayList{nullptr, nullptr}, n_rays, n_samples, variant, ert_threshold,
                                      white_bkgd, compat_chunk, rgb_map, disp_map, acc_map, depth_map, weights, stream);
}

int nb::composite_forward_compat2(const float* raw, const float* z_vals, const float* rays_d, int n_rays, int n_samples,
                                  int fast, float ert_threshold, int white_bkgd, int compat_chunk, float* rgb_map,
                                  float* disp_map, float* acc_map, float* depth_map, float* weights,
                                  uint8_t* low_flag, int32_t* chunk_any, void* stream) {
  NB_CHECK_ARG(n_rays <= 0 || (raw && z_vals && rays_d && rgb_map && disp_map && acc_map && depth_map && low_flag && chunk_any),
               "composite_forward (compat): null pointer");
  NB_CHECK_ARG(n_samples >= 1 && n_samples <= 32 * kMaxPer && compat_chunk > 0 && n_rays >= 0, "composite_forward (compat): bad sizes");
  if (n_rays == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = ceil_div(n_rays, kCompWarps);
  NB_CUDA(cudaMemsetAsync(chunk_any, 0, (size_t)ceil_div(n_rays, compat_chunk) * sizeof(int32_t), st));
  if (fast) {
    composite_kernel<true, true><<<blocks, kCompWarps * 32, 0, st>>>(raw, z_vals, rays_d, n_rays, n_samples, ert_threshold,
        white_bkgd, rgb_map, disp_map, acc_map, depth_map, weights, nullptr, nullptr, nullptr, low_flag, chunk_any, compat_chunk);
    NB_LAUNCH_OK("composite_kernel");
    composite_compat_fixup_kernel<true><<<blocks, kCompWarps * 32, 0, st>>>(raw, z_vals, rays_d, n_rays, n_samples,
        ert_threshold, white_bkgd, compat_chunk, rgb_map, disp_map, acc_map, depth_map, weights, low_flag, chunk_any);
  } else {
    composite_kernel<true, false><<<blocks, kCompWarps * 32, 0, st>>>(raw, z_vals, rays_d, n_rays, n_samples, ert_threshold,
        white_bkgd, rgb_map, disp_map, acc_map, depth_map, weights, nullptr, nullptr, nullptr, low_flag, chunk_any, compat_chunk);
    NB_LAUNCH_OK("composite_kernel");
    composite_compat_fixup_kernel<false><<<blocks, kCompWarps * 32, 0, st>>>(raw, z_vals, rays_d, n_rays, n_samples,
        ert_threshold, white_bkgd, compat_chunk, rgb_map, disp_map, acc_map, depth_map, weights, low_flag, chunk_any);
  }
  NB_LAUNCH_OK("composite_compat_fixup_kernel");
  return 0;
}

// + ray list (common.cuh: RayList); internal to the whole-pass driver
int nb::composite_forward_culled(const float* raw, const float* z_vals, const float* rays_d, const uint32_t* keep_bits,
                                 RayList rl, int n_rays, int n_samples, int variant, float ert_threshold,
                                 int white_bkgd, int compat_chunk, float* rgb_map, float* disp_map, float* acc_map,
                                 float* depth_map, float* weights, void* stream) {
  NB_CHECK_ARG(n_rays <= 0 || (raw && z_vals && rays_d && rgb_map && disp_map && acc_map && depth_map),
               "composite_forward: null pointer");
  NB_CHECK_ARG(n_samples >= 1 && n_samples <= 32 * kMaxPer, "composite_forward: n_samples=%d out of range [1,%d]",
               n_samples, 32 * kMaxPer);
  NB_CHECK_ARG(n_rays >= 0, "composite_forward: negative n_rays");
  const bool fast = (variant & NERFB200_COMPOSITE_FAST_MATH) != 0;
  variant &= ~NERFB200_COMPOSITE_FAST_MATH;
  NB_CHECK_ARG(variant >= 0 && variant <= 2, "composite_forward: unknown variant %d", variant);
  if (n_rays == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = ceil_div(n_rays, kCompWarps);
  if (rl.rays != nullptr && blocks > kPersistentBlocks) blocks = kPersistentBlocks;
#define NB_COMPOSITE(ERT, FAST, THR)                                                                              \
  composite_kernel<ERT, FAST><<<blocks, kCompWarps * 32, 0, st>>>(raw, z_vals, rays_d, n_rays, n_samples, THR, white_bkgd, \
                                                                  rgb_map, disp_map, acc_map, depth_map, weights, keep_bits, rl.rays, rl.count)
  if (variant == NERFB200_COMPOSITE_PLAIN) {
    if (fast) NB_COMPOSITE(false, true, 0.f); else NB_COMPOSITE(false, false, 0.f);
  } else if (variant == NERFB200_COMPOSITE_ERT) {
    if (fast) NB_COMPOSITE(true, true, ert_threshold); else NB_COMPOSITE(true, false, ert_threshold);
  } else {
    NB_CHECK_ARG(compat_chunk > 0, "composite_forward: compat_chunk must be > 0");
    NB_CHECK_ARG(keep_bits == nullptr && rl.rays == nullptr, "composite_forward: the ERT_COMPAT variant has no masked form");
    composite_ert_compat_kernel<<<ceil_div(n_rays, compat_chunk), 1024, 0, st>>>(
        raw, z_vals, rays_d, n_rays, n_samples, ert_threshold, white_bkgd, compat_chunk, rgb_map, disp_map, acc_map,
        depth_map, weights);
  }
  NB_LAUNCH_OK("composite_kernel");
  return 0;
}

extern "C" int nerfb200_composite_backward_z(const float* raw, const float* z_vals, const float* rays_d,
                                             int n_rays, int n_samples, int white_bkgd,
                                             const float* g_rgb_map, const float* g_acc_map,
                                             const float* g_depth_map, const float* g_weights,
                                             float* g_raw, float* g_z, void* stream) {
  NB_CHECK_ARG(n_rays <= 0 || (raw && z_vals && rays_d && g_raw), "composite_backward: null pointer");
  NB_CHECK_ARG(n_samples >= 1 && n_samples <= 32 * kMaxPer, "composite_backward: n_samples=%d out of range", n_samples);
  NB_CHECK_ARG(n_rays >= 0, "composite_backward: negative n_rays");
  if (n_rays == 0) return 0;
  composite_backward_kernel<<<ceil_div(n_rays, kCompWarps), kCompWarps * 32, 0, (cudaStream_t)stream>>>(
      raw, z_vals, rays_d, n_rays, n_samples, white_bkgd, g_rgb_map, g_acc_map, g_depth_map, g_weights, g_raw, g_z);
  NB_LAUNCH_OK("composite_backward_kernel");
  return 0;
}

extern "C" int nerfb200_composite_backward(const float* raw, const float* z_vals, const float* rays_d,
                                           int n_rays, int n_samples, int white_bkgd,
                                           const float* g_rgb_map, const float* g_acc_map,
                                           const float* g_depth_map, const float* g_weights,
                                           float* g_raw, void* stream) {
  return nerfb200_composite_backward_z(raw, z_vals, rays_d, n_rays, n_samples, white_bkgd, g_rgb_map, g_acc_map,
                                       g_depth_map, g_weights, g_raw, nullptr, stream);
}
