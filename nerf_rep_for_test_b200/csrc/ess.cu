// a8 occupancy-grid empty-space skipping.
// Reference: volume_renderer.py:830-873 (grid), :992-1007 (_is_empty_space), :1009-1087
// (_sample_coarse_with_ess), :963-985 (_update_occupancy_grid).
//
// Intended per-ray semantics of :1037-1077 (the reference writes through a stride-0 expand()ed
// view so all rays of a chunk alias one row -- SURVEY 8a8; that aliasing is NOT reproduced, the
// oracle restates both).  Byte/index work, HBM/L2-bound: 64 grid lookups (1 B each, 2 MB grid
// stays in L2) + 256 B of z per ray.
#include "common.cuh"

namespace nb {

constexpr int kEssWarps = 8;
constexpr int kEssMaxS = 256;

__device__ __forceinline__ int grid_index(float p, int res) {
  // :996-1000: ((p - min) / (max - min)).clamp(0,1) * (res-1) -> long (truncation) -> clamp
  // (max - min) = 4: scaling by a power of two is exact, so the multiply rounds exactly like the division
  float n = __fmul_rn(__fsub_rn(p, -2.0f), 0.25f);
  n = fminf(fmaxf(n, 0.f), 1.f);
  int c = (int)__fmul_rn(n, (float)(res - 1));
  return min(max(c, 0), res - 1);
}

__global__ void __launch_bounds__(kEssWarps * 32)
ess_resample_kernel(const uint8_t* __restrict__ grid, int res, const float* __restrict__ rays_o,
                    const float* __restrict__ rays_d, int n_rays, int S, float* __restrict__ z_vals,
                    int32_t* __restrict__ n_empty_out) {
  __shared__ float s_keep[kEssWarps][kEssMaxS];
  __shared__ float s_z[kEssWarps][kEssMaxS];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int ray = blockIdx.x * kEssWarps + warp;
  if (ray >= n_rays) return;
  float o[3], d[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) { o[c] = rays_o[(size_t)ray * 3 + c]; d[c] = rays_d[(size_t)ray * 3 + c]; }
  float* zrow = z_vals + (size_t)ray * S;
  float* keep = s_keep[warp];
  float* zs = s_z[warp];
  // pass 1: occupancy of every sample, ordered compaction of the occupied z's
  int n_keep = 0;
  for (int base = 0; base < S; base += 32) {
    int i = base + lane;
    bool occ = false;
    float z = 0.f;
    if (i < S) {
      z = zrow[i];
      zs[i] = z;
      int gx = grid_index(__fadd_rn(o[0], __fmul_rn(d[0], z)), res);
      int gy = grid_index(__fadd_rn(o[1], __fmul_rn(d[1], z)), res);
      int gz = grid_index(__fadd_rn(o[2], __fmul_rn(d[2], z)), res);
      occ = grid[((size_t)gx * res + gy) * res + gz] != 0;
    }
    unsigned m = __ballot_sync(0xffffffffu, occ);
    if (occ) keep[n_keep + __popc(m & ((1u << lane) - 1))] = z;
    n_keep += __popc(m);
  }
  __syncwarp();
  int n_empty = S - n_keep;
  if (n_empty_out && lane == 0) n_empty_out[ray] = n_empty;
  // empty_ratios > 0.5 (:1031-1033), and at least one occupied sample (:1045)
  if (!((float)n_empty / (float)S > 0.5f) || n_keep == 0) return;
  int n_add = S - n_keep;
  float lo = keep[0], hi = keep[n_keep - 1];  // occupied z's are ascending
  float step = n_add > 1 ? __fdiv_rn(__fsub_rn(hi, lo), (float)(n_add - 1)) : 0.f;
  // merged position of kept[i] = i + #added < kept[i]; of added[j] = j + #kept <= added[j]
  auto added = [&](int j) {  // torch.linspace(lo, hi, n_add)
    if (n_add == 1) return lo;
    return (j < n_add / 2) ? __fadd_rn(lo, __fmul_rn(step, (float)j))
                           : __fsub_rn(hi, __fmul_rn(step, (float)(n_add - 1 - j)));
  };
  for (int i = lane; i < n_keep; i += 32) {
    float x = keep[i];
    int l = 0, h = n_add;
    while (l < h) { int mid = (l + h) >> 1; if (added(mid) < x) l = mid + 1; else h = mid; }
    zrow[i + l] = x;
  }
  for (int j = lane; j < n_add; j += 32) {
    float x = added(j);
    int l = 0, h = n_keep;
    while (l < h) { int mid = (l + h) >> 1; if (keep[mid] <= x) l = mid + 1; else h = mid; }
    zrow[j + l] = x;
  }
}

// ---- the reference's LITERAL ESS (volume_renderer.py:1009-1077): z_vals is an expand()ed stride-0 view (:1020), so
// `z_vals[i] = combined_z_vals` (:1077) writes the ONE row every ray of the call shares.  Each highly-empty ray i, in
// order, reads the current shared row, keeps the entries where ITS occupancy mask -- evaluated up front at the
// unmodified linspace depths (:1024-1028) -- says "occupied", refills with linspace(min, max, n_add), sorts, and writes
// the row back; all rays of the call (a 2048-ray chunk, :147) end up with the row the last such ray left.
// Pass 1 (parallel): per ray a 64-bit EMPTY mask at the table depths (bit 63 of word... see below) + the flag.
// Pass 2 (one warp per chunk, sequential over its flagged rays): the row lives in shared memory.
// S <= 64 (the mask is one 64-bit word; the reference's N_samples is 64).
__global__ void __launch_bounds__(256)
ess_masks_kernel(const uint8_t* __restrict__ grid, int res, const float* __restrict__ rays_o,
                 const float* __restrict__ rays_d, const float* __restrict__ z_table, int n_rays, int S,
                 unsigned long long* __restrict__ masks) {
  int ray = blockIdx.x * blockDim.x + threadIdx.x;
  if (ray >= n_rays) return;
  float o[3], d[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) { o[c] = rays_o[(size_t)ray * 3 + c]; d[c] = rays_d[(size_t)ray * 3 + c]; }
  unsigned long long empty = 0ull;
  for (int i = 0; i < S; ++i) {
    const float z = z_table[i];
    int gx = grid_index(__fadd_rn(o[0], __fmul_rn(d[0], z)), res);
    int gy = grid_index(__fadd_rn(o[1], __fmul_rn(d[1], z)), res);
    int gz = grid_index(__fadd_rn(o[2], __fmul_rn(d[2], z)), res);
    if (grid[((size_t)gx * res + gy) * res + gz] == 0) empty |= 1ull << i;
  }
  masks[ray] = empty;
}

__global__ void __launch_bounds__(32)
ess_resample_compat_kernel(const unsigned long long* __restrict__ masks, const float* __restrict__ z_table, int n_rays,
                           int S, int chunk, float* __restrict__ z_vals) {
  __shared__ float row[64], keep[64], next[64];
  const int lane = threadIdx.x;
  const int r0 = blockIdx.x * chunk;
  const int r1 = min(n_rays, r0 + chunk);
  for (int i = lane; i < S; i += 32) row[i] = z_table[i];
  __syncwarp();
  const unsigned long long valid = S >= 64 ? ~0ull : ((1ull << S) - 1ull);
  for (int base = r0; base < r1; base += 32) {
    // 32 rays' masks at a time; the flagged ones are then processed strictly in ray order
    const int ray = base + lane;
    unsigned long long m = ray < r1 ? masks[ray] : 0ull;
    const bool flag = ray < r1 && ((float)__popcll(m & valid) / (float)S > 0.5f) && (~m & valid) != 0ull;   // :1031-1033, :1045
    unsigned todo = __ballot_sync(0xffffffffu, flag);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const unsigned long long em = __shfl_sync(0xffffffffu, m, src);
      const unsigned long long occ = ~em & valid;
      const int n_keep = __popcll(occ), n_add = S - n_keep;
      // ordered compaction of the occupied entries of the CURRENT row
      for (int i = lane; i < S; i += 32)
        if ((occ >> i) & 1ull) keep[__popcll(occ & ((1ull << i) - 1ull))] = row[i];
      __syncwarp();
      const float lo = keep[0], hi = keep[n_keep - 1];      // the row is ascending
      const float step = n_add > 1 ? __fdiv_rn(__fsub_rn(hi, lo), (float)(n_add - 1)) : 0.f;
      auto added = [&](int j) {   // torch.linspace(lo, hi, n_add), as in ess_resample_kernel
        if (n_add == 1) return lo;
        return (j < n_add / 2) ? __fadd_rn(lo, __fmul_rn(step, (float)j)) : __fsub_rn(hi, __fmul_rn(step, (float)(n_add - 1 - j)));
      };
      for (int i = lane; i < n_keep; i += 32) {
        const float x = keep[i];
        int l = 0, h = n_add;
        while (l < h) { int mid = (l + h) >> 1; if (added(mid) < x) l = mid + 1; else h = mid; }
        next[i + l] = x;
      }
      for (int j = lane; j < n_add; j += 32) {
        const float x = added(j);
        int l = 0, h = n_keep;
        while (l < h) { int mid = (l + h) >> 1; if (keep[mid] <= x) l = mid + 1; else h = mid; }
        next[j + l] = x;
      }
      __syncwarp();
      for (int i = lane; i < S; i += 32) row[i] = next[i];
      __syncwarp();
    }
  }
  // every ray of the chunk gets the shared row
  for (long long e = lane; e < (long long)(r1 - r0) * S; e += 32) z_vals[(size_t)r0 * S + e] = row[e % S];
}

// stratified jitter of ALREADY PLACED per-ray depths (volume_renderer.py:1079-1085: the reference jitters after the ESS
// resampling, from each row's own mid-points): z_i <- lower_i + (upper_i - lower_i) u,  u keyed on (seed, ray, i)
constexpr int kJitWarps = 8;
__global__ void __launch_bounds__(kJitWarps * 32)
jitter_rows_kernel(float* __restrict__ z_vals, int n_rays, int S, uint64_t seed) {
  const int lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kJitWarps + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  float* z = z_vals + (size_t)ray * S;
  float out[kEssMaxS / 32];
#pragma unroll
  for (int k = 0; k < kEssMaxS / 32; ++k) {
    const int i = k * 32 + lane;
    out[k] = 0.f;
    if (i < S) {
      const float zi = z[i];
      const float lower = i == 0 ? zi : __fmul_rn(0.5f, __fadd_rn(zi, z[i - 1]));
      const float upper = i == S - 1 ? zi : __fmul_rn(0.5f, __fadd_rn(z[i + 1], zi));
      out[k] = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), uniform01(seed, (uint32_t)ray, (uint32_t)i)));
    }
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < kEssMaxS / 32; ++k) {
    const int i = k * 32 + lane;
    if (i < S) z[i] = out[k];
  }
}

__global__ void ess_update_kernel(uint8_t* __restrict__ grid, int res, const float* __restrict__ rays_o,
                                  const float* __restrict__ rays_d, const float* __restrict__ z_vals,
                                  const float* __restrict__ raw, const float* __restrict__ weights,
                                  long long total, int S, int use_origin) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  if (!(weights[idx] > 1e-4f)) return;                 // :1149
  if (!(fmaxf(raw[idx * 4 + 3], 0.f) > 0.01f)) return;  // :976-977
  long long ray = idx / S;
  float z = z_vals[idx];
  int g[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float p = __fmul_rn(rays_d[ray * 3 + c], z);       // :1151 -- origin omitted in the reference
    if (use_origin) p = __fadd_rn(rays_o[ray * 3 + c], p);
    g[c] = grid_index(p, res);
  }
  grid[((size_t)g[0] * res + g[1]) * res + g[2]] = 1;
}

// ---- empty-space skipping proper: stream compaction of the rows worth evaluating ----------------
// A block handles 32 words of 32 consecutive rows (each warp four of them, so four independent z / grid loads
// are in flight per thread), counts its survivors with ballots + one 32-entry warp scan, and reserves its run of
// row_ids with ONE atomicAdd: the first version did one atomic per warp on the single counter and was bound by
// same-address atomic throughput (1.1 ms per 800x800 frame for 164 M rows, 0.55 rows per clock and SM).
// keep_bits (optional): the ballot words themselves, bit m = row m kept -- read by the masked compositor.
constexpr int kCompactThreads = 256;
constexpr int kCompactWords = 4;                                      // 32-row words per warp
constexpr int kCompactRows = kCompactThreads * kCompactWords;         // 1024 rows per block
static_assert(kCompactThreads / 32 * kCompactWords == 32, "the per-block scan is one warp wide");

__global__ void __launch_bounds__(kCompactThreads)
ess_compact_kernel(const uint8_t* __restrict__ grid, int res, const float* __restrict__ rays_o,
                   const float* __restrict__ rays_d, const float* __restrict__ z_vals,
                   const float* __restrict__ z_term, long long total, int S,
                   int32_t* __restrict__ row_ids, int32_t* __restrict__ n_active,
                   uint32_t* __restrict__ keep_bits, const int32_t* __restrict__ ray_list,
                   const int32_t* __restrict__ n_list) {
  __shared__ int s_off[32];
  __shared__ int s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ray_list != NULL (S % 32 == 0, checked by the launcher): persistent grid over the samples of the listed rays;
  // position idx of that sequence is sample idx % S of ray ray_list[idx / S], and a 32-sample word of the
  // sequence is a 32-row word of the full row space
  const long long total_eff = ray_list != nullptr ? (long long)*n_list * S : total;
  for (long long blk0 = (long long)blockIdx.x * kCompactRows; blk0 < total_eff; blk0 += (long long)gridDim.x * kCompactRows) {
    const long long pos0 = blk0 + (long long)warp * (kCompactWords * 32) + lane;
    unsigned m[kCompactWords];
    int row[kCompactWords];
#pragma unroll
    for (int j = 0; j < kCompactWords; ++j) {
      const long long idx = pos0 + j * 32;
      bool keep = false;
      row[j] = 0;
      if (idx < total_eff) {
        const unsigned li = (unsigned)idx / (unsigned)S;   // total < 2^31 (checked by the entry point): 32-bit division
        const unsigned ray = ray_list != nullptr ? (unsigned)ray_list[li] : li;
        row[j] = (int)(ray * (unsigned)S + ((unsigned)idx - li * (unsigned)S));
        float z = z_vals[row[j]];
        int g[3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
          g[c] = grid_index(__fadd_rn(rays_o[ray * 3u + c], __fmul_rn(rays_d[ray * 3u + c], z)), res);
        keep = grid[((size_t)g[0] * res + g[1]) * res + g[2]] != 0;
        if (keep && z_term != nullptr) keep = z <= z_term[ray];
      }
      m[j] = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) {
        s_off[warp * kCompactWords + j] = __popc(m[j]);
        if (keep_bits != nullptr && idx < total_eff) keep_bits[row[j] >> 5] = m[j];
      }
    }
    __syncthreads();
    if (warp == 0) {   // exclusive scan of the 32 word counts, one reservation for the block
      const int c = s_off[lane];
      int incl = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      s_off[lane] = incl - c;
      if (lane == 31) s_base = incl ? atomicAdd(n_active, incl) : 0;
    }
    __syncthreads();
    const int base = s_base;
#pragma unroll
    for (int j = 0; j < kCompactWords; ++j)
      if ((m[j] >> lane) & 1u)
        row_ids[base + s_off[warp * kCompactWords + j] + __popc(m[j] & ((1u << lane) - 1))] = row[j];
    __syncthreads();   // s_off / s_base are rewritten by the next round
  }
}

__global__ void ert_depth_kernel(const float* __restrict__ weights, const float* __restrict__ z_vals, int n_rays,
                                 int S, float thr, float* __restrict__ z_term, const int32_t* __restrict__ ray_list,
                                 const int32_t* __restrict__ n_list) {
  const int n_eff = ray_list != nullptr ? *n_list : n_rays;
  for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < n_eff; it += gridDim.x * blockDim.x) {
  const int ray = ray_list != nullptr ? ray_list[it] : it;
  float acc = 0.f, zt = __int_as_float(0x7f800000);
  for (int i = 0; i < S; ++i) {
    if (1.f - acc < thr) { zt = z_vals[(size_t)ray * S + i]; break; }
    acc += weights[(size_t)ray * S + i];
  }
  z_term[ray] = zt;
  }
}

// Ray-level culling for the skipping mode: slab test of the segment o + d*z, z in [z_near, z_far], against the
// axis-aligned box of the occupied cells.  The box comes from the caller (Renderer: min / max occupied cell index per
// axis, widened by a margin, open-ended on a side whose boundary cell is occupied because grid_index clamps), so the
// test is conservative with respect to the per-sample lookup: a culled ray has no sample in an occupied cell, at any
// z of the coarse or the fine pass.  On a scene that fills a fraction of the frame most rays stop here and never reach
// the compaction, sample_pdf or the compositor's per-sample work.
__global__ void __launch_bounds__(256)
ray_cull_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, int n_rays,
                const float* __restrict__ z_table, int S, float lx, float ly, float lz, float hx, float hy,
                float hz, uint8_t* __restrict__ flags, int32_t* __restrict__ list, int32_t* __restrict__ count,
                nerfb200_maps mc, nerfb200_maps mf, int white_bkgd) {
  __shared__ int s_cnt[8];
  __shared__ int s_base;
  const int ray = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  bool hit = false;
  if (ray < n_rays) {
    const float lo[3] = {lx, ly, lz}, hi[3] = {hx, hy, hz};
    float t0 = z_table[0], t1 = z_table[S - 1];
    if (t0 > t1) { float t = t0; t0 = t1; t1 = t; }
    const float pad = 1e-3f * fmaxf(1.f, fabsf(t1));   // rounding of o + d*z in the lookups is ~1e-6
    t0 -= pad; t1 += pad;
    hit = true;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float o = rays_o[(size_t)ray * 3 + c], d = rays_d[(size_t)ray * 3 + c];
      if (!(lo[c] <= hi[c])) hit = false;   // empty box (no occupied cell)
      if (fabsf(d) < 1e-12f) {
        hit = hit && (o >= lo[c]) && (o <= hi[c]);
      } else {
        float ta = (lo[c] - o) / d, tb = (hi[c] - o) / d;   // +-inf for an open side
        if (ta > tb) { float t = ta; ta = tb; tb = t; }
        t0 = fmaxf(t0, ta);
        t1 = fminf(t1, tb);
      }
    }
    hit = hit && t0 <= t1;
    if (flags != nullptr) flags[ray] = hit ? 1 : 0;
    if (!hit && list != nullptr) {
      // every sample of a culled ray would be skipped and all its weights are 0: these are exactly the maps the
      // compositor produces for such a ray (0/0 -> NaN disparity included)
      const float bg = white_bkgd ? 1.f : 0.f;
      const float nan_disp = __fdiv_rn(1.f, __fdiv_rn(0.f, 0.f));
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const nerfb200_maps& mp = pass ? mf : mc;
        if (mp.rgb) { mp.rgb[(size_t)ray * 3 + 0] = bg; mp.rgb[(size_t)ray * 3 + 1] = bg; mp.rgb[(size_t)ray * 3 + 2] = bg; }
        if (mp.disp) mp.disp[ray] = nan_disp;
        if (mp.acc) mp.acc[ray] = 0.f;
        if (mp.depth) mp.depth[ray] = 0.f;
      }
    }
  }
  if (list == nullptr) return;
  // append the surviving rays: one atomic per block, ray order kept inside the block
  const unsigned m = __ballot_sync(0xffffffffu, hit);
  if (lane == 0) s_cnt[warp] = __popc(m);
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) { const int c = s_cnt[w]; s_cnt[w] = t; t += c; }
    s_base = t ? atomicAdd(count, t) : 0;
  }
  __syncthreads();
  if (hit) list[s_base + s_cnt[warp] + __popc(m & ((1u << lane) - 1))] = ray;
}

__global__ void accumulate_counts_kernel(const int32_t* __restrict__ counts, long long* __restrict__ totals) {
  if (threadIdx.x < 2) totals[threadIdx.x] += counts[threadIdx.x];
}

}  // namespace nb

using namespace nb;

extern "C" int nerfb200_ess_compact(const uint8_t* grid, int res, const float* rays_o, const float* rays_d,
                                    const float* z_vals, const float* z_term, int n_rays, int n_samples,
                                    int32_t* row_ids, int32_t* n_active, uint32_t* keep_bits, void* stream) {
  return ess_compact_culled(grid, res, rays_o, rays_d, z_vals, z_term, RayList{nullptr, nullptr}, n_rays, n_samples, row_ids,
                            n_active, keep_bits, stream);
}

int nb::ess_compact_culled(const uint8_t* grid, int res, const float* rays_o, const float* rays_d, const float* z_vals,
                           const float* z_term, RayList rl, int n_rays, int n_samples, int32_t* row_ids,
                           int32_t* n_active, uint32_t* keep_bits, void* stream) {
  NB_CHECK_ARG(rl.rays == nullptr || n_samples % 32 == 0, "ess_compact: a ray list needs n_samples %% 32 == 0 (got %d)", n_samples);
  NB_CHECK_ARG(n_active, "ess_compact: null counter");
  NB_CHECK_ARG(n_rays <= 0 || (grid && rays_o && rays_d && z_vals && row_ids), "ess_compact: null pointer");
  NB_CHECK_ARG(res >= 1 && res <= 1024, "ess_compact: bad grid resolution %d", res);
  NB_CHECK_ARG(n_rays >= 0 && n_samples >= 1 && (long long)n_rays * n_samples < (1LL << 31), "ess_compact: bad sizes");
  NB_CUDA(cudaMemsetAsync(n_active, 0, sizeof(int32_t), (cudaStream_t)stream));
  if (n_rays == 0) return 0;
  long long total = (long long)n_rays * n_samples;
  int blocks = ceil_div(total, kCompactRows);
  if (rl.rays != nullptr && blocks > kPersistentBlocks) blocks = kPersistentBlocks;
  ess_compact_kernel<<<blocks, kCompactThreads, 0, (cudaStream_t)stream>>>(
      grid, res, rays_o, rays_d, z_vals, z_term, total, n_samples, row_ids, n_active, keep_bits, rl.rays, rl.count);
  NB_LAUNCH_OK("ess_compact_kernel");
  return 0;
}

extern "C" int nerfb200_ert_depth(const float* weights, const float* z_vals, int n_rays, int n_samples, float thr,
                                  float* z_term, void* stream) {
  return ert_depth_culled(weights, z_vals, RayList{nullptr, nullptr}, n_rays, n_samples, thr, z_term, stream);
}

extern "C" int nerfb200_ray_cull(const float* rays_o, const float* rays_d, int n_rays, const float* z_table, int n_samples,
                                 const float* box_lo, const float* box_hi, uint8_t* ray_active, void* stream) {
  NB_CHECK_ARG(n_rays <= 0 || ray_active, "ray_cull: null pointer");
  return ray_cull_list(rays_o, rays_d, n_rays, z_table, n_samples, box_lo, box_hi, ray_active, nullptr, nullptr, nullptr,
                       nullptr, 0, stream);
}

int nb::ray_cull_list(const float* rays_o, const float* rays_d, int n_rays, const float* z_table, int n_samples,
                      const float* box_lo, const float* box_hi, uint8_t* flags, int32_t* list, int32_t* count,
                      const nerfb200_maps* maps_c, const nerfb200_maps* maps_f, int white_bkgd, void* stream) {
  NB_CHECK_ARG(n_rays <= 0 || (rays_o && rays_d && z_table), "ray_cull: null pointer");
  NB_CHECK_ARG(box_lo && box_hi, "ray_cull: null box");
  NB_CHECK_ARG(n_rays >= 0 && n_samples >= 1, "ray_cull: bad sizes");
  NB_CHECK_ARG((list == nullptr) == (count == nullptr), "ray_cull: list and count go together");
  if (n_rays == 0) return 0;
  const nerfb200_maps none = {nullptr, nullptr, nullptr, nullptr};
  ray_cull_kernel<<<ceil_div(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(
      rays_o, rays_d, n_rays, z_table, n_samples, box_lo[0], box_lo[1], box_lo[2], box_hi[0], box_hi[1], box_hi[2], flags,
      list, count, maps_c ? *maps_c : none, maps_f ? *maps_f : none, white_bkgd);
  NB_LAUNCH_OK("ray_cull_kernel");
  return 0;
}

int nb::ert_depth_culled(const float* weights, const float* z_vals, RayList rl, int n_rays, int n_samples,
                         float thr, float* z_term, void* stream) {
  NB_CHECK_ARG(n_rays <= 0 || (weights && z_vals && z_term), "ert_depth: null pointer");
  NB_CHECK_ARG(n_rays >= 0 && n_samples >= 1, "ert_depth: bad sizes");
  if (n_rays == 0) return 0;
  int blocks = ceil_div(n_rays, 128);
  if (rl.rays != nullptr && blocks > kPersistentBlocks) blocks = kPersistentBlocks;
  ert_depth_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(weights, z_vals, n_rays, n_samples, thr, z_term, rl.rays, rl.count);
  NB_LAUNCH_OK("ert_depth_kernel");
  return 0;
}

extern "C" int nerfb200_accumulate_counts(const int32_t* counts, int64_t* totals, void* stream) {
  NB_CHECK_ARG(counts && totals, "accumulate_counts: null pointer");
  accumulate_counts_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(counts, reinterpret_cast<long long*>(totals));
  NB_LAUNCH_OK("accumulate_counts_kernel");
  return 0;
}

extern "C" int nerfb200_ess_resample(const uint8_t* grid, int res, const float* rays_o, const float* rays_d,
                                     int n_rays, int n_samples, float* z_vals, int32_t* n_empty, void* stream) {
  NB_CHECK_ARG(n_rays <= 0 || (grid && rays_o && rays_d && z_vals), "ess_resample: null pointer");
  NB_CHECK_ARG(res >= 1 && res <= 1024, "ess_resample: bad grid resolution %d", res);
  NB_CHECK_ARG(n_samples >= 1 && n_samples <= kEssMaxS, "ess_resample: n_samples=%d out of range", n_samples);
  NB_CHECK_ARG(n_rays >= 0, "ess_resample: negative n_rays");
  if (n_rays == 0) return 0;
  ess_resample_kernel<<<ceil_div(n_rays, kEssWarps), kEssWarps * 32, 0, (cudaStream_t)stream>>>(
      grid, res, rays_o, rays_d, n_rays, n_samples, z_vals, n_empty);
  NB_LAUNCH_OK("ess_resample_kernel");
  return 0;
}

extern "C" int nerfb200_ess_update(uint8_t* grid, int res, const float* rays_o, const float* rays_d,
                                   const float* z_vals, const float* raw, const float* weights, int n_rays,
                                   int n_samples, int use_origin, void* stream) {
  NB_CHECK_ARG(n_rays <= 0 || (grid && rays_d && z_vals && raw && weights), "ess_update: null pointer");
  NB_CHECK_ARG(!use_origin || rays_o, "ess_update: use_origin needs rays_o");
  NB_CHECK_ARG(res >= 1 && res <= 1024, "ess_update: bad grid resolution %d", res);
  NB_CHECK_ARG(n_rays >= 0 && n_samples >= 1, "ess_update: bad sizes");
  if (n_rays == 0) return 0;
  long long total = (long long)n_rays * n_samples;
  ess_update_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(grid, res, rays_o, rays_d, z_vals, raw,
                                                                            weights, total, n_samples, use_origin);
  NB_LAUNCH_OK("ess_update_kernel");
  return 0;
}

extern "C" int nerfb200_ess_resample_compat(const uint8_t* grid, int res, const float* rays_o, const float* rays_d,
                                            int n_rays, int n_samples, int chunk, const float* z_table, float* z_vals,
                                            uint64_t* scratch, void* stream) {
  NB_CHECK_ARG(n_rays <= 0 || (grid && rays_o && rays_d && z_table && z_vals && scratch), "ess_resample_compat: null pointer");
  NB_CHECK_ARG(res >= 1 && res <= 1024, "ess_resample_compat: bad grid resolution %d", res);
  NB_CHECK_ARG(n_samples >= 1 && n_samples <= 64, "ess_resample_compat: n_samples=%d out of range [1,64]", n_samples);
  NB_CHECK_ARG(n_rays >= 0 && chunk >= 1, "ess_resample_compat: bad sizes");
  NB_CHECK_ARG(((uintptr_t)scratch & 7) == 0, "ess_resample_compat: scratch must be 8-byte aligned");
  if (n_rays == 0) return 0;
  ess_masks_kernel<<<ceil_div(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(grid, res, rays_o, rays_d, z_table, n_rays, n_samples,
                                                                            reinterpret_cast<unsigned long long*>(scratch));
  NB_LAUNCH_OK("ess_masks_kernel");
  ess_resample_compat_kernel<<<ceil_div(n_rays, chunk), 32, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const unsigned long long*>(scratch), z_table, n_rays, n_samples, chunk, z_vals);
  NB_LAUNCH_OK("ess_resample_compat_kernel");
  return 0;
}

extern "C" int nerfb200_jitter_rows(float* z_vals, int n_rays, int n_samples, uint64_t seed, void* stream) {
  NB_CHECK_ARG(n_rays <= 0 || z_vals, "jitter_rows: null pointer");
  NB_CHECK_ARG(n_rays >= 0 && n_samples >= 1 && n_samples <= kEssMaxS, "jitter_rows: bad sizes n_rays=%d S=%d", n_rays, n_samples);
  if (n_rays == 0) return 0;
  jitter_rows_kernel<<<ceil_div(n_rays, kJitWarps), kJitWarps * 32, 0, (cudaStream_t)stream>>>(z_vals, n_rays, n_samples, seed);
  NB_LAUNCH_OK("jitter_rows_kernel");
  return 0;
}
