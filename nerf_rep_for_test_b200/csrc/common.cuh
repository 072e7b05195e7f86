// Shared host/device helpers for libnerfb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/nerfb200.h"

namespace nb {

// network.py:22-47 / lego.yaml geometry
constexpr int kW = 256;         // hidden width
constexpr int kD = 8;           // pts_linears depth
constexpr int kSkip = 4;        // concat after layer 4 -> layer 5 has K = 63 + 256
constexpr int kLx = 10;         // xyz frequencies
constexpr int kLd = 4;          // dir frequencies
constexpr int kChX = 3 + 6 * kLx;  // 63
constexpr int kChD = 3 + 6 * kLd;  // 27
constexpr int kWv = 128;        // views_linears width

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define NB_CHECK_ARG(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      nb::set_error(__VA_ARGS__);    \
      return 1;                      \
    }                                \
  } while (0)

#define NB_CUDA(call)                                                                  \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      nb::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,               \
                    cudaGetErrorString(e__));                                          \
      return 2;                                                                        \
    }                                                                                  \
  } while (0)

// launch check: catches bad configurations immediately (no device sync)
#define NB_LAUNCH_OK(name)                                                             \
  do {                                                                                 \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess) {                                                          \
      nb::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));         \
      return 3;                                                                        \
    }                                                                                  \
    nb::count_launch();                                                                \
  } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Internal variants of public entry points that work on a LIST of rays (ray indices + a device-side count, both NULL
// = every ray 0..n_rays-1), as produced by the ray culling of the skipping mode (ess.cu).  The kernels then run as
// persistent grids that loop over the list, so culled rays cost neither a launch slot nor a memory access.  Used by
// the sparse whole-pass driver only; the public signatures stay as they are.
struct RayList {
  const int32_t* rays;
  const int32_t* count;
};
constexpr int kPersistentBlocks = 148 * 16;   // grid cap of the list-driven launches
int composite_forward_culled(const float* raw, const float* z_vals, const float* rays_d, const uint32_t* keep_bits,
                             RayList rl, int n_rays, int n_samples, int variant, float ert_threshold,
                             int white_bkgd, int compat_chunk, float* rgb_map, float* disp_map, float* acc_map,
                             float* depth_map, float* weights, void* stream);
// ERT_COMPAT (reference chunk quirk) as two fully parallel launches; low_flag [n_rays] bytes and chunk_any
// [ceil(n_rays / compat_chunk)] ints are scratch
int composite_forward_compat2(const float* raw, const float* z_vals, const float* rays_d, int n_rays, int n_samples,
                              int fast, float ert_threshold, int white_bkgd, int compat_chunk, float* rgb_map,
                              float* disp_map, float* acc_map, float* depth_map, float* weights, uint8_t* low_flag,
                              int32_t* chunk_any, void* stream);
int sample_pdf_merge_culled(const float* z_coarse, const float* weights, const float* u, int u_per_ray,
                            RayList rl, int n_rays, int n_samples, int n_u, float* z_all, void* stream);
int ess_compact_culled(const uint8_t* grid, int res, const float* rays_o, const float* rays_d, const float* z_vals,
                       const float* z_term, RayList rl, int n_rays, int n_samples, int32_t* row_ids,
                       int32_t* n_active, uint32_t* keep_bits, void* stream);
int ert_depth_culled(const float* weights, const float* z_vals, RayList rl, int n_rays, int n_samples,
                     float thr, float* z_term, void* stream);
// slab test of every ray against the box; flags (may be NULL) get 0/1 per ray; when list/count are given the hit rays
// are appended to list (count must be zeroed by the caller) and every culled ray gets its background maps written
// (maps_c / maps_f: rgb, disp, acc, depth pointers, entries may be NULL)
int ray_cull_list(const float* rays_o, const float* rays_d, int n_rays, const float* z_table, int n_samples,
                  const float* box_lo, const float* box_hi, uint8_t* flags, int32_t* list, int32_t* count,
                  const nerfb200_maps* maps_c, const nerfb200_maps* maps_f, int white_bkgd, void* stream);

// Counter-based RNG shared by the stratified jitter (a2) and the density noise (a5): a uniform in
// [0,1) keyed on (seed, a, b).  The reference draws from torch's global generator, whose stream a
// kernel cannot reproduce; only the distribution is matched.
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ float uniform01(uint64_t seed, uint32_t a, uint32_t b) {
  uint32_t h = mix32((uint32_t)seed ^ mix32(a * 0x9E3779B9U + 0x85ebca6bU));
  h = mix32(h ^ (uint32_t)(seed >> 32) ^ mix32(b + 0xc2b2ae35U));
  return (float)(h >> 8) * (1.0f / 16777216.0f);  // [0,1)
}
#endif

}  // namespace nb
