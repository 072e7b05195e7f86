// C-ABI glue: error string, launch counter, MLP dispatch and the whole-pass drivers.
// Reference for the pass order: volume_renderer.py:154-205.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "common.cuh"
#include "train_layout.cuh"

// the ctypes binding (nerf_rep_for_test_b200/lib.py) mirrors these layouts field by field
static_assert(sizeof(nerfb200_render_params) == 104, "nerfb200_render_params layout changed: update lib.py");
static_assert(sizeof(nerfb200_mlp_weights) == 24 * sizeof(void*), "nerfb200_mlp_weights layout changed");
static_assert(sizeof(nerfb200_maps) == 4 * sizeof(void*), "nerfb200_maps layout changed");

namespace nb {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int launch_mlp_fp32(const void* packed, const float* rays_o, const float* rays_d, const float* z_vals,
                    int n_rays, int n_samples, float* raw, cudaStream_t st);
int launch_mlp_bf16(const void* packed, const float* rays_o, const float* rays_d, const float* z_vals,
                    int n_rays, int n_samples, float* raw, float* stage_dump, void* acts, void* masks,
                    const int* row_ids, const int* n_active, bool f16, cudaStream_t st);
int launch_mlp_f16x2(const void* packed, const float* rays_o, const float* rays_d, const float* z_vals, int n_rays,
                     int n_samples, float* raw, float* stage_dump, const int* row_ids, const int* n_active,
                     cudaStream_t st);
int launch_mlp_bwd_dgrad(const void* packed_bwd, const float* g_raw, const void* masks, void* dacts, long long M,
                         cudaStream_t st);
int launch_mlp_bwd_wgrad(const void* acts, const void* dacts, long long M, float* scratch, const nerfb200_mlp_weights* weights,
                         const nerfb200_mlp_grads* grads, cudaStream_t st);

// ---- optional MLP-kernel timing (bench.py roofline): CUDA events around every mlp launch, on
// the launching stream, while enabled.  Off by default; the only other global state besides the
// error string and the launch counter.
constexpr int kProfMax = 8192;
static bool g_prof_on = false;
static cudaEvent_t g_prof_ev[2 * kProfMax];
static int g_prof_created = 0, g_prof_n = 0;
static double g_prof_rows = 0.0;

static bool prof_begin(cudaStream_t st) {
  if (!g_prof_on || g_prof_n >= kProfMax) return false;
  while (g_prof_created <= g_prof_n) {
    if (cudaEventCreate(&g_prof_ev[2 * g_prof_created]) != cudaSuccess) return false;
    if (cudaEventCreate(&g_prof_ev[2 * g_prof_created + 1]) != cudaSuccess) return false;
    ++g_prof_created;
  }
  cudaEventRecord(g_prof_ev[2 * g_prof_n], st);
  return true;
}
static void prof_end(cudaStream_t st, double rows) {
  cudaEventRecord(g_prof_ev[2 * g_prof_n + 1], st);
  g_prof_rows += rows;
  ++g_prof_n;
}

// rays per internal chunk of the whole-pass driver: bounds the workspace (z, raw, weights: ~5.4 KB/ray = 177 MB).
// Between two MLP launches of a chunk sit three short kernels (compositing, sample_pdf, coarse z) during which the
// tensor pipe idles, so fewer, larger chunks win even though their intermediates no longer fit the 126 MB L2
// (measured, same box, 800x800 frame: 8192 rays 138.8 ms, 16384 137.2, 24576 136.9, 32768 136.7).
constexpr int kChunkRays = 32768;
// Empty-space skipping sends only ~10 % of the rows through the MLP, so an 8192-ray chunk is a fraction of a
// wave of the persistent kernel (measured: 22.9 ms per frame at 8192, 17.5 ms at 131 072 rays per chunk).
// Round 2: 327 680 rays = two chunks per 800x800 frame, ~28 stream operations instead of ~75 (same frame time within the
// run-to-run noise: 14.97 / 14.46 / 14.72 ms at 131 072 / 262 144 / 655 360); the workspace grows to 1.8 GB.
constexpr int kChunkRaysSparse = 327680;
// nerfb200_render_params.mode: bits 0-7 = fine-pass mode, bits 8-15 = 1 + coarse-pass mode (0: same as fine)
static int mode_fine(int mode) { return mode & 0xFF; }
static int mode_coarse(int mode) { return (mode >> 8) & 0xFF ? ((mode >> 8) & 0xFF) - 1 : (mode & 0xFF); }
static bool mode_known(int m) { return m == NERFB200_MODE_FP32 || m == NERFB200_MODE_BF16 || m == NERFB200_MODE_FP32_TC || m == NERFB200_MODE_FP16; }
static bool mode_tensor(int m) { return m == NERFB200_MODE_BF16 || m == NERFB200_MODE_FP32_TC || m == NERFB200_MODE_FP16; }   // persistent CTA-pair kernels
static bool mode_exact(int m) { return m == NERFB200_MODE_FP32 || m == NERFB200_MODE_FP32_TC; }   // fp32-accurate MLP arithmetic

static int chunk_rays(const nerfb200_render_params* p) {
  if (const char* e = getenv("NERFB200_CHUNK_RAYS")) { int c = atoi(e); if (c >= 2048 && c % 2048 == 0) return c; }   // tuning experiments
  if (p->occupancy_grid && p->ess_skip) return kChunkRaysSparse;
  if (p->occupancy_grid && p->ess_ref_compat) return kChunkRays;   // the shared-row blocks (compat_chunk rays) must tile the chunk
  // The literal ERT_COMPAT compositor groups rays in compat_chunk blocks, which must tile the chunk.
  if ((p->variant & ~NERFB200_COMPOSITE_FAST_MATH) == NERFB200_COMPOSITE_ERT_COMPAT || !mode_tensor(mode_fine(p->mode)) ||
      !mode_tensor(mode_coarse(p->mode))) return kChunkRays;
  // tensor-core dense path: the persistent MLP kernel deals 512-row quads (256-row pairs in the split-fp16 mode) to
  // (SMs / 2) CTA pairs, so a chunk whose coarse
  // launch (n*S/512 quads) is a whole number of rounds wastes no partial last round; with S = 64 and S + U = 192
  // that holds for the fine launch too (32 560 rays on 148 SMs: 55 and 165 full rounds instead of 55.35 and 166.05).
  static int unit = 0;
  if (unit == 0) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 2)
      sms = 148;
    unit = (sms / 2) * 8;   // rays per round of the coarse launch at S = 64
  }
  const int c = kChunkRays / unit * unit;
  return c > 0 ? c : kChunkRays;
}

struct Workspace {
  float* z_coarse;  // [c,S]
  float* raw_c;     // [c,S,4]
  float* weights;   // [c,S]
  float* z_all;     // [c,S+U]
  float* raw_f;     // [c,S+U,4]
  int32_t* row_ids; // [c,S+U] compacted active rows (sparse ESS mode)
  int32_t* counters;// [2] active-row counts of the coarse / fine pass
  float* z_term;    // [c] ERT depth from the coarse pass
  uint32_t* keep_bits;  // [ceil(c*(S+U)/32)] bit m = row m was evaluated (sparse ESS mode)
  int32_t* ray_list;    // [c] rays that survived the culling (sparse ESS mode with cull_rays); count = counters[2]
                        //     (ERT_COMPAT, never combined with culling: per-ray "went low" flags, as bytes)
  int32_t* compat_any;  // [c] ERT_COMPAT: per-compat-chunk "somebody went low" flags
  size_t bytes;
};

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static Workspace carve(void* base, int chunk, int S, int U) {
  Workspace w;
  size_t off = 0;
  auto take = [&](size_t nfloats) {
    float* p = base ? reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off) : nullptr;
    off += align256(nfloats * 4);
    return p;
  };
  w.z_coarse = take((size_t)chunk * S);
  w.raw_c = take((size_t)chunk * S * 4);
  w.weights = take((size_t)chunk * S);
  w.z_all = take((size_t)chunk * (S + U));
  w.raw_f = take((size_t)chunk * (S + U) * 4);
  w.row_ids = reinterpret_cast<int32_t*>(take((size_t)chunk * (S + U)));
  w.counters = reinterpret_cast<int32_t*>(take(64));
  w.z_term = take((size_t)chunk);
  w.keep_bits = reinterpret_cast<uint32_t*>(take(((size_t)chunk * (S + U) + 31) / 32));
  w.ray_list = reinterpret_cast<int32_t*>(take((size_t)chunk));
  w.compat_any = reinterpret_cast<int32_t*>(take((size_t)chunk));
  w.bytes = off;
  return w;
}

static int check_params(const nerfb200_render_params* p) {
  NB_CHECK_ARG(p, "render: null params");
  NB_CHECK_ARG(p->n_samples >= 3 && p->n_samples <= 256, "render: n_samples=%d out of range [3,256]", p->n_samples);
  NB_CHECK_ARG(p->n_importance >= 0 && p->n_importance <= 256 && p->n_samples + p->n_importance <= 256,
               "render: n_importance=%d out of range (n_samples+n_importance <= 256)", p->n_importance);
  NB_CHECK_ARG((p->mode >> 16) == 0 && mode_known(mode_fine(p->mode)) && mode_known(mode_coarse(p->mode)),
               "render: unknown mode 0x%x", p->mode);
  NB_CHECK_ARG(!(p->occupancy_grid && p->ess_skip) || (mode_tensor(mode_fine(p->mode)) && mode_tensor(mode_coarse(p->mode))),
               "render: ess_skip needs a tensor-core mode (sparse MLP launch)");
  NB_CHECK_ARG((p->variant & ~NERFB200_COMPOSITE_FAST_MATH) >= 0 && (p->variant & ~NERFB200_COMPOSITE_FAST_MATH) <= 2,
               "render: unknown composite variant %d", p->variant);
  NB_CHECK_ARG((p->variant & ~NERFB200_COMPOSITE_FAST_MATH) != NERFB200_COMPOSITE_ERT_COMPAT || (p->compat_chunk > 0 && kChunkRays % p->compat_chunk == 0),
               "render: compat_chunk=%d must divide %d", p->compat_chunk, kChunkRays);
  NB_CHECK_ARG(!(p->occupancy_grid && p->ess_ref_compat && !p->ess_skip) ||
                   (p->compat_chunk > 0 && kChunkRays % p->compat_chunk == 0 && p->n_samples <= 64),
               "render: ess_ref_compat needs n_samples <= 64 and a compat_chunk that divides %d", kChunkRays);
  static_assert(kChunkRaysSparse % kChunkRays == 0, "chunk sizes");
  NB_CHECK_ARG(p->raw_noise_std >= 0.f, "render: raw_noise_std=%g must be >= 0", (double)p->raw_noise_std);
  NB_CHECK_ARG(!(p->raw_noise_std > 0.f && p->occupancy_grid && p->ess_skip),
               "render: raw_noise_std > 0 (a training regulariser) cannot be combined with ess_skip (skipped rows have zero density)");
  return 0;
}

}  // namespace nb

using namespace nb;

extern "C" int nerfb200_abi_version(void) { return NERFB200_ABI_VERSION; }
extern "C" const char* nerfb200_get_last_error_string(void) { return g_err; }
extern "C" uint64_t nerfb200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int nerfb200_mlp_forward(const void* packed, int mode, const float* rays_o, const float* rays_d,
                                    const float* z_vals, int n_rays, int n_samples, float* raw, void* stream) {
  NB_CHECK_ARG(n_rays <= 0 || (packed && rays_o && rays_d && z_vals && raw), "mlp_forward: null pointer");
  NB_CHECK_ARG(n_rays >= 0 && n_samples >= 1, "mlp_forward: bad sizes n_rays=%d n_samples=%d", n_rays, n_samples);
  NB_CHECK_ARG(((uintptr_t)packed & 1023) == 0, "mlp_forward: packed weights must be 1024-byte aligned");
  if (n_rays == 0) return 0;
  NB_CHECK_ARG(mode_known(mode), "mlp_forward: unknown mode %d", mode);
  cudaStream_t st = (cudaStream_t)stream;
  bool prof = prof_begin(st);
  int rc = mode == NERFB200_MODE_FP32 ? launch_mlp_fp32(packed, rays_o, rays_d, z_vals, n_rays, n_samples, raw, st)
           : mode == NERFB200_MODE_FP32_TC ? launch_mlp_f16x2(packed, rays_o, rays_d, z_vals, n_rays, n_samples, raw, nullptr, nullptr, nullptr, st)
                                           : launch_mlp_bf16(packed, rays_o, rays_d, z_vals, n_rays, n_samples, raw, nullptr, nullptr, nullptr, nullptr, nullptr, mode == NERFB200_MODE_FP16, st);
  if (prof) prof_end(st, (double)n_rays * n_samples);
  return rc;
}

extern "C" size_t nerfb200_train_acts_bytes(long long n_rows) {
  return n_rows <= 0 ? 0 : (size_t)((n_rows + 127) / 128) * kActBlocks * kBlockBytes;
}
extern "C" size_t nerfb200_train_masks_bytes(long long n_rows) {
  return n_rows <= 0 ? 0 : (size_t)((n_rows + 127) / 128) * 128 * kMaskWords * 4 * kMaskPlanes;
}
extern "C" size_t nerfb200_mlp_backward_workspace_bytes(long long n_rows) {
  return n_rows <= 0 ? 0 : (size_t)((n_rows + 127) / 128) * kDactBlocks * kBlockBytes + kGradScratchFloats * sizeof(float);
}

extern "C" int nerfb200_mlp_forward_train(const void* packed, int mode, const float* rays_o, const float* rays_d,
                                          const float* z_vals, int n_rays, int n_samples, float* raw, void* acts,
                                          void* masks, void* stream) {
  NB_CHECK_ARG(n_rays <= 0 || (packed && rays_o && rays_d && z_vals && raw && acts && masks), "mlp_forward_train: null pointer");
  NB_CHECK_ARG(n_rays >= 0 && n_samples >= 1, "mlp_forward_train: bad sizes");
  NB_CHECK_ARG(mode == NERFB200_MODE_BF16, "mlp_forward_train: only NERFB200_MODE_BF16 saves activations");
  NB_CHECK_ARG(((uintptr_t)packed & 1023) == 0 && ((uintptr_t)acts & 127) == 0 && ((uintptr_t)masks & 15) == 0,
               "mlp_forward_train: misaligned buffer");
  if (n_rays == 0) return 0;
  return launch_mlp_bf16(packed, rays_o, rays_d, z_vals, n_rays, n_samples, raw, nullptr, acts, masks, nullptr, nullptr, false, (cudaStream_t)stream);
}

extern "C" int nerfb200_mlp_backward(const void* packed_bwd, const nerfb200_mlp_weights* weights, const float* g_raw,
                                     const void* acts, const void* masks, long long n_rows, void* workspace,
                                     size_t workspace_bytes, const nerfb200_mlp_grads* grads, void* stream) {
  NB_CHECK_ARG(n_rows >= 0, "mlp_backward: bad n_rows");
  NB_CHECK_ARG(grads, "mlp_backward: null grads");
  NB_CHECK_ARG(weights && weights->views_w && weights->feature_w && weights->feature_b,
               "mlp_backward: the fp32 views_linears.0 / feature_linear tensors are needed for the fused-tail chain rule");
  for (int i = 0; i < 8; ++i) NB_CHECK_ARG(grads->pts_w[i] && grads->pts_b[i], "mlp_backward: null gradient pts_linears.%d", i);
  NB_CHECK_ARG(grads->views_w && grads->views_b && grads->feature_w && grads->feature_b && grads->alpha_w && grads->alpha_b &&
                   grads->rgb_w && grads->rgb_b, "mlp_backward: null head gradient");
  NB_CHECK_ARG(n_rows == 0 || (packed_bwd && g_raw && acts && masks && workspace), "mlp_backward: null pointer");
  NB_CHECK_ARG(workspace_bytes >= nerfb200_mlp_backward_workspace_bytes(n_rows), "mlp_backward: workspace too small (%zu < %zu)",
               workspace_bytes, nerfb200_mlp_backward_workspace_bytes(n_rows));
  NB_CHECK_ARG(((uintptr_t)packed_bwd & 1023) == 0 && ((uintptr_t)acts & 127) == 0 && ((uintptr_t)workspace & 127) == 0 &&
                   ((uintptr_t)masks & 15) == 0 && ((uintptr_t)g_raw & 15) == 0, "mlp_backward: misaligned buffer");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t dacts_bytes = (size_t)((n_rows + 127) / 128) * kDactBlocks * kBlockBytes;
  float* scratch = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + dacts_bytes);
  if (n_rows > 0) {
    int rc = launch_mlp_bwd_dgrad(packed_bwd, g_raw, masks, workspace, n_rows, st);
    if (rc) return rc;
  } else {
    NB_CHECK_ARG(workspace && workspace_bytes >= kGradScratchFloats * sizeof(float), "mlp_backward: workspace too small");
  }
  return launch_mlp_bwd_wgrad(acts, workspace, n_rows, scratch, weights, grads, st);
}

// zero_fill = false: rows outside row_ids keep whatever raw held; the caller composites with the keep-bit mask
static int mlp_forward_sparse_impl(const void* packed, int mode, const float* rays_o, const float* rays_d,
                                   const float* z_vals, int n_rays, int n_samples, const int32_t* row_ids,
                                   const int32_t* n_active, float* raw, bool zero_fill, void* stream) {
  NB_CHECK_ARG(n_rays <= 0 || (packed && rays_o && rays_d && z_vals && raw && row_ids && n_active),
               "mlp_forward_sparse: null pointer");
  NB_CHECK_ARG(n_rays >= 0 && n_samples >= 1, "mlp_forward_sparse: bad sizes");
  NB_CHECK_ARG(mode_tensor(mode), "mlp_forward_sparse: only the tensor-core modes have the sparse launch");
  NB_CHECK_ARG(((uintptr_t)packed & 1023) == 0, "mlp_forward_sparse: packed weights must be 1024-byte aligned");
  if (n_rays == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  // skipped rows keep raw = 0: sigma_raw = 0 -> alpha = 0 -> no contribution (and rgb_raw is never used)
  if (zero_fill) NB_CUDA(cudaMemsetAsync(raw, 0, (size_t)n_rays * n_samples * 16, st));
  bool prof = prof_begin(st);
  int rc = mode == NERFB200_MODE_FP32_TC
               ? launch_mlp_f16x2(packed, rays_o, rays_d, z_vals, n_rays, n_samples, raw, nullptr, row_ids, n_active, st)
               : launch_mlp_bf16(packed, rays_o, rays_d, z_vals, n_rays, n_samples, raw, nullptr, nullptr, nullptr, row_ids, n_active, mode == NERFB200_MODE_FP16, st);
  if (prof) prof_end(st, 0.0);   // evaluated rows are data dependent; counted by the caller from n_active
  return rc;
}

extern "C" int nerfb200_mlp_forward_sparse(const void* packed, int mode, const float* rays_o, const float* rays_d,
                                           const float* z_vals, int n_rays, int n_samples, const int32_t* row_ids,
                                           const int32_t* n_active, float* raw, void* stream) {
  return mlp_forward_sparse_impl(packed, mode, rays_o, rays_d, z_vals, n_rays, n_samples, row_ids, n_active, raw, true, stream);
}

extern "C" int nerfb200_profile_enable(int on) {
  g_prof_on = on != 0;
  g_prof_n = 0;
  g_prof_rows = 0.0;
  return 0;
}

extern "C" int nerfb200_profile_read(double* mlp_ms, uint64_t* mlp_launches, double* mlp_rows) {
  NB_CHECK_ARG(mlp_ms && mlp_launches && mlp_rows, "profile_read: null pointer");
  double total = 0.0;
  for (int i = 0; i < g_prof_n; ++i) {
    NB_CUDA(cudaEventSynchronize(g_prof_ev[2 * i + 1]));
    float ms = 0.f;
    NB_CUDA(cudaEventElapsedTime(&ms, g_prof_ev[2 * i], g_prof_ev[2 * i + 1]));
    total += ms;
  }
  *mlp_ms = total;
  *mlp_launches = (uint64_t)g_prof_n;
  *mlp_rows = g_prof_rows;
  return 0;
}

extern "C" int nerfb200_mlp_forward_stages(const void* packed, int mode, const float* rays_o, const float* rays_d,
                                           const float* z_vals, int n_rays, int n_samples, float* raw,
                                           float* stage_dump, void* stream) {
  NB_CHECK_ARG(packed && rays_o && rays_d && z_vals && raw && stage_dump, "mlp_forward_stages: null pointer");
  NB_CHECK_ARG(n_rays >= 1 && n_samples >= 1, "mlp_forward_stages: bad sizes");
  NB_CHECK_ARG(mode == NERFB200_MODE_BF16 || mode == NERFB200_MODE_FP32_TC, "mlp_forward_stages: only NERFB200_MODE_BF16 / FP32_TC have a stage dump");
  NB_CHECK_ARG(((uintptr_t)packed & 1023) == 0, "mlp_forward_stages: packed weights must be 1024-byte aligned");
  if (mode == NERFB200_MODE_FP32_TC)
    return launch_mlp_f16x2(packed, rays_o, rays_d, z_vals, n_rays, n_samples, raw, stage_dump, nullptr, nullptr, (cudaStream_t)stream);
  return launch_mlp_bf16(packed, rays_o, rays_d, z_vals, n_rays, n_samples, raw, stage_dump, nullptr, nullptr, nullptr, nullptr, false, (cudaStream_t)stream);
}

extern "C" size_t nerfb200_render_workspace_bytes(int n_rays, const nerfb200_render_params* p) {
  if (!p || n_rays < 0) return 0;
  int chunk = n_rays < chunk_rays(p) ? n_rays : chunk_rays(p);
  if (chunk == 0) chunk = 1;
  return carve(nullptr, chunk, p->n_samples, p->n_importance).bytes;
}

// called after the last kernel of every chunk has been enqueued (rays [r0, r0 + n) of the call are final once the
// stream reaches this point); used by the host-buffer entry to start the device-to-host copies early
struct ChunkHook {
  virtual int done(int r0, int n) = 0;
};

static int render_rays_impl(const void* packed_coarse, const void* packed_fine, const float* rays_o,
                            const float* rays_d, int n_rays, const float* z_table, const float* u,
                            const nerfb200_render_params* p, void* workspace, size_t workspace_bytes,
                            const nerfb200_maps* mc, const nerfb200_maps* mf, void* stream, ChunkHook* hook) {
  if (int e = check_params(p)) return e;
  if (n_rays == 0) return 0;
  NB_CHECK_ARG(packed_coarse && rays_o && rays_d && z_table && mc, "render_rays: null pointer");
  NB_CHECK_ARG(mc->rgb && mc->disp && mc->acc && mc->depth, "render_rays: null coarse map");
  NB_CHECK_ARG(n_rays >= 0, "render_rays: negative n_rays");
  const int S = p->n_samples, U = p->n_importance;
  const int mode_c = mode_coarse(p->mode), mode_f = mode_fine(p->mode);
  // mixed precision: an accurate coarse pass keeps its exact compositor (its weights place the fine samples)
  const int variant_c = mode_exact(mode_c) ? (p->variant & ~NERFB200_COMPOSITE_FAST_MATH) : p->variant;
  if (U > 0) {
    NB_CHECK_ARG(packed_fine && u && mf && mf->rgb && mf->disp && mf->acc && mf->depth,
                 "render_rays: fine pass needs packed_fine, u and fine maps");
  }
  if (n_rays == 0) return 0;
  NB_CHECK_ARG(workspace && workspace_bytes >= nerfb200_render_workspace_bytes(n_rays, p),
               "render_rays: workspace too small (%zu < %zu)", workspace_bytes,
               nerfb200_render_workspace_bytes(n_rays, p));
  int chunk = n_rays < chunk_rays(p) ? n_rays : chunk_rays(p);
  Workspace ws = carve(workspace, chunk, S, U);
  for (int r0 = 0; r0 < n_rays; r0 += chunk) {
    int n = (n_rays - r0) < chunk ? (n_rays - r0) : chunk;
    const float* ro = rays_o + (size_t)r0 * 3;
    const float* rd = rays_d + (size_t)r0 * 3;
    int e;
    // per-ray jitter is keyed on the ray index inside the call: offset the seed per chunk
    const bool sparse = p->occupancy_grid && p->ess_skip;
    const bool resample = p->occupancy_grid && !sparse;
    const uint64_t jitter_seed = p->seed + (uint64_t)r0 * 0x9E3779B97F4A7C15ull;
    // reference order (:1009-1085): occupancy is tested at the UNPERTURBED depths, highly-empty rays are resampled, and
    // only then is the stratified jitter applied, from each row's own mid-points
    if ((e = nerfb200_sample_coarse(z_table, n, S, resample ? 0 : p->perturb, jitter_seed, ws.z_coarse, stream))) return e;
    if (resample) {
      if (p->ess_ref_compat) {
        if ((e = nerfb200_ess_resample_compat(p->occupancy_grid, p->grid_res, ro, rd, n, S, p->compat_chunk, z_table, ws.z_coarse,
                                              reinterpret_cast<uint64_t*>(ws.row_ids), stream))) return e;
      } else if ((e = nerfb200_ess_resample(p->occupancy_grid, p->grid_res, ro, rd, n, S, ws.z_coarse, nullptr, stream))) return e;
      if (p->perturb && (e = nerfb200_jitter_rows(ws.z_coarse, n, S, jitter_seed, stream))) return e;
    }
    // sparse: the compaction leaves one keep bit per row; the MLP writes only the listed rows and the masked
    // compositor never touches the others (no zero fill of raw, no 16 B read per skipped row)
    // (the literal ERT_COMPAT compositor has no masked form: that combination keeps the zero fill)
    const bool masked = sparse && (p->variant & ~NERFB200_COMPOSITE_FAST_MATH) != NERFB200_COMPOSITE_ERT_COMPAT;
    const uint32_t* keep = masked ? ws.keep_bits : nullptr;
    // ray-level culling against the occupied box, before any per-sample work
    // (list-driven kernels need whole 32-sample words per ray)
    RayList active = {nullptr, nullptr};
    if (masked && p->cull_rays && S % 32 == 0 && (S + U) % 32 == 0) {
      NB_CUDA(cudaMemsetAsync(ws.counters + 2, 0, sizeof(int32_t), (cudaStream_t)stream));
      nerfb200_maps c_maps = {mc->rgb + (size_t)r0 * 3, mc->disp + r0, mc->acc + r0, mc->depth + r0};
      nerfb200_maps f_maps = {nullptr, nullptr, nullptr, nullptr};
      if (U > 0) f_maps = nerfb200_maps{mf->rgb + (size_t)r0 * 3, mf->disp + r0, mf->acc + r0, mf->depth + r0};
      if ((e = ray_cull_list(ro, rd, n, z_table, S, p->cull_lo, p->cull_hi, nullptr, ws.ray_list, ws.counters + 2, &c_maps,
                             &f_maps, p->white_bkgd, stream))) return e;
      active = RayList{ws.ray_list, ws.counters + 2};
    }
    const uint64_t noise_seed = p->seed ^ ((uint64_t)r0 * 0xD6E8FEB86659FD93ull);
    if (sparse) {
      if ((e = ess_compact_culled(p->occupancy_grid, p->grid_res, ro, rd, ws.z_coarse, nullptr, active, n, S, ws.row_ids,
                                  ws.counters + 0, ws.keep_bits, stream))) return e;
      if ((e = mlp_forward_sparse_impl(packed_coarse, mode_c, ro, rd, ws.z_coarse, n, S, ws.row_ids, ws.counters + 0,
                                       ws.raw_c, !masked, stream))) return e;
    } else if ((e = nerfb200_mlp_forward(packed_coarse, mode_c, ro, rd, ws.z_coarse, n, S, ws.raw_c, stream))) return e;
    if (p->raw_noise_std > 0.f &&
        (e = nerfb200_sigma_noise(ws.raw_c, (long long)n * S, p->raw_noise_std, noise_seed + 0x632BE59BD9B4E019ull, stream))) return e;
    const bool compat = (p->variant & ~NERFB200_COMPOSITE_FAST_MATH) == NERFB200_COMPOSITE_ERT_COMPAT;
    const int fast = (p->variant & NERFB200_COMPOSITE_FAST_MATH) != 0;
    const int fast_c = (variant_c & NERFB200_COMPOSITE_FAST_MATH) != 0;
    if (compat) {
      if ((e = composite_forward_compat2(ws.raw_c, ws.z_coarse, rd, n, S, fast_c, p->ert_threshold, p->white_bkgd,
                                         p->compat_chunk, mc->rgb + (size_t)r0 * 3, mc->disp + r0, mc->acc + r0,
                                         mc->depth + r0, ws.weights, reinterpret_cast<uint8_t*>(ws.ray_list),
                                         ws.compat_any, stream))) return e;
    } else if ((e = composite_forward_culled(ws.raw_c, ws.z_coarse, rd, keep, active, n, S, variant_c, p->ert_threshold,
                                      p->white_bkgd, p->compat_chunk, mc->rgb + (size_t)r0 * 3, mc->disp + r0,
                                      mc->acc + r0, mc->depth + r0, ws.weights, stream))) return e;
    if (U > 0) {
      const float* uu = p->u_per_ray ? u + (size_t)r0 * U : u;
      if ((e = sample_pdf_merge_culled(ws.z_coarse, ws.weights, uu, p->u_per_ray, active, n, S, U, ws.z_all, stream))) return e;
      if (sparse) {
        // fine pass: skip samples in empty cells and, with ERT, samples behind the depth at which the
        // coarse transmittance fell below the threshold
        const float* zt = nullptr;
        if ((p->variant & ~NERFB200_COMPOSITE_FAST_MATH) != NERFB200_COMPOSITE_PLAIN) {
          if ((e = ert_depth_culled(ws.weights, ws.z_coarse, active, n, S, p->ert_threshold, ws.z_term, stream))) return e;
          zt = ws.z_term;
        }
        if ((e = ess_compact_culled(p->occupancy_grid, p->grid_res, ro, rd, ws.z_all, zt, active, n, S + U, ws.row_ids,
                                    ws.counters + 1, ws.keep_bits, stream))) return e;
        if ((e = mlp_forward_sparse_impl(packed_fine, mode_f, ro, rd, ws.z_all, n, S + U, ws.row_ids, ws.counters + 1,
                                         ws.raw_f, !masked, stream))) return e;
        if (p->eval_counts) {   // optional statistics: evaluated rows per pass, accumulated over the call
          if ((e = nerfb200_accumulate_counts(ws.counters, p->eval_counts, stream))) return e;
        }
      } else if ((e = nerfb200_mlp_forward(packed_fine, mode_f, ro, rd, ws.z_all, n, S + U, ws.raw_f, stream))) return e;
      if (p->raw_noise_std > 0.f &&
          (e = nerfb200_sigma_noise(ws.raw_f, (long long)n * (S + U), p->raw_noise_std, noise_seed + 0x94D049BB133111EBull, stream))) return e;
      if (compat) {
        if ((e = composite_forward_compat2(ws.raw_f, ws.z_all, rd, n, S + U, fast, p->ert_threshold, p->white_bkgd,
                                           p->compat_chunk, mf->rgb + (size_t)r0 * 3, mf->disp + r0, mf->acc + r0,
                                           mf->depth + r0, nullptr, reinterpret_cast<uint8_t*>(ws.ray_list),
                                           ws.compat_any, stream))) return e;
      } else if ((e = composite_forward_culled(ws.raw_f, ws.z_all, rd, keep, active, n, S + U, p->variant, p->ert_threshold,
                                        p->white_bkgd, p->compat_chunk, mf->rgb + (size_t)r0 * 3, mf->disp + r0,
                                        mf->acc + r0, mf->depth + r0, nullptr, stream))) return e;
    }
    if (hook && (e = hook->done(r0, n))) return e;
  }
  return 0;
}

extern "C" int nerfb200_render_rays(const void* packed_coarse, const void* packed_fine, const float* rays_o,
                                    const float* rays_d, int n_rays, const float* z_table, const float* u,
                                    const nerfb200_render_params* p, void* workspace, size_t workspace_bytes,
                                    const nerfb200_maps* mc, const nerfb200_maps* mf, void* stream) {
  return render_rays_impl(packed_coarse, packed_fine, rays_o, rays_d, n_rays, z_table, u, p, workspace, workspace_bytes,
                          mc, mf, stream, nullptr);
}

// ---- host-buffer entry ------------------------------------------------------------------------
// device scratch: pose(16)+K(9) | rays_o | rays_d | 8 maps | render workspace
namespace nb {
struct ImageScratch {
  float* cam;  // 32 floats
  float* rays_o;
  float* rays_d;
  float* maps;  // 2 x (3+1+1+1) x n
  void* ws;
  size_t ws_bytes;
  size_t bytes;
};
static ImageScratch carve_image(void* base, int n, const nerfb200_render_params* p) {
  ImageScratch s;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* q = base ? reinterpret_cast<char*>(base) + off : nullptr;
    off += align256(bytes);
    return q;
  };
  s.cam = (float*)take(32 * 4);
  s.rays_o = (float*)take((size_t)n * 12);
  s.rays_d = (float*)take((size_t)n * 12);
  s.maps = (float*)take((size_t)n * 12 * 4);
  s.ws_bytes = nerfb200_render_workspace_bytes(n, p);
  s.ws = take(s.ws_bytes);
  s.bytes = off;
  return s;
}
}  // namespace nb

extern "C" size_t nerfb200_render_image_workspace_bytes(int H, int W, const nerfb200_render_params* p) {
  if (!p || H <= 0 || W <= 0) return 0;
  return carve_image(nullptr, H * W, p).bytes;
}

extern "C" int nerfb200_render_image_host(const void* packed_coarse, const void* packed_fine,
                                          const float* pose_host, const float* intrinsics_host, int H, int W,
                                          const float* z_table, const float* u, const nerfb200_render_params* p,
                                          void* workspace, size_t workspace_bytes,
                                          const nerfb200_maps* mc_host, const nerfb200_maps* mf_host, void* stream) {
  if (int e = check_params(p)) return e;
  NB_CHECK_ARG(pose_host && intrinsics_host && mc_host, "render_image_host: null pointer");
  NB_CHECK_ARG(H > 0 && W > 0 && (long long)H * W < (1LL << 28), "render_image_host: bad H=%d W=%d", H, W);
  NB_CHECK_ARG(workspace && workspace_bytes >= nerfb200_render_image_workspace_bytes(H, W, p),
               "render_image_host: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int n = H * W;
  ImageScratch s = carve_image(workspace, n, p);
  NB_CUDA(cudaMemcpyAsync(s.cam, pose_host, 16 * 4, cudaMemcpyHostToDevice, st));
  NB_CUDA(cudaMemcpyAsync(s.cam + 16, intrinsics_host, 9 * 4, cudaMemcpyHostToDevice, st));
  int e;
  if ((e = nerfb200_raygen(s.cam, s.cam + 16, H, W, s.rays_o, s.rays_d, stream))) return e;
  nerfb200_maps dc, df;
  float* m = s.maps;
  dc.rgb = m; dc.disp = m + (size_t)n * 3; dc.acc = m + (size_t)n * 4; dc.depth = m + (size_t)n * 5;
  m += (size_t)n * 6;
  df.rgb = m; df.disp = m + (size_t)n * 3; df.acc = m + (size_t)n * 4; df.depth = m + (size_t)n * 5;
  // The maps of a chunk are final as soon as its last compositing kernel has run: their device-to-host copies go to
  // a side stream behind an event, so that all but the last chunk's 1.5 MB leave while the next chunk renders
  // (the 30.7 MB of an 800x800 frame used to trail the render by ~0.8 ms).
  static cudaStream_t copy_stream[64] = {};
  static cudaEvent_t chunk_event[64] = {};
  static std::mutex copy_mutex[64];   // the side stream and its event are per device: concurrent calls on one device take turns
  int dev = 0;
  NB_CUDA(cudaGetDevice(&dev));
  NB_CHECK_ARG(dev >= 0 && dev < 64, "render_image_host: device ordinal %d out of range", dev);
  std::lock_guard<std::mutex> lock(copy_mutex[dev]);
  if (!copy_stream[dev]) {
    NB_CUDA(cudaStreamCreateWithFlags(&copy_stream[dev], cudaStreamNonBlocking));
    NB_CUDA(cudaEventCreateWithFlags(&chunk_event[dev], cudaEventDisableTiming));
  }
  struct CopyOut : ChunkHook {
    cudaStream_t st, cs;
    cudaEvent_t ev;
    const nerfb200_maps *dc, *df, *hc, *hf;
    int done(int r0, int n) override {
      NB_CUDA(cudaEventRecord(ev, st));
      NB_CUDA(cudaStreamWaitEvent(cs, ev, 0));
      for (int pass = 0; pass < 2; ++pass) {
        const nerfb200_maps* d = pass ? df : dc;
        const nerfb200_maps* h = pass ? hf : hc;
        if (!d || !h) continue;
        if (h->rgb) NB_CUDA(cudaMemcpyAsync(h->rgb + (size_t)r0 * 3, d->rgb + (size_t)r0 * 3, (size_t)n * 12, cudaMemcpyDeviceToHost, cs));
        if (h->disp) NB_CUDA(cudaMemcpyAsync(h->disp + r0, d->disp + r0, (size_t)n * 4, cudaMemcpyDeviceToHost, cs));
        if (h->acc) NB_CUDA(cudaMemcpyAsync(h->acc + r0, d->acc + r0, (size_t)n * 4, cudaMemcpyDeviceToHost, cs));
        if (h->depth) NB_CUDA(cudaMemcpyAsync(h->depth + r0, d->depth + r0, (size_t)n * 4, cudaMemcpyDeviceToHost, cs));
      }
      return 0;
    }
  } hook;
  hook.st = st; hook.cs = copy_stream[dev]; hook.ev = chunk_event[dev];
  hook.dc = &dc; hook.df = p->n_importance > 0 ? &df : nullptr;
  hook.hc = mc_host; hook.hf = p->n_importance > 0 ? mf_host : nullptr;
  if ((e = render_rays_impl(packed_coarse, packed_fine, s.rays_o, s.rays_d, n, z_table, u, p, s.ws, s.ws_bytes, &dc,
                            p->n_importance > 0 ? &df : nullptr, stream, &hook))) {
    cudaStreamSynchronize(copy_stream[dev]);
    return e;
  }
  NB_CUDA(cudaStreamSynchronize(copy_stream[dev]));
  NB_CUDA(cudaStreamSynchronize(st));
  return 0;
}
