// a7, fp32-ACCURATE training path of the NeRF MLP: forward with saved activations and the full backward --
// parameter gradients AND the gradient with respect to the sample depth z (the reference does not detach its
// hierarchical sampler, volume_renderer.py:181-183, so the fine loss reaches z through the MLP input).
// Reference arithmetic: fp32 nn.Linear chain under autograd, network.py:49-74, freq.py:23-26.
//
// This is the parity twin of the bf16 tensor-core training kernels (mlp_bf16_tc2 <kSave>, mlp_bwd_dgrad, mlp_bwd_wgrad):
// true fp32 FFMA arithmetic, layer by layer, activations kept as plain row-major fp32 matrices.  One generic tiled
// SGEMM (64x64x16 tile, 4x4 register tile, arbitrary operand strides so that X W^T, G W and G^T X are the same
// kernel, split-K with atomics for the row-reduction of the weight gradients) plus three small kernels
// (positional encoding, its backward, column sums for the bias gradients).  It exists so that gradients can be
// compared with the reference's autograd at 1e-4 -- it is not the performance path.
#include "common.cuh"

namespace nb {
namespace f32t {

constexpr int TM = 64, TN = 64, TK = 16;
enum { EPI_BIAS = 1, EPI_RELU = 2, EPI_MASK = 4, EPI_ACCUM = 8, EPI_ATOMIC = 16 };

struct Gemm {
  const float* A; long long sAi, sAk;      // A(i,k) = A[i*sAi + k*sAk]
  const float* B; long long sBk, sBj;      // B(k,j) = B[k*sBk + j*sBj]
  float* C; long long ldc;                 // C(i,j) = C[i*ldc + j]
  const float* bias;                       // EPI_BIAS: + bias[j]
  const float* mask; long long ldmask;     // EPI_MASK: result zeroed where mask[i*ldmask + j] <= 0 (relu')
  int M, N, K, flags, k_per_split;
};

__global__ void __launch_bounds__(256) sgemm_kernel(Gemm g) {
  __shared__ __align__(16) float As[TK][TM + 4];
  __shared__ __align__(16) float Bs[TK][TN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long i0 = (long long)blockIdx.y * TM;
  const int j0 = blockIdx.x * TN;
  const int k_begin = blockIdx.z * g.k_per_split;
  const int k_end = min(g.K, k_begin + g.k_per_split);
  const bool a_k_contig = g.sAk == 1, b_j_contig = g.sBj == 1;
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
  for (int k0 = k_begin; k0 < k_end; k0 += TK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = tid + 256 * r;
      int ii, kk;
      if (a_k_contig) { kk = e & 15; ii = e >> 4; } else { ii = e & 63; kk = e >> 6; }
      const long long gi = i0 + ii;
      const int gk = k0 + kk;
      As[kk][ii] = (gi < g.M && gk < k_end) ? g.A[gi * g.sAi + (long long)gk * g.sAk] : 0.f;
      int jj, kb;
      if (b_j_contig) { jj = e & 63; kb = e >> 6; } else { kb = e & 15; jj = e >> 4; }
      const int gj = j0 + jj, gkb = k0 + kb;
      Bs[kb][jj] = (gj < g.N && gkb < k_end) ? g.B[(long long)gkb * g.sBk + (long long)gj * g.sBj] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const long long gi = i0 + ty * 4 + r;
    if (gi >= g.M) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int gj = j0 + tx * 4 + c;
      if (gj >= g.N) continue;
      float* dst = g.C + gi * g.ldc + gj;
      float v = acc[r][c];
      if (g.flags & EPI_ATOMIC) { atomicAdd(dst, v); continue; }
      if (g.flags & EPI_ACCUM) v += *dst;
      if (g.flags & EPI_BIAS) v += g.bias[gj];
      if (g.flags & EPI_RELU) v = fmaxf(v, 0.f);
      if ((g.flags & EPI_MASK) && !(g.mask[gi * g.ldmask + gj] > 0.f)) v = 0.f;
      *dst = v;
    }
  }
}

// out[j] += sum_i X[i*ld + j]   (out pre-zeroed)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, long long ld, long long M, int N,
                                                     float* __restrict__ out) {
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  const long long r0 = (long long)blockIdx.y * 2048, r1 = min(M, r0 + 2048);
  float s = 0.f;
  if (j < N)
    for (long long i = r0 + ty; i < r1; i += 8) s += X[i * ld + j];
  part[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && j < N) {
#pragma unroll
    for (int k = 1; k < 8; ++k) s += part[k][tx];
    atomicAdd(out + j, s);
  }
}

// PE [M,64] = (x, sin(2^l x), cos(2^l x))_{l<10} of x = o + d z (fadd(o, fmul(d,z)), :165), column 63 = 0;
// DPE [M,32] = the same with 4 octaves of the view direction, columns 27..31 = 0.  Full-range sinf / cosf.
__global__ void __launch_bounds__(256) pe_forward_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                         const float* __restrict__ z_vals, long long M, int S,
                                                         float* __restrict__ pe, float* __restrict__ dpe) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const long long ray = m / S;
  const float z = z_vals[m];
  float p[3], d[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    d[c] = rays_d[ray * 3 + c];
    p[c] = __fadd_rn(rays_o[ray * 3 + c], __fmul_rn(d[c], z));
  }
  float* o = pe + m * 64;
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c] = p[c];
  for (int l = 0; l < kLx; ++l)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float a = p[c] * (float)(1 << l);
      o[3 + 6 * l + c] = sinf(a);
      o[3 + 6 * l + 3 + c] = cosf(a);
    }
  o[63] = 0.f;
  float* q = dpe + m * 32;
#pragma unroll
  for (int c = 0; c < 3; ++c) q[c] = d[c];
  for (int l = 0; l < kLd; ++l)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float a = d[c] * (float)(1 << l);
      q[3 + 6 * l + c] = sinf(a);
      q[3 + 6 * l + 3 + c] = cosf(a);
    }
#pragma unroll
  for (int c = kChD; c < 32; ++c) q[c] = 0.f;
}

// g_z[m] = sum_c d_c * ( g_pe[c] + sum_l 2^l (g_sin_l[c] cos_l[c] - g_cos_l[c] sin_l[c]) )   (x = o + d z)
__global__ void __launch_bounds__(256) pe_backward_kernel(const float* __restrict__ g_pe, const float* __restrict__ pe,
                                                          const float* __restrict__ rays_d, long long M, int S,
                                                          float* __restrict__ g_z) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const long long ray = m / S;
  const float* g = g_pe + m * 64;
  const float* v = pe + m * 64;
  float out = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float gx = g[c];
    for (int l = 0; l < kLx; ++l)
      gx += (float)(1 << l) * (g[3 + 6 * l + c] * v[3 + 6 * l + 3 + c] - g[3 + 6 * l + 3 + c] * v[3 + 6 * l + c]);
    out += gx * rays_d[ray * 3 + c];
  }
  g_z[m] = out;
}

// activation store (floats per row): PE 64 | DPE 32 | H0..H7 8x256 | FEAT 256 | HV 128
constexpr long long kPe = 0, kDpe = 64, kH = 96, kFeat = kH + 8 * 256, kHv = kFeat + 256, kActFloats = kHv + 128;   // 2528
// backward workspace (floats per row): G0 256 | G1 256 | GHV 128 | GPE 64
constexpr long long kWsFloats = 256 + 256 + 128 + 64;

struct Ctx {
  cudaStream_t st;
  long long M;
  int rc;
};

static void gemm(Ctx& c, int M, int N, int K, const float* A, long long sAi, long long sAk, const float* B, long long sBk,
                 long long sBj, float* C, long long ldc, int flags, const float* bias = nullptr, const float* mask = nullptr,
                 long long ldmask = 0, int splits = 1) {
  if (c.rc || M <= 0 || N <= 0) return;
  Gemm g{A, sAi, sAk, B, sBk, sBj, C, ldc, bias, mask, ldmask, M, N, K, flags, 0};
  int kps = (K + splits - 1) / splits;
  kps = ((kps + TK - 1) / TK) * TK;
  if (kps < TK) kps = TK;
  g.k_per_split = kps;
  const int nz = K > 0 ? (K + kps - 1) / kps : 1;
  dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM, nz);
  sgemm_kernel<<<grid, 256, 0, c.st>>>(g);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("launch of sgemm_kernel failed: %s", cudaGetErrorString(e)); c.rc = 3; return; }
  count_launch();
}

// dW[N_out, K_in] (+)= G^T X over the M rows (split-K, atomics into a zeroed tensor), db[N_out] = column sums of G
static void wgrad(Ctx& c, const float* G, long long ldg, int n_out, const float* X, long long ldx, int k_in, float* dW,
                  long long lddw) {
  const long long M = c.M;
  int splits = (int)((M + 4095) / 4096);
  if (splits < 1) splits = 1;
  if (splits > 256) splits = 256;
  gemm(c, n_out, k_in, (int)M, G, 1, ldg, X, ldx, 1, dW, lddw, EPI_ATOMIC, nullptr, nullptr, 0, splits);
}
static void bgrad(Ctx& c, const float* G, long long ldg, int n_out, float* db) {
  if (c.rc) return;
  dim3 grid((n_out + 31) / 32, (unsigned)((c.M + 2047) / 2048));
  colsum_kernel<<<grid, 256, 0, c.st>>>(G, ldg, c.M, n_out, db);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("launch of colsum_kernel failed: %s", cudaGetErrorString(e)); c.rc = 3; return; }
  count_launch();
}

}  // namespace f32t
}  // namespace nb

using namespace nb;
using namespace nb::f32t;

extern "C" size_t nerfb200_train_fp32_acts_bytes(long long n_rows) {
  return n_rows <= 0 ? 0 : (size_t)n_rows * kActFloats * sizeof(float);
}
extern "C" size_t nerfb200_train_fp32_workspace_bytes(long long n_rows) {
  return n_rows <= 0 ? 0 : (size_t)n_rows * kWsFloats * sizeof(float);
}

static int check_weights(const nerfb200_mlp_weights* w, const char* who) {
  NB_CHECK_ARG(w, "%s: null weights", who);
  for (int i = 0; i < 8; ++i) NB_CHECK_ARG(w->pts_w[i] && w->pts_b[i], "%s: null pts_linears.%d", who, i);
  NB_CHECK_ARG(w->views_w && w->views_b && w->feature_w && w->feature_b && w->alpha_w && w->alpha_b && w->rgb_w && w->rgb_b,
               "%s: null head tensor", who);
  return 0;
}

extern "C" int nerfb200_mlp_forward_train_fp32(const nerfb200_mlp_weights* w, const float* rays_o, const float* rays_d,
                                               const float* z_vals, int n_rays, int n_samples, float* raw, void* acts,
                                               void* stream) {
  NB_CHECK_ARG(n_rays >= 0 && n_samples >= 1, "mlp_forward_train_fp32: bad sizes");
  if (n_rays == 0) return 0;
  if (int rc = check_weights(w, "mlp_forward_train_fp32")) return rc;
  NB_CHECK_ARG(rays_o && rays_d && z_vals && raw && acts, "mlp_forward_train_fp32: null pointer");
  NB_CHECK_ARG(((uintptr_t)acts & 15) == 0, "mlp_forward_train_fp32: acts must be 16-byte aligned");
  const long long M = (long long)n_rays * n_samples;
  NB_CHECK_ARG(M < (1LL << 31), "mlp_forward_train_fp32: too many rows");
  Ctx c{(cudaStream_t)stream, M, 0};
  float* a = reinterpret_cast<float*>(acts);
  float* PE = a + kPe * M;
  float* DPE = a + kDpe * M;
  float* H = a + kH * M;          // H[i] = H + i*256*M
  float* FEAT = a + kFeat * M;
  float* HV = a + kHv * M;
  pe_forward_kernel<<<ceil_div(M, 256), 256, 0, c.st>>>(rays_o, rays_d, z_vals, M, n_samples, PE, DPE);
  NB_LAUNCH_OK("pe_forward_kernel");
  const int Mi = (int)M;
  for (int i = 0; i < 8; ++i) {
    float* Hi = H + (long long)i * 256 * M;
    const float* Hp = H + (long long)(i - 1) * 256 * M;
    if (i == 0) {
      gemm(c, Mi, 256, kChX, PE, 64, 1, w->pts_w[0], 1, kChX, Hi, 256, EPI_BIAS | EPI_RELU, w->pts_b[0]);
    } else if (i == kSkip + 1) {   // input = cat(pe, h4), network.py:58-59
      gemm(c, Mi, 256, kChX, PE, 64, 1, w->pts_w[i], 1, kChX + 256, Hi, 256, 0);
      gemm(c, Mi, 256, 256, Hp, 256, 1, w->pts_w[i] + kChX, 1, kChX + 256, Hi, 256, EPI_ACCUM | EPI_BIAS | EPI_RELU, w->pts_b[i]);
    } else {
      gemm(c, Mi, 256, 256, Hp, 256, 1, w->pts_w[i], 1, 256, Hi, 256, EPI_BIAS | EPI_RELU, w->pts_b[i]);
    }
  }
  const float* H7 = H + 7LL * 256 * M;
  gemm(c, Mi, 1, 256, H7, 256, 1, w->alpha_w, 1, 256, raw + 3, 4, EPI_BIAS, w->alpha_b);
  gemm(c, Mi, 256, 256, H7, 256, 1, w->feature_w, 1, 256, FEAT, 256, EPI_BIAS, w->feature_b);
  gemm(c, Mi, kWv, 256, FEAT, 256, 1, w->views_w, 1, 256 + kChD, HV, kWv, 0);
  gemm(c, Mi, kWv, kChD, DPE, 32, 1, w->views_w + 256, 1, 256 + kChD, HV, kWv, EPI_ACCUM | EPI_BIAS | EPI_RELU, w->views_b);
  gemm(c, Mi, 3, kWv, HV, kWv, 1, w->rgb_w, 1, kWv, raw, 4, EPI_BIAS, w->rgb_b);
  return c.rc;
}

extern "C" int nerfb200_mlp_backward_fp32(const nerfb200_mlp_weights* w, const float* g_raw, const void* acts,
                                          const float* rays_d, int n_rays, int n_samples, void* workspace,
                                          size_t workspace_bytes, const nerfb200_mlp_grads* grads, float* g_z,
                                          void* stream) {
  NB_CHECK_ARG(n_rays >= 0 && n_samples >= 1, "mlp_backward_fp32: bad sizes");
  NB_CHECK_ARG(grads, "mlp_backward_fp32: null grads");
  for (int i = 0; i < 8; ++i) NB_CHECK_ARG(grads->pts_w[i] && grads->pts_b[i], "mlp_backward_fp32: null gradient pts_linears.%d", i);
  NB_CHECK_ARG(grads->views_w && grads->views_b && grads->feature_w && grads->feature_b && grads->alpha_w && grads->alpha_b &&
                   grads->rgb_w && grads->rgb_b, "mlp_backward_fp32: null head gradient");
  if (int rc = check_weights(w, "mlp_backward_fp32")) return rc;
  const long long M = (long long)n_rays * n_samples;
  NB_CHECK_ARG(M < (1LL << 31), "mlp_backward_fp32: too many rows");
  cudaStream_t st = (cudaStream_t)stream;
  // gradients are OVERWRITTEN: zero, then accumulate with atomics
  const int in_dim[8] = {kChX, 256, 256, 256, 256, kChX + 256, 256, 256};
  for (int i = 0; i < 8; ++i) {
    NB_CUDA(cudaMemsetAsync(grads->pts_w[i], 0, sizeof(float) * 256 * in_dim[i], st));
    NB_CUDA(cudaMemsetAsync(grads->pts_b[i], 0, sizeof(float) * 256, st));
  }
  NB_CUDA(cudaMemsetAsync(grads->views_w, 0, sizeof(float) * kWv * (256 + kChD), st));
  NB_CUDA(cudaMemsetAsync(grads->views_b, 0, sizeof(float) * kWv, st));
  NB_CUDA(cudaMemsetAsync(grads->feature_w, 0, sizeof(float) * 256 * 256, st));
  NB_CUDA(cudaMemsetAsync(grads->feature_b, 0, sizeof(float) * 256, st));
  NB_CUDA(cudaMemsetAsync(grads->alpha_w, 0, sizeof(float) * 256, st));
  NB_CUDA(cudaMemsetAsync(grads->alpha_b, 0, sizeof(float) * 1, st));
  NB_CUDA(cudaMemsetAsync(grads->rgb_w, 0, sizeof(float) * 3 * kWv, st));
  NB_CUDA(cudaMemsetAsync(grads->rgb_b, 0, sizeof(float) * 3, st));
  if (M == 0) return 0;
  NB_CHECK_ARG(g_raw && acts && workspace && (g_z == nullptr || rays_d), "mlp_backward_fp32: null pointer");
  NB_CHECK_ARG(workspace_bytes >= nerfb200_train_fp32_workspace_bytes(M), "mlp_backward_fp32: workspace too small (%zu < %zu)",
               workspace_bytes, nerfb200_train_fp32_workspace_bytes(M));
  NB_CHECK_ARG(((uintptr_t)acts & 15) == 0 && ((uintptr_t)workspace & 15) == 0, "mlp_backward_fp32: misaligned buffer");
  Ctx c{st, M, 0};
  const int Mi = (int)M;
  const float* a = reinterpret_cast<const float*>(acts);
  const float* PE = a + kPe * M;
  const float* DPE = a + kDpe * M;
  const float* H = a + kH * M;
  const float* FEAT = a + kFeat * M;
  const float* HV = a + kHv * M;
  const float* H7 = H + 7LL * 256 * M;
  float* ws = reinterpret_cast<float*>(workspace);
  float* G0 = ws;
  float* G1 = ws + 256 * M;
  float* GHV = ws + 512 * M;
  float* GPE = ws + 640 * M;
  // rgb_linear (network.py:69): raw[:, :3] = hv W_rgb^T + b
  wgrad(c, g_raw, 4, 3, HV, kWv, kWv, grads->rgb_w, kWv);
  bgrad(c, g_raw, 4, 3, grads->rgb_b);
  gemm(c, Mi, kWv, 3, g_raw, 4, 1, w->rgb_w, kWv, 1, GHV, kWv, EPI_MASK, nullptr, HV, kWv);        // d relu(views)
  // views_linears.0 on cat(feature, dir PE) (network.py:65-67)
  wgrad(c, GHV, kWv, kWv, FEAT, 256, 256, grads->views_w, 256 + kChD);
  wgrad(c, GHV, kWv, kWv, DPE, 32, kChD, grads->views_w + 256, 256 + kChD);
  bgrad(c, GHV, kWv, kWv, grads->views_b);
  gemm(c, Mi, 256, kWv, GHV, kWv, 1, w->views_w, 256 + kChD, 1, G0, 256, 0);                        // d feature
  // feature_linear, alpha_linear on h7 (network.py:62-63)
  wgrad(c, G0, 256, 256, H7, 256, 256, grads->feature_w, 256);
  bgrad(c, G0, 256, 256, grads->feature_b);
  wgrad(c, g_raw + 3, 4, 1, H7, 256, 256, grads->alpha_w, 256);
  bgrad(c, g_raw + 3, 4, 1, grads->alpha_b);
  gemm(c, Mi, 256, 256, G0, 256, 1, w->feature_w, 256, 1, G1, 256, 0);
  gemm(c, Mi, 256, 1, g_raw + 3, 4, 1, w->alpha_w, 256, 1, G1, 256, EPI_ACCUM | EPI_MASK, nullptr, H7, 256);   // d pre7
  float* cur = G1;
  float* oth = G0;
  for (int i = 7; i >= 1; --i) {
    const float* Hp = H + (long long)(i - 1) * 256 * M;
    const int in = in_dim[i];
    const int off = (i == kSkip + 1) ? kChX : 0;
    if (off) wgrad(c, cur, 256, 256, PE, 64, kChX, grads->pts_w[i], in);
    wgrad(c, cur, 256, 256, Hp, 256, 256, grads->pts_w[i] + off, in);
    bgrad(c, cur, 256, 256, grads->pts_b[i]);
    if (off && g_z) gemm(c, Mi, kChX, 256, cur, 256, 1, w->pts_w[i], in, 1, GPE, 64, 0);            // d pe through the skip
    gemm(c, Mi, 256, 256, cur, 256, 1, w->pts_w[i] + off, in, 1, oth, 256, EPI_MASK, nullptr, Hp, 256);
    float* t = cur; cur = oth; oth = t;
  }
  wgrad(c, cur, 256, 256, PE, 64, kChX, grads->pts_w[0], kChX);
  bgrad(c, cur, 256, 256, grads->pts_b[0]);
  if (g_z && !c.rc) {
    gemm(c, Mi, kChX, 256, cur, 256, 1, w->pts_w[0], kChX, 1, GPE, 64, EPI_ACCUM);
    if (c.rc) return c.rc;
    pe_backward_kernel<<<ceil_div(M, 256), 256, 0, st>>>(GPE, PE, rays_d, M, n_samples, g_z);
    NB_LAUNCH_OK("pe_backward_kernel");
  }
  return c.rc;
}
