// a7, the small end of the training step (trainers/nerf.py:52-65, trainer.py:56-60, optimizer.py:8-28): the loss
// gradient of mse(rgb_map_0, t) + mse(rgb_map, t) and the clip_grad_value_ + Adam update, each as ONE kernel over the
// flat buffers TrainStep keeps.  torch ran the former as ten elementwise / reduction launches (34 us of GPU time) and the
// latter as a 19-block multi_tensor_apply launch (one 65 536-element chunk per block: 80 us for 1.19 M parameters,
// latency bound) -- 2.5 % of a 4.5 ms step.
#include "common.cuh"

namespace nb {

// g0 = 2 (rgb0 - t) / (3n), g1 = 2 (rgb - t) / (3n); loss[0] += sum((rgb0 - t)^2 + (rgb - t)^2) / (3n)   (loss zeroed by the caller)
__global__ void __launch_bounds__(256) mse_pair_grad_kernel(const float* __restrict__ rgb0, const float* __restrict__ rgb,
                                                            const float* __restrict__ target, int count, float inv_count,
                                                            float* __restrict__ g0, float* __restrict__ g1,
                                                            float* __restrict__ loss) {
  float part = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    const float t = target[i];
    const float d0 = rgb0[i] - t, d1 = rgb[i] - t;
    g0[i] = 2.f * inv_count * d0;
    g1[i] = 2.f * inv_count * d1;
    part += d0 * d0 + d1 * d1;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    atomicAdd(loss, s * inv_count);
  }
}

// torch.optim.Adam (amsgrad = False, weight_decay = 0, maximize = False) with clip_grad_value_ in front, the arithmetic
// of torch's fused CUDA implementation: exp_avg = lerp(exp_avg, g, 1 - beta1); exp_avg_sq = beta2 exp_avg_sq + (1 - beta2) g g;
// p -= (lr / bc1) * exp_avg / (sqrt(exp_avg_sq) / sqrt(bc2) + eps).  g is first scaled (gradient averaging) and clamped
// to [-clip, clip]; the clamped value is written back so that .grad reads as after clip_grad_value_.
__global__ void __launch_bounds__(256) adam_clip_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, long long n, float lr_over_bc1, float one_minus_beta1,
                                                        float beta2, float one_minus_beta2, float eps, float inv_sqrt_bc2, float clip,
                                                        float grad_scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * grad_scale;
    if (clip > 0.f) gi = fminf(fmaxf(gi, -clip), clip);
    // 1 - beta is formed in double on the host as torch does (1.f - 0.999f is 4.7e-5 off 0.001)
    const float mi = m[i] + one_minus_beta1 * (gi - m[i]);
    const float vi = beta2 * v[i] + one_minus_beta2 * gi * gi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] -= lr_over_bc1 * (mi / denom);
    g[i] = gi; m[i] = mi; v[i] = vi;
  }
}

}  // namespace nb

using namespace nb;

extern "C" int nerfb200_mse_pair_grad(const float* rgb0, const float* rgb, const float* target, int n_rays, float* g_rgb0,
                                      float* g_rgb, float* loss, void* stream) {
  NB_CHECK_ARG(n_rays >= 0, "mse_pair_grad: negative n_rays");
  NB_CHECK_ARG(loss, "mse_pair_grad: null loss");
  NB_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), (cudaStream_t)stream));
  if (n_rays == 0) return 0;
  NB_CHECK_ARG(rgb0 && rgb && target && g_rgb0 && g_rgb, "mse_pair_grad: null pointer");
  const int count = n_rays * 3;
  int blocks = ceil_div(count, 256);
  if (blocks > 296) blocks = 296;
  mse_pair_grad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(rgb0, rgb, target, count, 1.0f / (float)count, g_rgb0, g_rgb, loss);
  NB_LAUNCH_OK("mse_pair_grad_kernel");
  return 0;
}

extern "C" int nerfb200_adam_clip_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, long long n, double lr,
                                       double beta1, double beta2, double eps, long long step, float clip_value, float grad_scale,
                                       void* stream) {
  NB_CHECK_ARG(n >= 0 && step >= 1, "adam_clip_step: bad n / step (step counts from 1)");
  NB_CHECK_ARG(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0, "adam_clip_step: bad hyper-parameters");
  if (n == 0) return 0;
  NB_CHECK_ARG(params && grads && exp_avg && exp_avg_sq, "adam_clip_step: null pointer");
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  int blocks = ceil_div(n, 256 * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  adam_clip_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, (float)(lr / bc1),
                                                            (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps,
                                                            (float)(1.0 / sqrt(bc2)), clip_value, grad_scale);
  NB_LAUNCH_OK("adam_clip_kernel");
  return 0;
}
