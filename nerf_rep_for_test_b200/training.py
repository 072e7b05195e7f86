"""Differentiable ray rendering for training.

Forward = the same CUDA kernels as inference (stratified sampling, fused PE+MLP on tcgen05 with the
stage activations kept in bf16, compositing, sample_pdf + merge).  Backward:
  * compositing: the analytic kernel `nerfb200_composite_backward` (a7);
  * MLP: dgrad / wgrad of the ten dense stages as plain bf16 GEMMs with fp32 accumulation through
    torch.matmul (cuBLAS) on the saved activations -- ROUND-1 STATUS: these are library GEMMs, not yet
    a fused tcgen05 backward kernel; everything else on the path is ours.
  * the hierarchical sampler is detached (original-NeRF semantics).  The reference does not detach
    (volume_renderer.py:181-183), so its fine loss also leaks into the coarse network
    through sample_pdf; that path is not reproduced.

Loss used by the benchmark: mse(rgb_map_0, t) + mse(rgb_map, t) (src/train/trainers/nerf.py:52-65).
"""
import torch

from . import lib as L
from . import ops

CH_XYZ, CH_DIR = 63, 27


def _pos_enc(x, n_freqs):
    """freq.py:23-26 ordering; only used to rebuild the (non-differentiated) MLP inputs for wgrad."""
    out = [x]
    for f in range(n_freqs):
        out.append(torch.sin(x * float(2 ** f)))
        out.append(torch.cos(x * float(2 ** f)))
    return torch.cat(out, -1)


def model_params(model):
    """The 24 tensors of one NeRF module in the order nerfb200_mlp_weights lists them."""
    ps = []
    for i in range(8):
        ps += [model.pts_linears[i].weight, model.pts_linears[i].bias]
    ps += [model.views_linears[0].weight, model.views_linears[0].bias, model.feature_linear.weight,
           model.feature_linear.bias, model.alpha_linear.weight, model.alpha_linear.bias,
           model.rgb_linear.weight, model.rgb_linear.bias]
    return ps


def _mm_f32(a, b):
    """a @ b with fp32 accumulation AND fp32 output (weight gradients)."""
    if a.dtype == torch.float32:
        return torch.mm(a, b)
    try:
        return torch.mm(a, b, out_dtype=torch.float32)
    except TypeError:      # older torch: bf16 output, widened afterwards
        return torch.mm(a, b).float()


def _pad_cols(t, n):
    """zero-pad the last dim to n columns (tensor-core GEMM kernels want multiples of 8 bf16 elements;
    K = 3 / 27 / 63 fall back to slow paths otherwise)."""
    if t.shape[-1] == n:
        return t.contiguous()
    out = t.new_zeros(t.shape[:-1] + (n,))
    out[..., : t.shape[-1]] = t
    return out


def mlp_backward(params, acts, pe, dpe, g_raw, compute_dtype=torch.bfloat16):
    """Gradients of the 24 tensors given dL/draw.

    params: list from model_params(); acts: [10, M, 256] stage outputs (relu(h0..h7), feature, relu(views));
    pe [M,63], dpe [M,27]: MLP inputs; g_raw [M,4] = dL/d(rgb_raw, sigma_raw).
    Activation gradients and GEMM operands are kept in `compute_dtype` (bf16: one [M,256] tensor is
    0.5 GB at 4096 rays x 256 samples); GEMMs accumulate in fp32; weight/bias gradients come out fp32.
    """
    cd = compute_dtype
    W = [p.detach().to(cd) for p in params]
    acts = acts.to(cd)
    pe, dpe = _pad_cols(pe.to(cd), 64), _pad_cols(dpe.to(cd), 32)
    g8 = _pad_cols(g_raw.to(cd), 8)                       # [M,8]: rgb(3) sigma(1) 0000
    h = [acts[i] for i in range(8)]                       # post-relu hidden activations
    feat, hv = acts[8], acts[9][:, :128]
    relu_bwd = torch.ops.aten.threshold_backward          # grad * (act > 0), one pass
    colsum = lambda t: t.sum(0, dtype=torch.float32)
    grads = [None] * 24
    # heads: raw = [hv @ Wrgb^T + b_rgb | h7 @ w_alpha^T + b_alpha]
    g_hv = _mm_f32(g8.t(), hv)                            # [8,128]: rows 0-2 = dW_rgb
    g_h7 = _mm_f32(g8.t(), h[7])                          # [8,256]: row 3 = dw_alpha
    gsum = colsum(g8)
    grads[22], grads[23] = g_hv[:3].contiguous(), gsum[:3].contiguous()
    grads[20], grads[21] = g_h7[3:4].contiguous(), gsum[3:4].contiguous()
    w_heads_hv = torch.zeros(8, 128, dtype=cd, device=g8.device)
    w_heads_hv[:3] = W[22]
    w_heads_h7 = torch.zeros(8, 256, dtype=cd, device=g8.device)
    w_heads_h7[3] = W[20][0]
    d_hv = relu_bwd(torch.mm(g8, w_heads_hv), hv, 0.0)
    # views_linears.0 on [feature | dpe]
    grads[16] = torch.cat([_mm_f32(d_hv.t(), feat), _mm_f32(d_hv.t(), dpe)[:, :CH_DIR]], 1)
    grads[17] = colsum(d_hv)
    d_feat = torch.mm(d_hv, W[16][:, :256].contiguous())
    # feature_linear (linear, no activation) on h7
    grads[18] = _mm_f32(d_feat.t(), h[7])
    grads[19] = colsum(d_feat)
    d_h = torch.addmm(torch.mm(g8, w_heads_h7), d_feat, W[18])   # d_feat @ Wf + g_sigma * w_alpha
    for i in range(7, -1, -1):
        d_pre = relu_bwd(d_h, h[i], 0.0)
        if i == 0:
            grads[0] = _mm_f32(d_pre.t(), pe)[:, :CH_XYZ].contiguous()
        elif i == 5:                                      # skip concat: input = [pe | h4]  (network.py:57-58)
            grads[10] = torch.cat([_mm_f32(d_pre.t(), pe)[:, :CH_XYZ], _mm_f32(d_pre.t(), h[4])], 1)
        else:
            grads[2 * i] = _mm_f32(d_pre.t(), h[i - 1])
        grads[2 * i + 1] = colsum(d_pre)
        if i > 0:
            d_h = torch.mm(d_pre, W[2 * i][:, CH_XYZ:].contiguous() if i == 5 else W[2 * i])
    return grads


class _RenderRays(torch.autograd.Function):
    @staticmethod
    def forward(ctx, renderer, rays_o, rays_d, *params):
        r = renderer
        dev = rays_o.device
        n = rays_o.shape[0]
        S, U = r.N_samples, r.N_importance
        r.seed += 1
        z_c = ops.sample_coarse(r._table("z"), n, perturb=float(r.perturb) > 0, seed=r.seed)
        pk_c, pk_f = r.packed("coarse", "bf16"), r.packed("fine", "bf16")   # cached per parameter version
        raw_c, acts_c = ops.mlp_forward_train(pk_c, rays_o, rays_d, z_c)
        rgb0, disp0, acc0, w_c, depth0 = ops.composite_forward(raw_c, z_c, rays_d, L.COMPOSITE_PLAIN,
                                                               white_bkgd=r.white_bkgd)
        u = torch.rand((n, U), device=dev) if r.net.training else r._table("u")
        z_all, _, _, _ = ops.sample_pdf_merge(z_c, w_c, u, want_aux=False)
        raw_f, acts_f = ops.mlp_forward_train(pk_f, rays_o, rays_d, z_all)
        rgb, disp, acc, _, depth = ops.composite_forward(raw_f, z_all, rays_d, L.COMPOSITE_PLAIN,
                                                         white_bkgd=r.white_bkgd, want_weights=False)
        ctx.renderer = r
        ctx.save_for_backward(rays_o, rays_d, z_c, z_all, raw_c, raw_f, acts_c, acts_f, *params)
        ctx.mark_non_differentiable(disp0, disp)
        return rgb0, acc0, depth0, disp0, rgb, acc, depth, disp

    @staticmethod
    def backward(ctx, g_rgb0, g_acc0, g_depth0, _gd0, g_rgb, g_acc, g_depth, _gd):
        r = ctx.renderer
        rays_o, rays_d, z_c, z_all, raw_c, raw_f, acts_c, acts_f = ctx.saved_tensors[:8]
        params = ctx.saved_tensors[8:]
        c = lambda t: None if t is None else t.contiguous()
        grads = []
        for z, raw, acts, ps, gr, ga, gd in ((z_c, raw_c, acts_c, params[:24], g_rgb0, g_acc0, g_depth0),
                                            (z_all, raw_f, acts_f, params[24:], g_rgb, g_acc, g_depth)):
            g_raw = ops.composite_backward(raw, z, rays_d, c(gr), c(ga), c(gd), None, white_bkgd=r.white_bkgd)
            pts = (rays_o[:, None, :] + rays_d[:, None, :] * z[..., None]).reshape(-1, 3)
            dirs = rays_d[:, None, :].expand(z.shape[0], z.shape[1], 3).reshape(-1, 3)
            grads += mlp_backward(ps, acts, _pos_enc(pts, 10), _pos_enc(dirs, 4), g_raw.reshape(-1, 4))
        return (None, None, None) + tuple(grads)


_NAMES = sum((["pts_linears.%d.weight" % i, "pts_linears.%d.bias" % i] for i in range(8)), []) + [
    "views_linears.0.weight", "views_linears.0.bias", "feature_linear.weight", "feature_linear.bias",
    "alpha_linear.weight", "alpha_linear.bias", "rgb_linear.weight", "rgb_linear.bias"]


def render_rays_train(renderer, rays_o, rays_d):
    """Differentiable counterpart of Renderer.render_rays: dict of [N,...] maps carrying grad to the
    parameters of net.model and net.model_fine."""
    if renderer.enable_ess or renderer.enable_ert:
        raise L.NerfB200Error("training path implements the plain compositor (enable_ess/enable_ert off)")
    params = model_params(renderer.coarse_model) + model_params(renderer.fine_model)
    rays_o = rays_o.to(renderer.device, torch.float32).contiguous()
    rays_d = rays_d.to(renderer.device, torch.float32).contiguous()
    rgb0, acc0, depth0, disp0, rgb, acc, depth, disp = _RenderRays.apply(renderer, rays_o, rays_d, *params)
    return {"rgb_map_0": rgb0, "acc_map_0": acc0, "depth_map_0": depth0, "disp_map_0": disp0,
            "rgb_map": rgb, "acc_map": acc, "depth_map": depth, "disp_map": disp}


def nerf_loss(out, target_rgb):
    """src/train/trainers/nerf.py:52-65: mse on the coarse and the fine rgb maps."""
    return torch.nn.functional.mse_loss(out["rgb_map_0"], target_rgb) + torch.nn.functional.mse_loss(out["rgb_map"], target_rgb)


class TrainStep:
    """One data-parallel training step: fwd + bwd + flat-gradient all-reduce + clip_grad_value_(40)
    (trainer.py:59) + Adam(lr 5e-4, eps 1e-8) (src/train/optimizer.py:8-28)."""

    def __init__(self, renderer, lr=5e-4):
        from .parallel import FlatGradAllReduce
        self.r = renderer
        self.params = list(renderer.net.model.parameters()) + list(renderer.net.model_fine.parameters())
        self.opt = torch.optim.Adam(self.params, lr=lr, eps=1e-8)
        self.allreduce = FlatGradAllReduce(self.params)

    def __call__(self, rays_o, rays_d, target_rgb):
        self.opt.zero_grad(set_to_none=True)
        out = render_rays_train(self.r, rays_o, rays_d)
        loss = nerf_loss(out, target_rgb)
        loss.backward()
        self.allreduce()
        torch.nn.utils.clip_grad_value_(self.params, 40)
        self.opt.step()
        return loss.detach()
