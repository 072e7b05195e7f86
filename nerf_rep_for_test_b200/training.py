"""Differentiable ray rendering for training.

Forward = the same CUDA kernels as inference (stratified sampling, fused PE+MLP on tcgen05, compositing,
sample_pdf + merge); the training forward additionally leaves every stage's bf16 output in HBM as
operand-ready tile images plus the relu sign bits (csrc/train_layout.cuh).  Backward:
  * compositing: the analytic kernel `nerfb200_composite_backward` (a7);
  * MLP: `nerfb200_mlp_backward` -- a tcgen05 activation-gradient chain (csrc/mlp_bwd_dgrad.cu) and ten
    split-K weight-gradient GEMMs with the bias / head gradients riding along (csrc/mlp_bwd_wgrad.cu);
    bf16 operands, fp32 accumulation in TMEM, fp32 gradients.  No library GEMM is involved.
  * the hierarchical sampler is detached (original-NeRF semantics).  The reference does not detach
    (volume_renderer.py:181-183), so its fine loss also leaks into the coarse network
    through sample_pdf; that path is not reproduced.

Loss used by the benchmark: mse(rgb_map_0, t) + mse(rgb_map, t) (src/train/trainers/nerf.py:52-65).
"""
import torch

from . import lib as L
from . import ops


def model_params(model):
    """The 24 tensors of one NeRF module in the order nerfb200_mlp_weights lists them."""
    ps = []
    for i in range(8):
        ps += [model.pts_linears[i].weight, model.pts_linears[i].bias]
    ps += [model.views_linears[0].weight, model.views_linears[0].bias, model.feature_linear.weight,
           model.feature_linear.bias, model.alpha_linear.weight, model.alpha_linear.bias,
           model.rgb_linear.weight, model.rgb_linear.bias]
    return ps


def _density_noise(renderer, raw, which):
    """raw_noise_std (volume_renderer.py:310-314): sigma_raw += N(0,1)*std, in place on the MLP output, so the
    compositing forward and its analytic backward see the same noisy density."""
    std = float(renderer.raw_noise_std or 0.0)
    if std > 0.0:
        ops.sigma_noise(raw, std, seed=renderer.seed * 0x9E3779B97F4A7C15 + which)


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
        raw_c, store_c = ops.mlp_forward_train(pk_c, rays_o, rays_d, z_c)
        _density_noise(r, raw_c, 1)
        rgb0, disp0, acc0, w_c, depth0 = ops.composite_forward(raw_c, z_c, rays_d, L.COMPOSITE_PLAIN,
                                                               white_bkgd=r.white_bkgd)
        u = torch.rand((n, U), device=dev) if r.net.training else r._table("u")
        z_all, _, _, _ = ops.sample_pdf_merge(z_c, w_c, u, want_aux=False)
        raw_f, store_f = ops.mlp_forward_train(pk_f, rays_o, rays_d, z_all)
        _density_noise(r, raw_f, 2)
        rgb, disp, acc, _, depth = ops.composite_forward(raw_f, z_all, rays_d, L.COMPOSITE_PLAIN,
                                                         white_bkgd=r.white_bkgd, want_weights=False)
        ctx.renderer = r
        ctx.stores = (store_c, store_f)
        ctx.save_for_backward(rays_d, z_c, z_all, raw_c, raw_f)
        ctx.mark_non_differentiable(disp0, disp)
        return rgb0, acc0, depth0, disp0, rgb, acc, depth, disp

    @staticmethod
    def backward(ctx, g_rgb0, g_acc0, g_depth0, _gd0, g_rgb, g_acc, g_depth, _gd):
        r = ctx.renderer
        rays_d, z_c, z_all, raw_c, raw_f = ctx.saved_tensors
        c = lambda t: None if t is None else t.contiguous()
        grads = []
        for which, z, raw, store, gr, ga, gd in (("coarse", z_c, raw_c, ctx.stores[0], g_rgb0, g_acc0, g_depth0),
                                                 ("fine", z_all, raw_f, ctx.stores[1], g_rgb, g_acc, g_depth)):
            g_raw = ops.composite_backward(raw, z, rays_d, c(gr), c(ga), c(gd), None, white_bkgd=r.white_bkgd)
            grads += ops.mlp_backward(r.packed_bwd(which), g_raw, store)
        ctx.stores = None
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
    (trainer.py:59) + Adam(lr 5e-4, eps 1e-8) (src/train/optimizer.py:8-28).

    The step calls the kernels directly instead of going through autograd (render_rays_train above stays the
    autograd entry for generic use): the loss is mse(rgb_map_0, t) + mse(rgb_map, t), so dL/d(rgb map) is an
    elementwise expression, and nerfb200_mlp_backward writes every gradient straight into its slice of ONE flat
    fp32 buffer that the parameters' .grad tensors alias -- the NCCL all-reduce, the clip and the fused Adam
    all work on that buffer with no per-parameter copies.

    graph=True (single process only): after three eager warm-up steps the whole step -- weight re-packs, sampling,
    both MLP forwards with their activation stores, compositing, loss, both backward chains, clip and Adam -- is
    captured into ONE CUDA graph and replayed.  Measured on B200: 213 / 208 / 206 it/s against 209 / 211 / 210 it/s
    eager on boxes of the same pool, i.e. no gain -- the eager step's host side is three times ahead of the device
    (1.5 ms of enqueue for 4.7 ms of kernels), and the ~0.3 ms the step spends outside its kernels are dependency
    bubbles between ~50 short kernels that a graph does not remove.  Kept as an opt-in (slow hosts).  Everything random inside the
    graph comes from torch's graph-aware generator: the stratified jitter is torch.rand through the reference's own
    formula (:228-235) instead of the kernel's hash, u is torch.rand as before.  The learning rate lives in a device
    tensor (set_lr); raw_noise_std > 0 falls back to the eager step (its seed is a kernel argument).  Inputs must keep
    their shape; a capture failure falls back to the eager step with a warning."""

    def __init__(self, renderer, lr=5e-4, graph=False):
        from .parallel import FlatGradAllReduce
        self.r = renderer
        self.models = (("coarse", renderer.coarse_model), ("fine", renderer.fine_model))
        self.params = [p for _, m in self.models for p in model_params(m)]
        dev = self.params[0].device
        n_total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n_total, dtype=torch.float32, device=dev)
        # The parameters themselves are re-pointed at slices of one flat fp32 buffer as well (names, shapes and values of
        # the state_dict are unchanged), and Adam runs on that single tensor: torch's fused Adam over the 48 small
        # tensors took 84 + 48 us per step (multi_tensor_apply chunks), over one 1.19 M-element tensor it is one
        # short launch.  The update rule is elementwise, so the result is identical.
        self.flat_param = torch.nn.Parameter(torch.empty(n_total, dtype=torch.float32, device=dev), requires_grad=True)
        off = 0
        with torch.no_grad():
            for p in self.params:
                n = p.numel()
                self.flat_param.data[off:off + n].copy_(p.data.reshape(-1))
                p.data = self.flat_param.data[off:off + n].view_as(p)
                p.grad = self.flat[off:off + n].view_as(p)
                off += n
        self.flat_param.grad = self.flat
        self.grad_views = {name: [p.grad for p in model_params(m)] for name, m in self.models}
        world = torch.distributed.get_world_size() if (torch.distributed.is_available() and torch.distributed.is_initialized()) else 1
        self.use_graph = bool(graph) and world == 1
        lr_arg = torch.tensor(float(lr), dtype=torch.float32, device=dev) if self.use_graph else lr
        self.opt = torch.optim.Adam([self.flat_param], lr=lr_arg, eps=1e-8, fused=True, capturable=self.use_graph)
        self.allreduce = FlatGradAllReduce(self.params, flat=self.flat)
        self._graph, self._static, self._loss, self._warm = None, None, None, 0

    def set_lr(self, lr):
        for g in self.opt.param_groups:
            if torch.is_tensor(g["lr"]):
                g["lr"].fill_(float(lr))
            else:
                g["lr"] = float(lr)

    def _coarse_z(self, n, in_graph):
        r = self.r
        jitter = float(r.perturb) > 0
        if not (in_graph and jitter):
            r.seed += 1
            return ops.sample_coarse(r._table("z"), n, perturb=jitter, seed=r.seed)
        tab = r._table("z")                                   # :228-235 with torch's generator
        mids = 0.5 * (tab[1:] + tab[:-1])
        upper, lower = torch.cat([mids, tab[-1:]]), torch.cat([tab[:1], mids])
        return (lower + (upper - lower) * torch.rand((n, tab.numel()), device=tab.device)).contiguous()

    def _step(self, rays_o, rays_d, target_rgb, in_graph=False):
        r = self.r
        dev = r.device
        n = rays_o.shape[0]
        z_c = self._coarse_z(n, in_graph)
        pk_c, pk_f = r.packed("coarse", "bf16"), r.packed("fine", "bf16")
        raw_c, store_c = ops.mlp_forward_train(pk_c, rays_o, rays_d, z_c)
        _density_noise(r, raw_c, 1)
        rgb0, _, _, w_c, _ = ops.composite_forward(raw_c, z_c, rays_d, L.COMPOSITE_PLAIN, white_bkgd=r.white_bkgd)
        u = torch.rand((n, r.N_importance), device=dev) if r.net.training else r._table("u")
        z_all = ops.sample_pdf_merge(z_c, w_c, u, want_aux=False)[0]
        raw_f, store_f = ops.mlp_forward_train(pk_f, rays_o, rays_d, z_all)
        _density_noise(r, raw_f, 2)
        rgb = ops.composite_forward(raw_f, z_all, rays_d, L.COMPOSITE_PLAIN, white_bkgd=r.white_bkgd, want_weights=False)[0]
        # loss = mean((rgb0 - t)^2) + mean((rgb - t)^2)  (trainers/nerf.py:52-65)  ->  dL/d map = 2 (map - t) / (3 n)
        d0, d1 = rgb0 - target_rgb, rgb - target_rgb
        loss = (d0 * d0).mean() + (d1 * d1).mean()
        scale = 2.0 / d0.numel()
        for which, z, raw, store, d in (("coarse", z_c, raw_c, store_c, d0), ("fine", z_all, raw_f, store_f, d1)):
            g_raw = ops.composite_backward(raw, z, rays_d, d * scale, None, None, None, white_bkgd=r.white_bkgd)
            ops.mlp_backward(r.packed_bwd(which), g_raw, store, grads=self.grad_views[which])
        self.allreduce()
        self.flat.clamp_(-40.0, 40.0)          # clip_grad_value_(params, 40) on the aliased buffer
        self.opt.step()
        r.invalidate_weights()
        return loss

    def __call__(self, rays_o, rays_d, target_rgb):
        r = self.r
        if r.enable_ess or r.enable_ert:
            raise L.NerfB200Error("training path implements the plain compositor (enable_ess/enable_ert off)")
        dev = r.device
        rays_o = rays_o.to(dev, torch.float32).contiguous()
        rays_d = rays_d.to(dev, torch.float32).contiguous()
        target_rgb = target_rgb.to(dev, torch.float32).contiguous()
        if not self.use_graph or float(r.raw_noise_std or 0.0) > 0.0:
            return self._step(rays_o, rays_d, target_rgb)
        if self._static is None or self._static[0].shape != rays_o.shape:
            self._static = (torch.empty_like(rays_o), torch.empty_like(rays_d), torch.empty_like(target_rgb))
            self._graph, self._warm = None, 0
        for dst, src in zip(self._static, (rays_o, rays_d, target_rgb)):
            dst.copy_(src)
        if self._graph is None:
            if self._warm < 3:                 # eager warm-up: kernel attributes, allocator pools, Adam state
                self._warm += 1
                return self._step(*self._static)
            try:
                r.invalidate_weights()         # the re-packs must be part of the captured step
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._loss = self._step(*self._static, in_graph=True)
                self._graph = g
            except Exception as e:             # capture is an optimisation, never a requirement
                import warnings
                warnings.warn("TrainStep: CUDA-graph capture failed (%s); running the eager step" % (e,))
                self.use_graph = False
                torch.cuda.synchronize()
                r.invalidate_weights()
                return self._step(rays_o, rays_d, target_rgb)
        self._graph.replay()
        r.invalidate_weights()
        return self._loss
