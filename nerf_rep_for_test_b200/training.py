"""Differentiable ray rendering for training.

Forward = the same CUDA kernels as inference (stratified sampling, fused PE+MLP, compositing, sample_pdf + merge).
Two arithmetic paths, selected with `precision`:
  * "bf16" (performance): tcgen05 forward that leaves every stage's bf16 output in HBM as operand-ready tile images
    plus the relu sign bits (csrc/train_layout.cuh); backward = `nerfb200_mlp_backward` -- a tcgen05
    activation-gradient chain (csrc/mlp_bwd_dgrad.cu) and nine split-K weight-gradient GEMMs with the bias / head
    gradients riding along (csrc/mlp_bwd_wgrad.cu); bf16 operands, fp32 accumulation in TMEM, fp32 gradients.
  * "fp32" (parity): true fp32 FFMA arithmetic layer by layer (csrc/mlp_fp32_train.cu) -- what the reference computes
    under autograd (network.py:49-74, trainer.py:56-60); gradients agree with the reference's autograd to ~1e-5.
Compositing backward: the analytic kernel `nerfb200_composite_backward[_z]` (a7).  No library GEMM is involved.

`ref_compat_sampler`: the reference does NOT detach its hierarchical sampler (volume_renderer.py:181-183 passes
sample_pdf's output, :239-268, straight into the fine pass), so the fine loss also reaches the COARSE network:
dL/dz of the merged depths -- through the fine MLP's input x = o + d z (positional-encoding backward) and through the
fine compositor's interval lengths -- flows through the inverse-CDF sampling (`nerfb200_sample_pdf_backward`) into the
coarse weights and on into the coarse compositor / MLP.  True reproduces that graph (both precisions; the bf16 path gets
the MLP-input gradient from `nerfb200_mlp_backward_input`); False (default) is the original-NeRF semantics with the
sampler detached.

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


def _tensors(model):
    return [p.detach() for p in model_params(model)]


def _forward_passes(r, rays_o, rays_d, precision, z_c=None, _inject=None):
    """Coarse + fine forward with everything the backward needs.  Returns a dict.  _inject (tests only): a dict with
    the coarse weights "w_c" and merged depths "z_all" of another run, used instead of this run's own so that two
    arithmetic paths can be compared at identical sample positions."""
    dev = rays_o.device
    n = rays_o.shape[0]
    if z_c is None:
        r.seed += 1
        z_c = ops.sample_coarse(r._table("z"), n, perturb=float(r.perturb) > 0, seed=r.seed)
    st = {"z_c": z_c, "precision": precision}
    if precision == "bf16":
        raw_c, st["store_c"] = ops.mlp_forward_train(r.packed("coarse", "bf16"), rays_o, rays_d, z_c)
    else:
        raw_c, st["store_c"] = ops.mlp_forward_train_fp32(_tensors(r.coarse_model), rays_o, rays_d, z_c)
    _density_noise(r, raw_c, 1)
    rgb0, disp0, acc0, w_c, depth0 = ops.composite_forward(raw_c, z_c, rays_d, L.COMPOSITE_PLAIN, white_bkgd=r.white_bkgd)
    u = torch.rand((n, r.N_importance), device=dev) if r.net.training else r._table("u")
    z_all = ops.sample_pdf_merge(z_c, w_c, u, want_aux=False)[0]
    if _inject is not None:
        w_c, z_all = _inject["w_c"], _inject["z_all"]
    if precision == "bf16":
        raw_f, st["store_f"] = ops.mlp_forward_train(r.packed("fine", "bf16"), rays_o, rays_d, z_all)
    else:
        raw_f, st["store_f"] = ops.mlp_forward_train_fp32(_tensors(r.fine_model), rays_o, rays_d, z_all)
    _density_noise(r, raw_f, 2)
    rgb, disp, acc, _, depth = ops.composite_forward(raw_f, z_all, rays_d, L.COMPOSITE_PLAIN, white_bkgd=r.white_bkgd,
                                                     want_weights=False)
    st.update(raw_c=raw_c, raw_f=raw_f, w_c=w_c, u=u, z_all=z_all,
              maps=(rgb0, acc0, depth0, disp0, rgb, acc, depth, disp))
    return st


def _mlp_backward(r, which, precision, g_raw, store, rays_o, rays_d, z, grads, want_g_z=False):
    """-> (24 gradients, g_z or None).  g_z = dL/dz through the MLP input (reference graph only)."""
    n, S = z.shape
    if precision == "bf16":
        keep = {} if want_g_z else None
        bwd = r.packed_bwd(which)
        grads = ops.mlp_backward(bwd, g_raw, store, grads=grads, keep_workspace=keep)
        return grads, (ops.mlp_backward_input(bwd, keep["ws"], rays_o, rays_d, z) if want_g_z else None)
    model = r.coarse_model if which == "coarse" else r.fine_model
    return ops.mlp_backward_fp32(_tensors(model), g_raw, store, rays_d, n, S, grads=grads, want_g_z=want_g_z)


def _backward_passes(r, st, rays_o, rays_d, g_coarse, g_fine, ref_compat_sampler, grads_c=None, grads_f=None,
                     after_fine=None):
    """g_coarse / g_fine: (g_rgb, g_acc, g_depth) of the two passes (entries may be None).  Fine pass first: with the
    reference's non-detached sampler its dL/dz feeds the coarse compositor through dL/d(coarse weights)."""
    precision = st["precision"]
    g_w_c = None
    if ref_compat_sampler:
        g_raw_f, g_z = ops.composite_backward_z(st["raw_f"], st["z_all"], rays_d, *g_fine, None, white_bkgd=r.white_bkgd)
        grads_f, g_z_mlp = _mlp_backward(r, "fine", precision, g_raw_f, st["store_f"], rays_o, rays_d, st["z_all"], grads_f,
                                         want_g_z=True)
        g_z += g_z_mlp
        g_w_c = ops.sample_pdf_backward(st["z_c"], st["w_c"], st["u"], g_z)
    else:
        g_raw_f = ops.composite_backward(st["raw_f"], st["z_all"], rays_d, *g_fine, None, white_bkgd=r.white_bkgd)
        grads_f, _ = _mlp_backward(r, "fine", precision, g_raw_f, st["store_f"], rays_o, rays_d, st["z_all"], grads_f)
    if after_fine is not None:
        after_fine()        # the fine model's gradients are final: their all-reduce overlaps the coarse backward
    g_raw_c = ops.composite_backward(st["raw_c"], st["z_c"], rays_d, *g_coarse, g_w_c, white_bkgd=r.white_bkgd)
    grads_c, _ = _mlp_backward(r, "coarse", precision, g_raw_c, st["store_c"], rays_o, rays_d, st["z_c"], grads_c)
    return grads_c, grads_f


class _RenderRays(torch.autograd.Function):
    @staticmethod
    def forward(ctx, renderer, precision, ref_compat_sampler, rays_o, rays_d, *params):
        st = _forward_passes(renderer, rays_o, rays_d, precision)
        rgb0, acc0, depth0, disp0, rgb, acc, depth, disp = st.pop("maps")
        ctx.renderer, ctx.st, ctx.rays_o, ctx.rays_d, ctx.ref_compat_sampler = renderer, st, rays_o, rays_d, ref_compat_sampler
        ctx.mark_non_differentiable(disp0, disp)
        return rgb0, acc0, depth0, disp0, rgb, acc, depth, disp

    @staticmethod
    def backward(ctx, g_rgb0, g_acc0, g_depth0, _gd0, g_rgb, g_acc, g_depth, _gd):
        c = lambda t: None if t is None else t.contiguous()
        grads_c, grads_f = _backward_passes(ctx.renderer, ctx.st, ctx.rays_o, ctx.rays_d, (c(g_rgb0), c(g_acc0), c(g_depth0)),
                                            (c(g_rgb), c(g_acc), c(g_depth)), ctx.ref_compat_sampler)
        ctx.st = None
        return (None, None, None, None, None) + tuple(grads_c) + tuple(grads_f)


_NAMES = sum((["pts_linears.%d.weight" % i, "pts_linears.%d.bias" % i] for i in range(8)), []) + [
    "views_linears.0.weight", "views_linears.0.bias", "feature_linear.weight", "feature_linear.bias",
    "alpha_linear.weight", "alpha_linear.bias", "rgb_linear.weight", "rgb_linear.bias"]


def render_rays_train(renderer, rays_o, rays_d, precision="bf16", ref_compat_sampler=False):
    """Differentiable counterpart of Renderer.render_rays: dict of [N,...] maps carrying grad to the
    parameters of net.model and net.model_fine.  precision / ref_compat_sampler: module docstring."""
    if renderer.enable_ess or renderer.enable_ert:
        raise L.NerfB200Error("training path implements the plain compositor (enable_ess/enable_ert off)")
    if precision not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    params = model_params(renderer.coarse_model) + model_params(renderer.fine_model)
    rays_o = rays_o.to(renderer.device, torch.float32).contiguous()
    rays_d = rays_d.to(renderer.device, torch.float32).contiguous()
    rgb0, acc0, depth0, disp0, rgb, acc, depth, disp = _RenderRays.apply(renderer, precision, bool(ref_compat_sampler),
                                                                         rays_o, rays_d, *params)
    return {"rgb_map_0": rgb0, "acc_map_0": acc0, "depth_map_0": depth0, "disp_map_0": disp0,
            "rgb_map": rgb, "acc_map": acc, "depth_map": depth, "disp_map": disp}


def nerf_loss(out, target_rgb):
    """src/train/trainers/nerf.py:52-65: mse on the coarse and the fine rgb maps."""
    return torch.nn.functional.mse_loss(out["rgb_map_0"], target_rgb) + torch.nn.functional.mse_loss(out["rgb_map"], target_rgb)


def split_adam_state(flat_sd, named_params, offsets):
    """torch.optim.Adam.state_dict() of an optimizer over ONE flat parameter -> the layout the reference's optimizer
    writes (src/train/optimizer.py:14-19: one param group per named parameter, in net.named_parameters() order), so that
    net_utils.save_model / load_model (:288-343) and a stock torch.optim.Adam over those groups read it.
    offsets: {id(param): offset into the flat tensor}."""
    g0 = {k: (float(v) if torch.is_tensor(v) and v.numel() == 1 else v) for k, v in flat_sd["param_groups"][0].items()
          if k != "params"}
    fs = flat_sd["state"].get(0, {})
    state, groups = {}, []
    for i, (name, p) in enumerate(named_params):
        off, n = offsets[id(p)], p.numel()
        if fs:
            state[i] = {"step": fs["step"].detach().clone().cpu() if torch.is_tensor(fs["step"]) else torch.tensor(float(fs["step"])),
                        "exp_avg": fs["exp_avg"][off:off + n].view_as(p).clone(),
                        "exp_avg_sq": fs["exp_avg_sq"][off:off + n].view_as(p).clone()}
        groups.append(dict(g0, params=[i]))
    return {"state": state, "param_groups": groups}


def merge_adam_state(sd, named_params, offsets, n_total, like):
    """Inverse of split_adam_state (also accepts the flat single-tensor layout unchanged): per-parameter Adam state in
    net.named_parameters() order -> state_dict of an Adam over one flat tensor of n_total elements (`like`: a tensor
    giving device / dtype of the moments)."""
    groups = sd["param_groups"]
    named = list(named_params)
    if len(groups) == 1 and len(groups[0]["params"]) == 1 and len(named) != 1:
        return sd                                               # already the flat layout
    order = [i for g in groups for i in g["params"]]
    if len(order) != len(named):
        raise ValueError("optimizer state has %d parameters, the network has %d" % (len(order), len(named)))
    g0 = {k: v for k, v in groups[0].items() if k != "params"}
    state = {}
    if sd["state"]:
        ea = torch.zeros(n_total, dtype=like.dtype, device=like.device)
        es = torch.zeros(n_total, dtype=like.dtype, device=like.device)
        step = None
        for idx, (name, p) in zip(order, named):
            st = sd["state"].get(idx)
            if st is None:
                continue
            off, n = offsets[id(p)], p.numel()
            if tuple(st["exp_avg"].shape) != tuple(p.shape):
                raise ValueError("optimizer state of %s has shape %s, parameter has %s" % (name, tuple(st["exp_avg"].shape), tuple(p.shape)))
            ea[off:off + n] = st["exp_avg"].reshape(-1).to(ea)
            es[off:off + n] = st["exp_avg_sq"].reshape(-1).to(es)
            step = st["step"] if step is None else step
        state[0] = {"step": step if torch.is_tensor(step) else torch.tensor(float(step)), "exp_avg": ea, "exp_avg_sq": es}
    return {"state": state, "param_groups": [dict(g0, params=[0])]}


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

    def __init__(self, renderer, lr=5e-4, graph=False, precision="bf16", ref_compat_sampler=False, fused_update=True):
        from .parallel import FlatGradAllReduce
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.precision, self.ref_compat_sampler = precision, bool(ref_compat_sampler)
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
        self.n_coarse = sum(p.numel() for p in model_params(renderer.coarse_model))
        self.grad_views = {name: [p.grad for p in model_params(m)] for name, m in self.models}
        world = torch.distributed.get_world_size() if (torch.distributed.is_available() and torch.distributed.is_initialized()) else 1
        self.use_graph = bool(graph) and world == 1
        lr_arg = torch.tensor(float(lr), dtype=torch.float32, device=dev) if self.use_graph else lr
        self.opt = torch.optim.Adam([self.flat_param], lr=lr_arg, eps=1e-8, fused=True, capturable=self.use_graph)
        self.allreduce = FlatGradAllReduce(self.params, flat=self.flat)
        self._graph, self._static, self._loss, self._warm = None, None, None, 0
        # fused_update: the loss gradient and clip + Adam run as one kernel of this library each (nerfb200_mse_pair_grad,
        # nerfb200_adam_clip_step) instead of ten torch launches and torch's 19-block multi-tensor Adam; the moments live
        # in self.opt's own state tensors, so state_dict() / load_state_dict() are unchanged.  The graph-captured step
        # keeps torch's capturable Adam (its step counter has to live on the device).
        self.fused_update = bool(fused_update) and not self.use_graph
        self._t = 0

    def _adam_state(self):
        st = self.opt.state[self.flat_param]
        if "exp_avg" not in st:
            st["step"] = torch.zeros((), dtype=torch.float32, device=self.flat_param.device)
            st["exp_avg"] = torch.zeros_like(self.flat_param.data)
            st["exp_avg_sq"] = torch.zeros_like(self.flat_param.data)
        return st

    # ---- checkpoint interface of the optimizer the reference hands to net_utils.save_model / load_model ------------
    def _offsets(self):
        off, out = 0, {}
        for p in self.params:
            out[id(p)] = off
            off += p.numel()
        return out

    def state_dict(self):
        """Adam state in the REFERENCE's layout (one param group per named parameter, net.named_parameters() order,
        src/train/optimizer.py:14-19): pass the TrainStep itself as `optim` to extras.save_model / load_model."""
        if self.fused_update and self._t > 0:
            self._adam_state()["step"].fill_(float(self._t))
        return split_adam_state(self.opt.state_dict(), list(self.r.net.named_parameters()), self._offsets())

    def load_state_dict(self, sd):
        """Accepts a checkpoint's `optim` entry written by the reference (per-parameter groups) or by an older
        TrainStep.opt (one flat tensor)."""
        flat = merge_adam_state(sd, list(self.r.net.named_parameters()), self._offsets(), self.flat_param.numel(), self.flat_param.data)
        lr = self.opt.param_groups[0]["lr"]
        self.opt.load_state_dict(flat)
        if torch.is_tensor(lr):                                 # capturable / graph mode keeps lr in a device tensor
            lr.fill_(float(self.opt.param_groups[0]["lr"]))
            self.opt.param_groups[0]["lr"] = lr
        st = self.opt.state.get(self.flat_param, {})
        self._t = int(round(float(st["step"]))) if "step" in st else 0

    @property
    def param_groups(self):
        return self.opt.param_groups

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
        st = _forward_passes(r, rays_o, rays_d, self.precision, z_c=self._coarse_z(rays_o.shape[0], in_graph))
        rgb0, rgb = st["maps"][0], st["maps"][4]
        # loss = mean((rgb0 - t)^2) + mean((rgb - t)^2)  (trainers/nerf.py:52-65)  ->  dL/d map = 2 (map - t) / (3 n)
        if self.fused_update:
            loss, g0, g1 = ops.mse_pair_grad(rgb0, rgb, target_rgb)
            loss = loss[0]
        else:
            d0, d1 = rgb0 - target_rgb, rgb - target_rgb
            loss = (d0 * d0).mean() + (d1 * d1).mean()
            g0, g1 = d0 * (2.0 / d0.numel()), d1 * (2.0 / d0.numel())
        # gradient all-reduce in two pieces of the flat buffer (coarse parameters first, then fine): the fine piece is
        # started as soon as the fine backward is enqueued and overlaps the coarse network's backward
        handles = []
        _backward_passes(r, st, rays_o, rays_d, (g0, None, None), (g1, None, None), self.ref_compat_sampler,
                         grads_c=self.grad_views["coarse"], grads_f=self.grad_views["fine"],
                         after_fine=lambda: handles.append(self.allreduce.start(self.n_coarse, self.flat.numel())))
        handles.append(self.allreduce.start(0, self.n_coarse))
        if self.fused_update:
            # averaging over the ranks, clip_grad_value_(params, 40) (trainer.py:59) and Adam in one kernel
            inv_world = self.allreduce.finish(handles, scale=False)
            ast, g = self._adam_state(), self.opt.param_groups[0]
            self._t += 1
            ops.adam_clip_step(self.flat_param.data, self.flat, ast["exp_avg"], ast["exp_avg_sq"], g["lr"], g["betas"][0],
                               g["betas"][1], g["eps"], self._t, clip_value=40.0, grad_scale=inv_world)
        else:
            self.allreduce.finish(handles)
            self.flat.clamp_(-40.0, 40.0)          # clip_grad_value_(params, 40) on the aliased buffer
            self.opt.step()
        r.invalidate_weights()
        return loss

    def __call__(self, rays_o, rays_d, target_rgb):
        if self.r.device.index is not None and self.r.device.index != torch.cuda.current_device():
            with torch.cuda.device(self.r.device):      # the C ABI launches on the current device
                return self._call(rays_o, rays_d, target_rgb)
        return self._call(rays_o, rays_d, target_rgb)

    def _call(self, rays_o, rays_d, target_rgb):
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
