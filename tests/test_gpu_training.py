"""GPU: gradients of the training path against autograd through the CPU oracle (fp32), and one
optimisation smoke run."""
import pytest
import torch

from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from nerf_rep_for_test_b200 import Network, RenderConfig, Renderer
    from nerf_rep_for_test_b200 import training as T
    DEV = torch.device("cuda:0")


def _setup(seed=3, gain=30.0, bias=0.2, n=96):
    sd = O.make_state_dict(seed, gain, bias)
    net = Network(device=DEV)
    net.load_state_dict(sd)
    net.to(DEV).eval()
    r = Renderer(net, RenderConfig(perturb=0, enable_ess=False, enable_ert=False), mode="bf16")
    b = O.lego_batch(32, 32)
    ro, rd = O.get_rays(32, 32, b["pose"][0], b["intrinsics"][0])
    sel = torch.randperm(ro.shape[0], generator=torch.Generator().manual_seed(0))[:n]
    target = torch.rand(n, 3, generator=torch.Generator().manual_seed(1))
    return sd, net, r, ro[sel].contiguous(), rd[sel].contiguous(), target


def _oracle_grads(sd, ro, rd, target, z_all=None):
    """fp32 autograd through the oracle with the sampler detached.  With `z_all` given, the fine pass
    is evaluated at exactly those sample positions: sin(2^9 p) turns a 1e-3 shift of a fine sample
    (caused by bf16 coarse weights) into a 0.5 rad phase change, which would decorrelate the layer-0
    weight gradients of the two pipelines without saying anything about the backward itself."""
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    if z_all is None:
        out = O.render_rays(sdg, ro, rd, detach_sampler=True)
    else:
        z_c = O.sample_coarse(ro.shape[0])
        raw_c = O.query_network(sdg, "model.", ro[:, None, :] + rd[:, None, :] * z_c[..., None], rd)
        rgb0 = O.raw2outputs(raw_c, z_c, rd)[0]
        raw_f = O.query_network(sdg, "model_fine.", ro[:, None, :] + rd[:, None, :] * z_all[..., None], rd)
        out = {"rgb_map_0": rgb0, "rgb_map": O.raw2outputs(raw_f, z_all, rd)[0]}
    loss = torch.nn.functional.mse_loss(out["rgb_map_0"], target) + torch.nn.functional.mse_loss(out["rgb_map"], target)
    loss.backward()
    return float(loss.detach()), {k: v.grad for k, v in sdg.items()}


def test_training_gradients_vs_oracle_autograd():
    sd, net, r, ro, rd, target = _setup()
    # the fine sample positions our (bf16) forward uses, rebuilt with the same kernels
    from nerf_rep_for_test_b200 import lib as L, ops
    z_c = ops.sample_coarse(r._table("z"), ro.shape[0])
    raw_c, _ = ops.mlp_forward_train(r.packed("coarse"), ro.to(DEV), rd.to(DEV), z_c)   # the training forward
    w_c = ops.composite_forward(raw_c, z_c, rd.to(DEV))[3]
    z_all = ops.sample_pdf_merge(z_c, w_c, r._table("u"), want_aux=False)[0].cpu()
    loss_ref, gref = _oracle_grads(sd, ro, rd, target, z_all)
    out = T.render_rays_train(r, ro.to(DEV), rd.to(DEV))
    loss = T.nerf_loss(out, target.to(DEV))
    loss.backward()
    assert abs(float(loss) - loss_ref) < 2e-3 * max(1.0, abs(loss_ref))
    worst = 1.0
    for name, p in net.state_dict(keep_vars=True).items():
        g, ref = p.grad.float().cpu(), gref[name]
        cos = float((g * ref).sum() / (g.norm() * ref.norm() + 1e-20))
        rel = float((g - ref).norm() / (ref.norm() + 1e-20))
        worst = min(worst, cos)
        print("%-36s cos %.5f  rel %.4f" % (name, cos, rel))
        # bf16 operands in forward and backward GEMMs: a few 1e-2 relative on the gradient norm
        assert cos > 0.995 and rel < 0.1, (name, cos, rel)
    print("worst gradient cosine vs fp32 reference autograd: %.5f" % worst)


def test_train_steps_reduce_loss():
    sd, net, r, ro, rd, target = _setup(seed=5, gain=10.0, bias=0.0, n=256)
    net.train()
    step = T.TrainStep(r)
    r.perturb = 1
    losses = [float(step(ro.to(DEV), rd.to(DEV), target.to(DEV))) for _ in range(30)]
    assert losses[-1] < 0.7 * losses[0], losses[::5]
    assert all(torch.isfinite(p).all() for p in net.parameters())


def test_training_with_density_noise():
    """raw_noise_std > 0 (volume_renderer.py:310-314) in the training path: the noise is added in place to the
    MLP output, so forward and analytic backward see the same density; maps and gradients stay finite, differ
    from the noise-free step, and the optimisation still reduces the loss."""
    sd, net, r, ro, rd, target = _setup(seed=5, gain=10.0, bias=0.0, n=256)
    out0 = T.render_rays_train(r, ro.to(DEV), rd.to(DEV))
    T.nerf_loss(out0, target.to(DEV)).backward()
    g0 = net.model_fine.alpha_linear.weight.grad.clone()
    net.zero_grad()
    r.raw_noise_std = 1.0
    out1 = T.render_rays_train(r, ro.to(DEV), rd.to(DEV))
    T.nerf_loss(out1, target.to(DEV)).backward()
    g1 = net.model_fine.alpha_linear.weight.grad.clone()
    assert torch.isfinite(out1["rgb_map"]).all() and torch.isfinite(g1).all()
    assert float((out1["acc_map"] - out0["acc_map"]).abs().max()) > 1e-3
    assert float((g1 - g0).norm()) > 1e-3 * float(g0.norm())
    net.zero_grad()
    net.train()
    r.perturb = 1
    r.raw_noise_std = 0.2     # sigma_raw of this fixture is ~0.2: std 1.0 drowns it (measured: loss 0.183 -> 0.163 in 30 steps)
    step = T.TrainStep(r)
    losses = [float(step(ro.to(DEV), rd.to(DEV), target.to(DEV))) for _ in range(30)]
    assert losses[-1] < 0.95 * losses[0], losses[::5]
    assert all(torch.isfinite(p).all() for p in net.parameters())


def test_graph_captured_step_equals_eager_step():
    """TrainStep(graph=True): after three eager warm-up steps the whole step is one CUDA graph.  With jitter off and
    the network in eval mode (u table) the only run-to-run difference of a step is the order of the split-K
    `red.global.add`s in wgrad (two EAGER runs already differ in the 8th digit of the loss), so graph replays must
    track the eager steps to that level -- parameters and losses -- including the weight re-packs inside the graph."""
    outs = []
    for graph in (False, True):
        sd, net, r, ro, rd, target = _setup(seed=5, gain=10.0, bias=0.0, n=256)
        net.eval()
        r.perturb = 0
        step = T.TrainStep(r, graph=graph, fused_update=False)     # same optimizer kernel on both sides (torch's fused Adam)
        losses = [float(step(ro.to(DEV), rd.to(DEV), target.to(DEV))) for _ in range(8)]
        if graph:
            assert step._graph is not None, "capture fell back to the eager step"
        outs.append((losses, [p.detach().clone() for p in net.parameters()]))
    for le, lg in zip(*[o[0] for o in outs]):
        assert abs(le - lg) <= 2e-5 * max(abs(le), 1e-3), (outs[0][0], outs[1][0])
    n_far = n_all = 0
    for a, b in zip(outs[0][1], outs[1][1]):
        # Adam normalises the gradient: where it is ~0 a last-digit difference of the (atomically summed) gradient can flip
        # the sign of a step of size lr, and such an entry can drift by up to 8 x lr = 4e-3 in these steps.  So: no entry
        # beyond what eight steps can move, and all but a sliver of the 1.19 M entries within a fraction of it (observed:
        # max 1.5e-4 and no entry beyond 3e-4 in most runs; in about one run of eight one hidden unit's relu flips for a
        # sample and the ~150 weights around it end up to 1.2e-3 apart).
        d = (a - b).abs()
        assert float(d.max()) <= 8 * 5e-4 * 1.05
        n_far += int((d > 3e-4).sum())
        n_all += d.numel()
    assert n_far <= 1e-3 * n_all, (n_far, n_all)
    assert outs[0][0][-1] < outs[0][0][0]


def test_graph_captured_step_trains_with_jitter_and_lr_tensor():
    sd, net, r, ro, rd, target = _setup(seed=5, gain=10.0, bias=0.0, n=256)
    net.train()
    r.perturb = 1
    step = T.TrainStep(r, graph=True)
    losses = [float(step(ro.to(DEV), rd.to(DEV), target.to(DEV))) for _ in range(30)]
    assert step._graph is not None
    assert losses[-1] < 0.7 * losses[0], losses[::5]
    assert len(set(losses[5:10])) > 1                       # the jitter differs between replays
    step.set_lr(0.0)                                        # learning rate lives in a device tensor: takes effect in the graph
    before = [p.detach().clone() for p in net.parameters()]
    step(ro.to(DEV), rd.to(DEV), target.to(DEV))
    for a, b in zip(before, net.parameters()):
        assert torch.equal(a, b)


# ------------------------------------------------------------------------------- round 2: the reference's own graph
def _grad_report(net, gref, tag):
    rows = []
    for name, p in net.state_dict(keep_vars=True).items():
        g, ref = p.grad.float().cpu(), gref[name]
        cos = float((g * ref).sum() / (g.norm() * ref.norm() + 1e-30))
        rel = float((g - ref).norm() / (ref.norm() + 1e-30))
        rows.append((name, cos, rel, float(ref.norm())))
        print("%s %-36s |ref| %.3e  cos %.7f  rel %.2e" % (tag, name, float(ref.norm()), cos, rel))
    return rows


def _unmodified_oracle_grads(sd, ro, rd, target, detach_sampler, fine_only=False, dtype=torch.float32):
    """Autograd through the UNMODIFIED restatement of the reference render (O.render_rays is bit-identical to the
    reference's Renderer.render on CPU, tests/test_oracle.py); detach_sampler=False is the reference's graph.
    dtype=float64 evaluates the same graph in double precision: the distance between the two is the reference's OWN
    rounding noise on these inputs."""
    sdg = {k: v.clone().to(dtype).requires_grad_(True) for k, v in sd.items()}
    out = O.render_rays(sdg, ro.to(dtype), rd.to(dtype), detach_sampler=detach_sampler)
    loss = torch.nn.functional.mse_loss(out["rgb_map"], target.to(dtype))
    if not fine_only:
        loss = loss + torch.nn.functional.mse_loss(out["rgb_map_0"], target.to(dtype))
    loss.backward()
    return float(loss.detach()), {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in sdg.items()}


def _check_against_reference_graph(net, sd, ro, rd, target, detach_sampler, fine_only, tag):
    """Every gradient tensor against the reference's fp32 autograd, gated at max(2e-4, 4 x the reference's own fp32
    rounding on that tensor) -- the importance samples of the reference are ill-conditioned in fp32 (t = (u - c0) /
    (c1 - c0) with c1 - c0 down to 1e-5 turns a 1-ulp change of the cdf into a 1e-4 shift of a fine sample, which the
    2^9 octave of the positional encoding turns into a 0.05 rad phase change), so its OWN fp32 gradients sit 1e-3..1e-2
    away from the float64 evaluation of the same graph on the tensors that depend on the fine sample positions."""
    _, g32 = _unmodified_oracle_grads(sd, ro, rd, target, detach_sampler, fine_only)
    _, g64 = _unmodified_oracle_grads(sd, ro, rd, target, detach_sampler, fine_only, dtype=torch.float64)
    worst = 0.0
    for name, p in net.state_dict(keep_vars=True).items():
        g, ref = p.grad.float().cpu().double(), g32[name].double()
        nrm = float(ref.norm())
        rel = float((g - ref).norm() / (nrm + 1e-30))
        floor = float((ref - g64[name]).norm() / (float(g64[name].norm()) + 1e-30))
        print("%s %-36s |ref| %.3e  ours vs ref(fp32) %.2e   ref(fp32) vs ref(fp64) %.2e" % (tag, name, nrm, rel, floor))
        if nrm == 0.0:
            assert float(g.abs().max()) == 0.0, name
            continue
        assert rel <= max(2e-4, 4.0 * floor), (name, rel, floor)
        worst = max(worst, rel / max(floor, 1e-7))
    return g32


@pytest.mark.parametrize("ref_compat_sampler", [True, False])
def test_fp32_training_gradients_vs_unmodified_oracle_autograd(ref_compat_sampler):
    """gradient parity against the reference's autograd on <= 256 rays in the fp32-accurate mode.
    ref_compat_sampler=True is the reference's own graph (sampler NOT detached, volume_renderer.py:181-183): the
    fine loss reaches the coarse network through sample_pdf.  Measured on B200: every tensor that does not depend on
    the fine sample positions (the whole coarse model with the detached sampler, the heads) agrees to 1e-7..1e-6; the
    others sit at the reference's own fp32 rounding (_check_against_reference_graph)."""
    sd, net, r, ro, rd, target = _setup(n=192)
    loss_ref, _ = _unmodified_oracle_grads(sd, ro, rd, target, detach_sampler=not ref_compat_sampler)
    out = T.render_rays_train(r, ro.to(DEV), rd.to(DEV), precision="fp32", ref_compat_sampler=ref_compat_sampler)
    loss = T.nerf_loss(out, target.to(DEV))
    loss.backward()
    assert abs(float(loss.detach()) - loss_ref) <= 1e-5 * max(1.0, abs(loss_ref))
    _check_against_reference_graph(net, sd, ro, rd, target, not ref_compat_sampler, False, "fp32 compat=%d" % ref_compat_sampler)
    if not ref_compat_sampler:      # detached sampler: the coarse model sees identical inputs -> fp32 summation order only
        _, g32 = _unmodified_oracle_grads(sd, ro, rd, target, True)
        for name, p in net.state_dict(keep_vars=True).items():
            if name.startswith("model."):
                rel = float((p.grad.cpu() - g32[name]).norm() / g32[name].norm())
                assert rel <= 5e-6, (name, rel)


def test_fp32_gradients_at_identical_sample_positions():
    """The kernels themselves, with the conditioning of the reference's sampler taken out: the oracle's graph is kept
    (sampler not detached) but the VALUES of its importance samples are replaced by ours (t + (ours - t).detach()), so
    both pipelines evaluate the fine network at the same depths.  Every one of the 48 gradients then agrees to 1e-4 of
    its norm (measured 1e-5..3e-5) -- including the coarse model's, which sees the fine loss only through d z_fine / d weights."""
    from nerf_rep_for_test_b200 import lib as L, ops
    sd, net, r, ro, rd, target = _setup(n=192)
    n = ro.shape[0]
    z_c = ops.sample_coarse(r._table("z"), n)
    raw_c, _ = ops.mlp_forward_train_fp32(T._tensors(r.coarse_model), ro.to(DEV), rd.to(DEV), z_c)
    w_c = ops.composite_forward(raw_c, z_c, rd.to(DEV))[3]
    ours_t = ops.sample_pdf_merge(z_c, w_c, r._table("u"))[1].cpu()                 # our importance samples [n,128]
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    t_vals = O.sample_coarse(n)
    raw = O.query_network(sdg, "model.", ro[:, None, :] + rd[:, None, :] * t_vals[..., None], rd)
    rgb0, _, _, weights, _ = O.raw2outputs(raw, t_vals, rd)
    t_fine, _, _ = O.sample_fine(.5 * (t_vals[..., 1:] + t_vals[..., :-1]), weights[..., 1:-1])
    assert float((t_fine.detach() - ours_t).abs().max()) <= 1e-2      # same samples up to the conditioning of t = (u - c0) / (c1 - c0)
    t_fine = t_fine + (ours_t - t_fine).detach()                                    # same graph, our values
    z_all, _ = torch.sort(torch.cat([t_vals, t_fine], -1), -1)
    raw_f = O.query_network(sdg, "model_fine.", ro[:, None, :] + rd[:, None, :] * z_all[..., None], rd)
    rgb = O.raw2outputs(raw_f, z_all, rd)[0]
    (torch.nn.functional.mse_loss(rgb0, target) + torch.nn.functional.mse_loss(rgb, target)).backward()
    out = T.render_rays_train(r, ro.to(DEV), rd.to(DEV), precision="fp32", ref_compat_sampler=True)
    T.nerf_loss(out, target.to(DEV)).backward()
    for name, p in net.state_dict(keep_vars=True).items():
        ref = sdg[name].grad
        rel = float((p.grad.cpu() - ref).norm() / ref.norm())
        print("same-z %-36s |ref| %.3e rel %.2e" % (name, float(ref.norm()), rel))
        assert rel <= 1e-4, (name, rel)


def test_fp32_fine_only_loss_reaches_the_coarse_network_like_the_reference():
    """A loss on the FINE image alone: with the reference's graph the coarse network still gets a gradient (SURVEY a7:
    sum|g| = 16.3 on the coarse model), with the detached sampler it gets exactly none."""
    sd, net, r, ro, rd, target = _setup(n=128)
    out = T.render_rays_train(r, ro.to(DEV), rd.to(DEV), precision="fp32", ref_compat_sampler=True)
    torch.nn.functional.mse_loss(out["rgb_map"], target.to(DEV)).backward()
    g32 = _check_against_reference_graph(net, sd, ro, rd, target, False, True, "fine-only")
    assert sum(float(v.norm()) for k, v in g32.items() if k.startswith("model.")) > 1e-3     # the path exists in the reference
    net.zero_grad()
    out = T.render_rays_train(r, ro.to(DEV), rd.to(DEV), precision="fp32", ref_compat_sampler=False)
    torch.nn.functional.mse_loss(out["rgb_map"], target.to(DEV)).backward()
    for name, p in net.state_dict(keep_vars=True).items():
        if name.startswith("model."):
            assert float(p.grad.abs().max()) == 0.0, name


def test_composite_backward_z_vs_float64_autograd():
    from nerf_rep_for_test_b200 import ops
    g = torch.Generator().manual_seed(5)
    n, S = 77, 192
    raw = torch.randn(n, S, 4, generator=g) * torch.tensor([1.0, 1.0, 1.0, 3.0])
    z, _ = torch.sort(torch.rand(n, S, generator=g) * 4 + 2, -1)
    rd = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1) * 1.3
    g_rgb, g_acc, g_depth = torch.randn(n, 3, generator=g), torch.randn(n, generator=g), torch.randn(n, generator=g)
    z64 = z.double().requires_grad_(True)
    raw64 = raw.double().requires_grad_(True)
    dists = torch.cat([z64[:, 1:] - z64[:, :-1], torch.full((n, 1), 1e10, dtype=torch.float64)], -1) * rd.double().norm(dim=-1, keepdim=True)
    alpha = 1 - torch.exp(-torch.relu(raw64[..., 3]) * dists)
    T_ = torch.cumprod(torch.cat([torch.ones(n, 1, dtype=torch.float64), 1 - alpha + 1e-10], -1), -1)[:, :-1]
    w = alpha * T_
    rgb_map = (w[..., None] * torch.sigmoid(raw64[..., :3])).sum(-2) + (1 - w.sum(-1, keepdim=True))
    obj = (rgb_map * g_rgb.double()).sum() + (w.sum(-1) * g_acc.double()).sum() + ((w * z64).sum(-1) * g_depth.double()).sum()
    obj.backward()
    g_raw, g_z = ops.composite_backward_z(raw.to(DEV), z.to(DEV), rd.to(DEV), g_rgb.to(DEV), g_acc.to(DEV), g_depth.to(DEV))
    ez = (g_z.cpu().double() - z64.grad).abs().max() / z64.grad.abs().max()
    er = (g_raw.cpu().double() - raw64.grad).abs().max() / raw64.grad.abs().max()
    print("composite_backward_z: max err / max |g|: g_z %.2e, g_raw %.2e" % (float(ez), float(er)))
    assert float(ez) <= 2e-5 and float(er) <= 2e-5
    # the plain entry is the same kernel without the z output
    g_raw2 = ops.composite_backward(raw.to(DEV), z.to(DEV), rd.to(DEV), g_rgb.to(DEV), g_acc.to(DEV), g_depth.to(DEV))
    assert torch.equal(g_raw, g_raw2)


@pytest.mark.parametrize("per_ray_u", [False, True])
def test_sample_pdf_backward_vs_autograd(per_ray_u):
    """d z_all / d weights through torch autograd of the oracle's sample_fine + sort (volume_renderer.py:239-268,
    :181-183) in float64 against nerfb200_sample_pdf_backward; peaked and flat pdfs, table and per-ray u."""
    from nerf_rep_for_test_b200 import ops
    g = torch.Generator().manual_seed(9)
    n, S, U = 150, 64, 128
    z_c = O.sample_coarse(n)
    w = torch.rand(n, S, generator=g) ** 6
    w[: n // 3] = w[: n // 3] * 1e-4                      # nearly flat pdf (the +1e-5 floor dominates)
    w[n // 3: 2 * n // 3, 20:24] += 3.0                   # peaked
    u = torch.rand(n, U, generator=g) if per_ray_u else O.fine_u_table(U)
    g_z_all = torch.randn(n, S + U, generator=g)
    w64 = w.double().requires_grad_(True)
    t_mid = 0.5 * (z_c[:, 1:] + z_c[:, :-1]).double()
    u64 = (u if per_ray_u else u.expand(n, U)).double()
    # evaluate the graph AT THE KERNEL'S OWN VALUES: its fp32 cdf is injected as the value of the float64 cdf (same graph,
    # cdf + (cdf_kernel - cdf).detach()), its bin indices are used, so the `denom < 1e-5` guard (:264), t and the merge
    # order are decided on identical numbers -- t = (u - c0) / (c1 - c0) is too ill-conditioned in flat bins to compare
    # derivatives taken at cdfs that differ in the last fp32 bit
    z_all_k, zs_k, inds_k, cdf_k = ops.sample_pdf_merge(z_c.to(DEV), w.to(DEV), u.to(DEV))
    cdf64 = O.pdf_to_cdf(w64[:, 1:-1])
    assert float((cdf64.detach() - cdf_k.cpu().double()).abs().max()) <= 3e-7
    cdf64 = cdf64 + (cdf_k.cpu().double() - cdf64).detach()
    inds = inds_k.cpu().long()
    below, above = (inds - 1).clamp(min=0), inds.clamp(max=S - 2)
    c0, c1 = torch.gather(cdf64, 1, below), torch.gather(cdf64, 1, above)
    b0, b1 = torch.gather(t_mid, 1, below), torch.gather(t_mid, 1, above)
    denom = c1 - c0
    denom = torch.where(denom.float() < 1e-5, torch.ones_like(denom), denom)
    samples = b0 + (u64 - c0) / denom * (b1 - b0)
    z_all, order = torch.sort(torch.cat([z_c.double(), samples], -1), -1)
    (z_all * g_z_all.double()).sum().backward()
    assert float((z_all.float() - z_all_k.cpu()).abs().max()) <= 1e-5      # same forward
    g_w = ops.sample_pdf_backward(z_c.to(DEV), w.to(DEV), u.to(DEV), g_z_all.to(DEV)).cpu().double()
    ref = w64.grad
    assert float(g_w[:, 0].abs().max()) == 0.0 and float(g_w[:, -1].abs().max()) == 0.0
    err = (g_w - ref).abs().max(-1)[0] / ref.abs().max(-1)[0].clamp_min(1e-12)
    print("sample_pdf_backward per_ray_u=%d: per-ray max err / max |g|: median %.2e max %.2e" % (per_ray_u, float(err.median()), float(err.max())))
    # rays on which a sorted position differs between the float64 graph and the fp32 kernel (ties within 1e-7 between
    # a coarse depth and a sample) permute two entries of g_z_all: allowed on < 2 % of the rays
    assert float(err.median()) <= 1e-4
    assert float((err > 1e-3).float().mean()) <= 0.02


def test_bf16_training_with_the_reference_graph():
    """precision='bf16', ref_compat_sampler=True: the tensor-core training path with the reference's non-detached sampler
    (nerfb200_mlp_backward_input supplies dL/dz through the MLP input).  Checked against the fp32 path -- itself pinned
    to the reference's autograd above -- AT IDENTICAL SAMPLE POSITIONS (the fp32 run is handed the bf16 run's coarse
    weights and merged depths): dL/dz of a random-init field with a 2^9 positional-encoding octave decorrelates under
    the 1e-3 displacement that bf16 coarse weights cause, which says nothing about the backward.  The fine model and
    the coarse heads then agree as bf16 operands allow (cosine > 0.99); the coarse trunk, whose gradient arrives almost
    entirely through sample_pdf, agrees to cosine > 0.85 (see the comment at the assertion)."""
    sd, net, r, ro, rd, target = _setup(n=192)
    rod, rdd, tgt = ro.to(DEV), rd.to(DEV), target.to(DEV)
    names = ["model." + k for k in T._NAMES] + ["model_fine." + k for k in T._NAMES]

    def run(prec, compat, inject=None):
        st = T._forward_passes(r, rod, rdd, prec, _inject=inject)
        rgb0, rgb = st["maps"][0], st["maps"][4]
        scale = 2.0 / rgb0.numel()
        gc, gf = T._backward_passes(r, st, rod, rdd, ((rgb0 - tgt) * scale, None, None), ((rgb - tgt) * scale, None, None), compat)
        return st, dict(zip(names, [g.clone() for g in list(gc) + list(gf)]))

    st16, g16 = run("bf16", True)
    _, g16_detached = run("bf16", False)
    _, g32 = run("fp32", True, inject={"w_c": st16["w_c"], "z_all": st16["z_all"]})
    rows = []
    for name in names:
        a, b, c = g16[name], g32[name], g16_detached[name]
        cos = float((a * b).sum() / (a.norm() * b.norm() + 1e-30))
        rel = float((a - b).norm() / (b.norm() + 1e-30))
        via = float((a - c).norm() / (a.norm() + 1e-30))
        rows.append((name, cos, rel, via))
        print("bf16 compat %-36s cos %.5f rel %.3f   share arriving through the sampler %.3f" % (name, cos, rel, via))
    for name, cos, rel, via in rows:
        if name.startswith("model_fine.") or via == 0.0:
            assert cos > 0.99 and rel < 0.15, (name, cos, rel)       # direct gradients: bf16 operand accuracy
        elif name.endswith("weight") and "pts_linears" in name:
            # through the sampler: dL/dz of the fine pass is bf16-accurate (cos 0.997, scripts/debug_gz.py), dL/d(coarse
            # weights) still 0.992, but the coarse compositor's backward then takes differences of suffix sums of a
            # g_weights 300x larger than the direct dL/dw, which amplifies the 12 % to 20-50 % (the reference's own
            # fp32 rounding already shows 1-2 % on these tensors, test_fp32_training_gradients_...[True])
            assert cos > 0.85, (name, cos, rel)
    # the sampler path carries most of the coarse trunk's gradient in the reference graph
    assert dict((x[0], x[3]) for x in rows)["model.pts_linears.7.weight"] > 0.1
    step = T.TrainStep(r, precision="bf16", ref_compat_sampler=True)
    net.train()
    r.perturb = 1
    losses = [float(step(rod, rdd, tgt)) for _ in range(30)]
    # the reference graph trains more slowly than the detached one on this fixture (0.170 -> 0.160 in 30 steps against
    # 0.170 -> 0.10): the coarse trunk's update direction is dominated by the ill-conditioned sampler path
    assert losses[-1] < 0.98 * losses[0], losses[::5]
    assert all(torch.isfinite(p).all() for p in net.parameters())


def test_fused_loss_and_adam_kernels_match_torch():
    """nerfb200_mse_pair_grad against torch's mse + autograd, nerfb200_adam_clip_step against clip_grad_value_ +
    torch.optim.Adam(fused=True) over five steps (same formulas: parameters agree to fp32 rounding), and a TrainStep with
    the fused update against one with torch's optimizer."""
    from nerf_rep_for_test_b200 import ops
    g = torch.Generator().manual_seed(0)
    n = 777
    rgb0, rgb, t = (torch.rand(n, 3, generator=g).to(DEV) for _ in range(3))
    a, b = rgb0.clone().requires_grad_(True), rgb.clone().requires_grad_(True)
    ref = torch.nn.functional.mse_loss(a, t) + torch.nn.functional.mse_loss(b, t)
    ref.backward()
    loss, g0, g1 = ops.mse_pair_grad(rgb0, rgb, t)
    assert abs(float(loss) - float(ref)) <= 1e-6 * float(ref)
    assert torch.allclose(g0, a.grad, rtol=1e-6, atol=1e-9) and torch.allclose(g1, b.grad, rtol=1e-6, atol=1e-9)
    # against torch.optim.Adam as the reference constructs it (optimizer.py:22-25: the default, non-fused implementation):
    # moments bit-identical / within one fp32 rounding, parameters within 1.2e-7 (measured).  torch's own FUSED CUDA
    # kernel -- what TrainStep used before -- rounds differently (v 1.3e-5 away at a clipped gradient): looser bound.
    m = 100003
    for torch_fused, tol_m, tol_v, tol_p in ((False, 1e-7, 1e-6, 5e-7), (True, 1e-5, 1e-4, 2e-6)):
        g = torch.Generator().manual_seed(1)
        p_ref = torch.nn.Parameter(torch.randn(m, generator=g).to(DEV))
        p, mom, var = p_ref.detach().clone(), torch.zeros(m, device=DEV), torch.zeros(m, device=DEV)
        opt = torch.optim.Adam([p_ref], lr=5e-4, eps=1e-8, fused=torch_fused)
        for step in range(1, 6):
            grad = (torch.randn(m, generator=g) * (60.0 if step == 2 else 1.0)).to(DEV)     # step 2 exercises the clip
            p_ref.grad = grad.clone()
            torch.nn.utils.clip_grad_value_([p_ref], 40.0)
            opt.step()
            gbuf = grad.clone() * 2.0
            ops.adam_clip_step(p, gbuf, mom, var, 5e-4, 0.9, 0.999, 1e-8, step, clip_value=40.0, grad_scale=0.5)
            assert torch.equal(gbuf, p_ref.grad)                                            # scaled + clamped gradient written back
        st = opt.state[p_ref]
        dm, dv = float((mom - st["exp_avg"]).abs().max()), float((var - st["exp_avg_sq"]).abs().max())
        dp = float((p - p_ref.detach()).abs().max())
        print("adam_clip_step vs torch Adam(fused=%s): max |d exp_avg| %.1e |d exp_avg_sq| %.1e |d p| %.1e" % (torch_fused, dm, dv, dp))
        assert dm <= tol_m * 40 and dv <= tol_v * 2 and dp <= tol_p
    # whole step: fused update vs torch's optimizer from the same initial state
    outs = []
    for fused in (True, False):
        sd, net, r, ro, rd, target = _setup(seed=5, gain=10.0, bias=0.0, n=256)
        net.eval()
        r.perturb = 0
        step = T.TrainStep(r, fused_update=fused)
        losses = [float(step(ro.to(DEV), rd.to(DEV), target.to(DEV))) for _ in range(6)]
        outs.append((losses, [q.detach().clone() for q in net.parameters()], step))
    for la, lb in zip(outs[0][0], outs[1][0]):
        assert abs(la - lb) <= 1e-4 * max(abs(lb), 1e-3), (outs[0][0], outs[1][0])
    # two Adam implementations: where a gradient is ~0 (|g| ~ eps) the normalised step is chaotic and a parameter can
    # differ by several lr = 5e-4; everywhere else the trajectories coincide
    diff = torch.cat([(qa - qb).abs().flatten() for qa, qb in zip(outs[0][1], outs[1][1])])
    print("fused update vs torch optimizer after 6 steps: mean |dp| %.2e, fraction > 2e-4: %.2e, max %.2e" % (
        float(diff.mean()), float((diff > 2e-4).float().mean()), float(diff.max())))
    assert float(diff.mean()) <= 2e-5 and float((diff > 2e-4).float().mean()) <= 0.01 and float(diff.max()) <= 6 * 5e-4 + 1e-6
    sa, sb = outs[0][2].state_dict(), outs[1][2].state_dict()
    assert float(sa["state"][0]["step"]) == float(sb["state"][0]["step"]) == 6.0 and len(sa["state"]) == 48
