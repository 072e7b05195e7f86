"""GPU: the tcgen05 MLP backward (a7) -- activation store of the training forward, the dgrad chain and the
split-K wgrad GEMMs -- against a torch restatement of the same arithmetic (bf16 operands, wide accumulate)
evaluated on the activations the forward saved, and against fp32 autograd through the CPU oracle."""
import pytest
import torch

from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from nerf_rep_for_test_b200 import Network, RenderConfig, Renderer, ops
    from nerf_rep_for_test_b200 import training as T
    DEV = torch.device("cuda:0")


def _renderer(seed=3, gain=30.0, bias=0.2):
    sd = O.make_state_dict(seed, gain, bias)
    net = Network(device=DEV)
    net.load_state_dict(sd)
    net.to(DEV).eval()
    return sd, Renderer(net, RenderConfig(perturb=0, enable_ess=False, enable_ert=False), mode="bf16")


def _rays(n, S, seed=0):
    g = torch.Generator().manual_seed(seed)
    b = O.lego_batch(40, 40)
    ro, rd = O.get_rays(40, 40, b["pose"][0], b["intrinsics"][0])
    sel = torch.randperm(ro.shape[0], generator=g)[:n]
    z = torch.sort(2.0 + 4.0 * torch.rand(n, S, generator=g), -1)[0]
    return ro[sel].contiguous().to(DEV), rd[sel].contiguous().to(DEV), z.to(DEV)


def _mask_bits(store, plane, n_rows, width):
    """decode the sign-bit planes (train_layout.cuh): True where the gradient is blocked."""
    n_tiles = (n_rows + 127) // 128
    words = store.masks.view(torch.int32).view(9, n_tiles * 128, 8)[plane, :n_rows]
    cols = torch.arange(width, device=DEV)
    g, c = cols // 32, cols % 32
    bit = 8 * (c % 4) + 7 - (c // 4)
    return ((words[:, g] >> bit[None, :]) & 1).bool()


@pytest.mark.parametrize("n,S", [(2, 64), (5, 64), (37, 192)])
def test_train_store_matches_stage_outputs(n, S):
    """tile images written by bulk stores == the stage outputs the diagnostic forward dumps (rows 0..127),
    PE planes == torch PE, relu sign bits consistent with the saved activations; ragged last tile."""
    sd, r = _renderer()
    ro, rd, z = _rays(n, S)
    pk = r.packed("coarse", "bf16")
    raw, store = ops.mlp_forward_train(pk, ro, rd, z)
    raw2, dump = ops.mlp_forward_stages(pk, ro, rd, z)          # unfused ten-stage diagnostic image
    raw3 = ops.mlp_forward(pk, ro, rd, z)                       # fused inference image
    torch.cuda.synchronize()
    M = n * S
    assert torch.equal(raw, raw3)                               # the training forward IS the inference forward
    rows = min(M, 128)
    for i in range(8):
        got = store.plane("h%d" % i)
        assert got.shape == (M, 256)
        assert torch.equal(got[:rows].float(), dump[i, :rows].bfloat16().float()), i
    # the tail is fused (W' = Wv Wf rounded once): relu(views) agrees with the two-step image to bf16 rounding
    hv, hv_ref = store.plane("hv")[:rows].float(), dump[9, :rows, :128]
    assert float((hv - hv_ref).abs().max()) <= 2e-2 * float(hv_ref.abs().max()) + 1e-3
    pts = (ro[:, None, :] + rd[:, None, :] * z[..., None]).reshape(-1, 3)
    pe = O.pos_enc(pts.cpu(), 10).to(DEV)
    dpe = O.pos_enc(rd.cpu(), 4).to(DEV)[:, None, :].expand(n, S, 27).reshape(-1, 27)
    # |x| reaches 6 and the top octave multiplies the fp32 argument error by 512: compare at bf16 + that slack
    assert (store.plane("pe")[:, :63].float() - pe).abs().max() < 4e-2
    assert (store.plane("pe")[:, 63].float()).abs().max() == 0
    assert (store.plane("dpe")[:, :27].float() - dpe).abs().max() < 1e-2
    assert (store.plane("dpe")[:, 27:32].float()).abs().max() == 0
    for i in range(9):
        act = store.plane("h%d" % i if i < 8 else "hv")
        blocked = _mask_bits(store, i, M, act.shape[1])
        assert not bool((blocked & (act > 0)).any()), i          # a positive activation never has its bit set
        frac = float((~blocked & (act == 0)).float().mean())     # +0 / underflow only
        assert frac < 1e-3, (i, frac)


def _torch_backward(sd, prefix, store, g_raw):
    """The arithmetic of mlp_bwd_dgrad.cu / mlp_bwd_wgrad.cu restated with torch (float64 accumulation) on
    the saved planes: bf16 weights, every dL/d(pre-activation) rounded to bf16 before it is reused."""
    f64 = torch.float64
    bf = lambda t: t.to(torch.float32).to(torch.bfloat16).to(f64)
    W = {k[len(prefix):]: v.to(DEV) for k, v in sd.items() if k.startswith(prefix)}
    P = {k: store.plane(k).to(f64) for k in store.PLANES}
    g = g_raw.reshape(-1, 4).to(f64)
    on = lambda a: (a > 0).to(f64)
    d_hv = bf(on(P["hv"]) * (g[:, :3].float() @ W["rgb_linear.weight"]).to(f64))
    Wva, Wf, bfeat = W["views_linears.0.weight"][:, :256], W["feature_linear.weight"], W["feature_linear.bias"]
    Wp = bf(Wva @ Wf)                                            # fused tail weight, rounded once (pack_bf16_bwd_kernel)
    d_pre = [None] * 8
    d_pre[7] = bf(on(P["h7"]) * (d_hv @ Wp + g[:, 3:4] * W["alpha_linear.weight"].to(f64)))
    for i in range(7, 0, -1):
        w = W["pts_linears.%d.weight" % i]
        w = w[:, 63:] if i == 5 else w
        d_pre[i - 1] = bf(on(P["h%d" % (i - 1)]) * (d_pre[i] @ bf(w)))
    G = bf(g)
    grads = {}
    for i in range(8):
        if i == 0:
            gw = d_pre[0].t() @ P["pe"][:, :63]
        elif i == 5:
            gw = torch.cat([d_pre[5].t() @ P["pe"][:, :63], d_pre[5].t() @ P["h4"]], 1)
        else:
            gw = d_pre[i].t() @ P["h%d" % (i - 1)]
        grads["pts_linears.%d.weight" % i] = gw
        grads["pts_linears.%d.bias" % i] = d_pre[i].sum(0)
    dWp, dbp = d_hv.t() @ P["h7"], d_hv.sum(0)                   # gradient of the fused map, then the chain rule
    grads["views_linears.0.weight"] = torch.cat([dWp @ Wf.to(f64).t() + dbp[:, None] * bfeat.to(f64)[None, :],
                                                 d_hv.t() @ P["dpe"][:, :27]], 1)
    grads["views_linears.0.bias"] = dbp
    grads["feature_linear.weight"] = Wva.to(f64).t() @ dWp
    grads["feature_linear.bias"] = Wva.to(f64).t() @ dbp
    grads["alpha_linear.weight"] = G[:, 3:4].t() @ P["h7"]
    grads["alpha_linear.bias"] = G[:, 3].sum(0, keepdim=True)
    grads["rgb_linear.weight"] = G[:, :3].t() @ P["hv"]
    grads["rgb_linear.bias"] = G[:, :3].sum(0)
    inter = {"d9": torch.cat([d_hv, G], 1)}
    inter.update({"dpre%d" % i: d_pre[i] for i in range(8)})
    return grads, inter


@pytest.mark.parametrize("n,S,which", [(2, 64, "coarse"), (5, 64, "coarse"), (37, 192, "fine"), (600, 64, "coarse")])
def test_mlp_backward_kernels_vs_torch(n, S, which):
    sd, r = _renderer()
    ro, rd, z = _rays(n, S, seed=1)
    raw, store = ops.mlp_forward_train(r.packed(which, "bf16"), ro, rd, z)
    M = n * S
    g_raw = torch.randn(M, 4, generator=torch.Generator().manual_seed(7)).to(DEV) * 0.1
    keep = {}
    grads = ops.mlp_backward(r.packed_bwd(which), g_raw, store, keep_workspace=keep)
    torch.cuda.synchronize()
    ref, inter = _torch_backward(sd, "model." if which == "coarse" else "model_fine.", store, g_raw)
    # dgrad planes first (they localise a failure)
    n_tiles = (M + 127) // 128
    dacts = keep["ws"][: n_tiles * store.DACT_BLOCKS * 16384]
    for name in ["d9"] + ["dpre%d" % i for i in range(7, -1, -1)]:
        blk, width = store.DACT_PLANES[name]
        got = ops.untile(dacts, M, store.DACT_BLOCKS, blk, width).double()
        want = inter[name]
        if name == "d9":
            assert float(got[:, 132:].abs().max()) == 0.0
            got = got[:, :132]
        scale = float(want.abs().max()) + 1e-30
        err = float((got - want).abs().max()) / scale
        bad = float(((got - want).abs() > 2e-2 * scale).double().mean())
        print("%-6s max err / scale %.3e   frac > 2%%: %.2e" % (name, err, bad))
        # a bf16 rounding tie or a relu sign flip at |x| ~ 0 moves single elements; the bulk must agree
        assert bad < 2e-3, (name, err, bad)
    for name, g in zip(T._NAMES, grads):
        want = ref[name].reshape(g.shape)
        rel = float((g.double() - want).norm() / (want.norm() + 1e-30))
        print("%-26s rel %.3e" % (name, rel))
        assert rel < 1e-2, (name, rel)


def test_mlp_backward_vs_oracle_autograd():
    """fp32 autograd through the oracle MLP on the same points: bf16 kernels agree to a few 1e-2."""
    sd, r = _renderer(seed=4)
    n, S = 64, 64
    ro, rd, z = _rays(n, S, seed=2)
    raw, store = ops.mlp_forward_train(r.packed("coarse", "bf16"), ro, rd, z)
    g_raw = torch.randn(n * S, 4, generator=torch.Generator().manual_seed(3)) * 0.1
    grads = ops.mlp_backward(r.packed_bwd("coarse"), g_raw.to(DEV), store)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith("model.")}
    pts = (ro[:, None, :] + rd[:, None, :] * z[..., None]).cpu()
    raw_ref = O.query_network(sdg, "model.", pts, rd.cpu())
    (raw_ref.reshape(-1, 4) * g_raw).sum().backward()
    for name, g in zip(T._NAMES, grads):
        want = sdg["model." + name].grad
        cos = float((g.cpu() * want).sum() / (g.cpu().norm() * want.norm() + 1e-30))
        rel = float((g.cpu() - want).norm() / (want.norm() + 1e-30))
        print("%-26s cos %.5f rel %.4f" % (name, cos, rel))
        # the trunk gradients sit behind up to eight bf16 roundings of dL/d(pre-activation) and bf16 forward
        # activations: ~1e-1 relative on a 4096-row batch; the heads see one rounding
        lim = (0.99, 0.15) if name.startswith("pts_linears.") else (0.995, 0.1)
        assert cos > lim[0] and rel < lim[1], (name, cos, rel)


def test_mlp_backward_full_batch_is_additive_over_rays():
    """BASELINE configs[2] size (4096 rays x 192 fine rows = 786 432 MLP rows, every CTA pair busy, thousands of
    red.global.add flushes): the weight gradients of the whole batch equal the sum over two half batches run
    separately -- a size-independent property of the split-K reduction (row-to-tile assignment differs)."""
    sd, r = _renderer(seed=5)
    n, S = 4096, 192
    ro, rd, z = _rays(1600, S, seed=4)
    idx = torch.arange(n, device=DEV) % 1600
    ro, rd, z = ro[idx].contiguous(), rd[idx].contiguous(), (z[idx] + 0.001 * torch.rand(n, S, device=DEV)).contiguous()
    g_raw = (torch.randn(n * S, 4, generator=torch.Generator().manual_seed(9)) * 0.05).to(DEV)
    pk, bw = r.packed("fine", "bf16"), r.packed_bwd("fine")

    def run(lo, hi):
        raw, store = ops.mlp_forward_train(pk, ro[lo:hi], rd[lo:hi], z[lo:hi])
        return ops.mlp_backward(bw, g_raw[lo * S:hi * S], store)

    full = run(0, n)
    a, b = run(0, 1999), run(1999, n)          # odd split: the second half starts in the middle of a 128-row tile
    torch.cuda.synchronize()
    for name, g, ga, gb in zip(T._NAMES, full, a, b):
        want = ga + gb
        rel = float((g - want).norm() / (want.norm() + 1e-30))
        assert rel < 2e-4, (name, rel)
        assert bool(torch.isfinite(g).all()), name


@pytest.mark.parametrize("n,S", [(3, 64), (37, 192), (700, 64)])
def test_mlp_backward_input_vs_torch(n, S):
    """nerfb200_mlp_backward_input (reference graph, bf16 path): g_z = d . PE'(x)^T (dpre0 W0 + dpre5 W5[:, :63]) from the
    dgrad planes the backward left in its workspace, against a float64 torch evaluation on the SAME bf16 planes (so the
    only differences are the bf16 rounding of the two weight blocks and fp32 accumulation)."""
    sd, r = _renderer()
    ro, rd, z = _rays(n, S, seed=3)
    raw, store = ops.mlp_forward_train(r.packed("fine", "bf16"), ro, rd, z)
    M = n * S
    g_raw = torch.randn(M, 4, generator=torch.Generator().manual_seed(11)).to(DEV) * 0.1
    keep = {}
    bwd = r.packed_bwd("fine")
    ops.mlp_backward(bwd, g_raw, store, keep_workspace=keep)
    g_z = ops.mlp_backward_input(bwd, keep["ws"], ro, rd, z)
    n_tiles = (M + 127) // 128
    dacts = keep["ws"][: n_tiles * store.DACT_BLOCKS * 16384]
    d0 = ops.untile(dacts, M, store.DACT_BLOCKS, *store.DACT_PLANES["dpre0"]).double().cpu()
    d5 = ops.untile(dacts, M, store.DACT_BLOCKS, *store.DACT_PLANES["dpre5"]).double().cpu()
    W0 = sd["model_fine.pts_linears.0.weight"].bfloat16().double()
    W5 = sd["model_fine.pts_linears.5.weight"][:, :63].bfloat16().double()
    g_pe = d0 @ W0 + d5 @ W5                                                  # [M, 63]
    x = (ro.cpu()[:, None, :] + rd.cpu()[:, None, :] * z.cpu()[..., None]).reshape(M, 3).double()
    gx = g_pe[:, :3].clone()
    for l in range(10):
        f = 2.0 ** l
        gx += f * (g_pe[:, 3 + 6 * l: 6 + 6 * l] * torch.cos(f * x) - g_pe[:, 6 + 6 * l: 9 + 6 * l] * torch.sin(f * x))
    want = (gx * rd.cpu().double()[:, None, :].expand(n, S, 3).reshape(M, 3)).sum(-1).reshape(n, S)
    err = float((g_z.cpu().double() - want).abs().max() / want.abs().max())
    print("mlp_backward_input n=%d S=%d: max err / max |g_z| = %.2e" % (n, S, err))
    assert err <= 2e-4
