"""GPU: error behaviour of the C ABI entry points added for a7 / a9 -- a bad argument is a non-zero return with
a message (NerfB200Error here), never a crash or a silent no-op; empty inputs are accepted."""
import ctypes as C

import pytest
import torch

import fixtures as FX

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from nerf_rep_for_test_b200 import Network, RenderConfig, Renderer, kilo, lib as L, ops
    DEV = torch.device("cuda:0")


def _renderer():
    net = Network(device=DEV)
    net.load_state_dict(FX.make_state_dict(1))
    return Renderer(net.to(DEV).eval(), RenderConfig(perturb=0, enable_ess=False, enable_ert=False), mode="bf16")


def test_mlp_backward_argument_checks():
    r = _renderer()
    ro = torch.zeros(3, 3, device=DEV); rd = torch.ones(3, 3, device=DEV) / 3 ** 0.5
    z = torch.linspace(2, 6, 64, device=DEV).expand(3, 64).contiguous()
    raw, store = ops.mlp_forward_train(r.packed("coarse", "bf16"), ro, rd, z)
    with pytest.raises(L.NerfB200Error):                      # g_raw of the wrong batch
        ops.mlp_backward(r.packed_bwd("coarse"), torch.zeros(5, 4, device=DEV), store)
    with pytest.raises(L.NerfB200Error):                      # gradient buffers of the wrong shape
        ops.mlp_backward(r.packed_bwd("coarse"), torch.zeros(192, 4, device=DEV), store,
                         grads=[torch.zeros(7, device=DEV)] * 24)
    with pytest.raises(L.NerfB200Error):                      # CPU tensors are rejected: there is no CPU path
        ops.mlp_backward(r.packed_bwd("coarse"), torch.zeros(192, 4), store)
    lib = L.load()
    packed_bwd, w = r.packed_bwd("coarse")
    g = L.MlpGrads()
    # workspace too small / null gradient pointers -> error code + message, nothing launched
    rc = lib.nerfb200_mlp_backward(packed_bwd, C.byref(w), L.dev(torch.zeros(192, 4, device=DEV)), L.dev(store.acts), L.dev(store.masks),
                                   192, L.dev(torch.zeros(256, dtype=torch.uint8, device=DEV)), 256, C.byref(g), L.stream_ptr())
    assert rc != 0 and b"null gradient" in lib.nerfb200_get_last_error_string()
    grads = ops.mlp_backward(r.packed_bwd("coarse"), torch.zeros(192, 4, device=DEV), store)      # zero upstream gradient
    torch.cuda.synchronize()
    assert all(float(t.abs().max()) == 0.0 for t in grads)


def test_mlp_forward_train_rejects_fp32_mode_and_accepts_empty():
    r = _renderer()
    lib = L.load()
    z = torch.zeros(0, 64, device=DEV)
    raw, store = ops.mlp_forward_train(r.packed("coarse", "bf16"), torch.zeros(0, 3, device=DEV), torch.zeros(0, 3, device=DEV), z)
    assert raw.shape == (0, 64, 4) and store.acts.numel() == 0
    pk = r.packed("coarse", "fp32")
    rc = lib.nerfb200_mlp_forward_train(pk.ptr, L.MODE_FP32, None, None, None, 1, 64, None, None, None, L.stream_ptr())
    assert rc != 0


def test_kilo_argument_checks():
    sc = FX.make_kilo_scene(seed=0, net_res=4, grid_res=16)
    with pytest.raises(L.NerfB200Error):                      # parameter block of the wrong size
        kilo.KiloRenderer(sc["grid"], sc["params"][:, :100], sc["domain_mins"], sc["domain_maxs"], sc["gmin"], sc["gmax"], 0.01, 384, 2.0, device=DEV)
    r = kilo.KiloRenderer(sc["grid"], sc["params"], sc["domain_mins"], sc["domain_maxs"], sc["gmin"], sc["gmax"], 0.01, 384, 2.0, device=DEV)
    b = FX.lego_batch(8, 8)
    out = r.render({k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in b.items()})
    torch.cuda.synchronize()
    assert out["rgb_map"].shape == (8, 8, 3) and bool(torch.isfinite(out["rgb_map"]).all())
    r.dbp = -1.0                                              # non-positive step
    with pytest.raises(L.NerfB200Error):
        r.render({k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in b.items()})
    r.dbp, r.max_depth = 0.01, 1 << 22                        # H*W*max_depth overflows the int32 query index at 800x800
    b800 = FX.lego_batch(800, 800)
    with pytest.raises(L.NerfB200Error):
        r.render({k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in b800.items()})
    # grid on the wrong device / dtype
    with pytest.raises(L.NerfB200Error):
        kilo.generate_query_indices_on_ray([0, 0, 4], torch.zeros(4, 3, device=DEV), torch.zeros(4, 4, 4, dtype=torch.int32, device=DEV),
                                           torch.ones(4, dtype=torch.uint8, device=DEV), torch.zeros(4, dtype=torch.int32, device=DEV),
                                           [-1, -1, -1], [1, 1, 1], 0.01, 4, 16, 2.0, True)


def test_round2_entry_points_empty_ragged_and_bad_arguments():
    """The entry points added in round 2 (fp32 training twin, dL/dz, sample_pdf backward, literal ESS, jitter): empty
    inputs are accepted and leave defined outputs, ragged sizes run, bad arguments are a non-zero return + message."""
    from nerf_rep_for_test_b200 import training as T
    r = _renderer()
    lib = L.load()
    tensors = T._tensors(r.coarse_model)
    # ---- empty
    e3, e64 = torch.zeros(0, 3, device=DEV), torch.zeros(0, 64, device=DEV)
    raw, acts = ops.mlp_forward_train_fp32(tensors, e3, e3, e64)
    assert raw.shape == (0, 64, 4) and acts.numel() == 0
    grads, g_z = ops.mlp_backward_fp32(tensors, torch.zeros(0, 4, device=DEV), acts, e3, 0, 64, want_g_z=True)
    torch.cuda.synchronize()
    assert all(float(g.abs().max()) == 0.0 for g in grads) and g_z.shape == (0, 64)   # gradients are OVERWRITTEN: zeros
    g_raw, gz = ops.composite_backward_z(torch.zeros(0, 64, 4, device=DEV), e64, e3, torch.zeros(0, 3, device=DEV))
    assert g_raw.shape == (0, 64, 4) and gz.shape == (0, 64)
    assert ops.sample_pdf_backward(e64, e64, torch.linspace(0, 1, 128, device=DEV), torch.zeros(0, 192, device=DEV)).shape == (0, 64)
    grid = torch.ones(8, 8, 8, dtype=torch.uint8, device=DEV)
    assert ops.ess_resample_compat(grid, e3, e3, torch.linspace(2, 6, 64, device=DEV)).shape == (0, 64)
    assert ops.jitter_rows(e64).shape == (0, 64)
    # ---- ragged: one ray, three samples (the smallest sample_pdf accepts), odd counts
    for n, S, U in ((1, 3, 1), (5, 17, 33), (130, 64, 128)):
        ro = torch.randn(n, 3, device=DEV) * 0.1
        rd = torch.nn.functional.normalize(torch.randn(n, 3, device=DEV), dim=-1)
        z = torch.sort(torch.rand(n, S, device=DEV) * 4 + 2, -1)[0]
        raw, acts = ops.mlp_forward_train_fp32(tensors, ro, rd, z)
        grads, g_z = ops.mlp_backward_fp32(tensors, torch.randn(n * S, 4, device=DEV), acts, rd, n, S, want_g_z=True)
        w = ops.composite_forward(raw, z, rd)[3]
        u = torch.linspace(0, 1, U, device=DEV)
        z_all = ops.sample_pdf_merge(z, w, u, want_aux=False)[0]
        g_w = ops.sample_pdf_backward(z, w, u, torch.randn(n, S + U, device=DEV))
        torch.cuda.synchronize()
        assert bool(torch.isfinite(g_z).all()) and bool(torch.isfinite(g_w).all()) and all(bool(torch.isfinite(g).all()) for g in grads)
        assert bool((z_all[:, 1:] >= z_all[:, :-1]).all())
    # ---- bad arguments
    with pytest.raises(L.NerfB200Error):                      # g_z_all of the wrong width
        ops.sample_pdf_backward(torch.zeros(4, 64, device=DEV), torch.zeros(4, 64, device=DEV), torch.linspace(0, 1, 128, device=DEV),
                                torch.zeros(4, 100, device=DEV))
    with pytest.raises(L.NerfB200Error):                      # literal ESS needs n_samples <= 64 (one 64-bit mask per ray)
        ops.ess_resample_compat(grid, torch.zeros(2, 3, device=DEV), torch.ones(2, 3, device=DEV), torch.linspace(2, 6, 65, device=DEV))
    with pytest.raises(L.NerfB200Error):                      # fp32 backward: g_raw of the wrong batch
        ops.mlp_backward_fp32(tensors, torch.zeros(7, 4, device=DEV), acts, rd, n, S)
    w_struct, keep = ops.weights_struct(tensors)
    rc = lib.nerfb200_mlp_backward_fp32(C.byref(w_struct), L.dev(torch.zeros(64, 4, device=DEV)), L.dev(torch.zeros(1 << 20, dtype=torch.uint8, device=DEV)),
                                        None, 1, 64, L.dev(torch.zeros(16, dtype=torch.uint8, device=DEV)), 16, C.byref(L.MlpGrads()), None, L.stream_ptr())
    assert rc != 0 and b"null gradient" in lib.nerfb200_get_last_error_string()
    with pytest.raises(L.NerfB200Error):                      # reference graph in bf16 needs the workspace of the backward
        ops.mlp_backward_input(r.packed_bwd("fine"), None, torch.zeros(2, 3, device=DEV), torch.ones(2, 3, device=DEV),
                               torch.zeros(2, 64, device=DEV))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_renderer_on_a_non_current_device():
    """the C ABI launches on the CURRENT device.  A network on cuda:1 rendered while cuda:0 is current must
    run on cuda:1 (the Python entry points switch devices around every call; kernel attributes are per device) and give
    the same image as the same network on cuda:0."""
    from oracle import nerf_oracle as O
    from nerf_rep_for_test_b200 import Network, RenderConfig, Renderer, ops
    sd = O.make_state_dict(0, 30.0, 0.2)
    outs = []
    for idx in (0, 1):
        dev = torch.device("cuda", idx)
        net = Network(device=dev)
        net.load_state_dict(sd)
        net.to(dev).eval()
        torch.cuda.set_device(0)                                  # cuda:0 stays current throughout
        for mode in ("bf16", "fp32tc", "fp32"):
            r = Renderer(net, RenderConfig(perturb=0, enable_ess=False, enable_ert=False), mode=mode)
            b = O.lego_batch(12, 12)
            out = r.render({k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()})
            assert out["rgb_map"].device == dev
            outs.append(out["rgb_map"].cpu())
        z = ops.sample_coarse(torch.linspace(2, 6, 64, device=dev), 7)
        assert z.device == dev
    for a, b in zip(outs[:3], outs[3:]):
        assert torch.equal(a, b)
