"""GPU parity of the BF16 (tcgen05) mode.

north_star: rgb/acc within 1e-3 absolute, depth within 1e-3 relative to its range, < 0.05 dB PSNR
against the reference renderer on identical rays and weights.  The kernel itself is additionally
pinned, stage by stage, against a CPU emulation of its exact arithmetic (tests/bf16_emul.py)."""
import pytest
import torch

from conftest import golden
from oracle import nerf_oracle as O
from bf16_emul import mlp_bf16_stages

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from nerf_rep_for_test_b200 import lib as L
    from nerf_rep_for_test_b200 import ops
    DEV = torch.device("cuda:0")


def _rays(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    b = O.lego_batch(64, 64)
    ro, rd = O.get_rays(64, 64, b["pose"][0], b["intrinsics"][0])
    sel = torch.randperm(ro.shape[0], generator=g)[:n]
    return ro[sel].contiguous(), rd[sel].contiguous()


def test_stage_outputs_match_bf16_emulation():
    sd = O.make_state_dict(0)
    ro, rd = _rays(2)
    z = O.sample_coarse(2)
    packed = ops.pack_from_state_dict(sd, "model_fine.", L.MODE_BF16, DEV)
    raw, dump = ops.mlp_forward_stages(packed, ro.to(DEV), rd.to(DEV), z.to(DEV))
    pts = (ro[:, None, :] + rd[:, None, :] * z[..., None]).reshape(-1, 3)
    dirs = rd[:, None, :].expand(2, 64, 3).reshape(-1, 3)
    ref_raw, stages = mlp_bf16_stages(sd, "model_fine.", pts, dirs)
    for i, st in enumerate(stages):
        err = (dump[i, :128, : st.shape[1]].cpu() - st).abs().max()
        # differences come only from bf16 roundings that flip on last-bit input differences
        assert float(err) < 2e-3 * max(1.0, float(st.abs().max())), (i, float(err))
    assert float((raw.cpu().reshape(-1, 4) - ref_raw).abs().max()) < 5e-4


@pytest.mark.parametrize("n,S", [(1, 1), (3, 5), (2, 64), (5, 192), (700, 64), (1300, 192)])
def test_mlp_bf16_vs_emulation_and_oracle(n, S):
    """ragged sizes: partial tiles, odd tile counts, several tile pairs per CTA (1300x192 = 1950 tiles)."""
    sd = O.make_state_dict(1)
    ro, rd = _rays(n, seed=S)
    z, _ = torch.sort(torch.rand(n, S, generator=torch.Generator().manual_seed(n)) * 4 + 2, -1)
    packed = ops.pack_from_state_dict(sd, "model.", L.MODE_BF16, DEV)
    raw = ops.mlp_forward(packed, ro.to(DEV), rd.to(DEV), z.to(DEV)).cpu()
    pts = (ro[:, None, :] + rd[:, None, :] * z[..., None]).reshape(-1, 3)
    dirs = rd[:, None, :].expand(n, S, 3).reshape(-1, 3)
    with torch.no_grad():
        emu, _ = mlp_bf16_stages(sd, "model.", pts, dirs)
        ref = O.nerf_mlp(sd, "model.", torch.cat([O.pos_enc(pts, 10), O.pos_enc(dirs, 4)], -1))
    e_emu = (raw.reshape(-1, 4) - emu).abs()
    e_ref = (raw.reshape(-1, 4) - ref).abs()
    print("n=%d S=%d  vs emulation max %.2e  vs fp32 oracle max %.2e median %.2e" % (
        n, S, float(e_emu.max()), float(e_ref.max()), float(e_ref.median())))
    assert float(e_emu.max()) < 1e-3
    assert float(e_ref.max()) < 5e-3
    # same inputs twice -> bit-identical (no race between the two tile slots)
    raw2 = ops.mlp_forward(packed, ro.to(DEV), rd.to(DEV), z.to(DEV)).cpu()
    assert torch.equal(raw, raw2)


def _renderer(sd, mode="bf16", **cfg):
    from nerf_rep_for_test_b200 import Network, RenderConfig, Renderer
    net = Network(device=DEV)
    net.load_state_dict(sd)
    net.to(DEV).eval()
    base = dict(perturb=0, enable_ess=False, enable_ert=False)
    base.update(cfg)
    return Renderer(net, RenderConfig(**base), mode=mode)


def _psnr(a, b):
    return float(-10.0 * torch.log10(((a - b) ** 2).mean().clamp_min(1e-20)))


# rays of each golden whose LAST reference sigma_raw (coarse, fine pass) lies within the mode's sigma error bound of
# zero -- a property of the frozen reference outputs, so the counts are pinned, not bounded by a blanket allowance
FLIP_CANDIDATES = {("bf16", "lego16_randinit"): (1, 28), ("bf16", "lego8_dense"): (0, 0),
                   ("fp16", "lego16_randinit"): (1, 4), ("fp16", "lego8_dense"): (0, 0)}
SIGMA_BOUND = {"bf16": 2e-3, "fp16": 2.5e-4}


@pytest.mark.parametrize("mode", ["bf16", "fp16"])
@pytest.mark.parametrize("name", ["lego16_randinit", "lego8_dense"])
def test_render_bf16_vs_reference_golden(name, mode):
    g = golden(name)
    H, W, seed, gain, bias, ert = g["meta"]
    sd = O.make_state_dict(int(seed), float(gain), float(bias))
    out = _renderer(sd, mode=mode).render({"pose": torch.from_numpy(g["pose"]).to(DEV),
                                           "intrinsics": torch.from_numpy(g["intrinsics"]).to(DEV), "H": int(H), "W": int(W)})
    # outlier rule (SURVEY 8c' item 3): the last interval is 1e10 wide, so alpha_last = [sigma_last > 0]
    # is a step function of an MLP output; rays whose reference |sigma_raw_last| is below the mode's
    # sigma error bound (bf16 2e-3, fp16 2.5e-4) can flip acc by ~1 and are excluded and counted.
    sig_last_f = torch.from_numpy(g["aux_raw_fine"])[:, -1, 3].abs()
    sig_last_c = torch.from_numpy(g["aux_raw_coarse"])[:, -1, 3].abs()
    thr = SIGMA_BOUND[mode]
    assert (int((sig_last_c <= thr).sum()), int((sig_last_f <= thr).sum())) == FLIP_CANDIDATES[(mode, name)]
    for k in ("rgb_map_0", "acc_map_0", "depth_map_0", "rgb_map", "acc_map", "depth_map"):
        ref = torch.from_numpy(g["out_" + k])
        a = out[k].cpu()
        stable = ((sig_last_c if k.endswith("_0") else sig_last_f) > thr).reshape(ref.shape[:2])
        err = (a - ref).abs()
        if err.dim() == 3:
            err = err.max(-1)[0]
        scale = 6.0 if "depth" in k else 1.0
        print("%s %-11s abs err/scale median %.2e p99 %.2e max(stable) %.2e  excluded %d of %d rays" % (
            name, k, float(err.median()) / scale, float(err.flatten().kthvalue(int(0.99 * err.numel()))[0]) / scale,
            float(err[stable].max()) / scale, int((~stable).sum()), stable.numel()))
        assert float(err[stable].max()) <= 1e-3 * scale, k
    ref_rgb = torch.from_numpy(g["out_rgb_map"])
    stable = (sig_last_f > thr).reshape(ref_rgb.shape[:2])
    # PSNR of our image against the reference's image; "< 0.05 dB" is about PSNR vs ground truth,
    # which needs a trained checkpoint (absent); the image-to-image PSNR is reported instead
    print("%s PSNR(ours, reference) over stable rays: %.1f dB" % (name, _psnr(out["rgb_map"].cpu()[stable], ref_rgb[stable])))
    assert _psnr(out["rgb_map"].cpu()[stable], ref_rgb[stable]) > 60.0


# ------------------------------------------------------------------------------- full-size properties
def test_full_frame_800x800_independence_and_oracle_subset():
    """BASELINE.json configs[1] at full size (640 000 rays): every ray's result is independent of where
    it sits in the batch (32768-ray chunks, 512-row quads, tile slots, CTA pairs) -- a random subset
    rendered on its own is BIT-IDENTICAL to the same pixels of the full frame -- and that subset agrees
    with the CPU oracle within the bf16 tolerances."""
    sd = O.make_state_dict(0, 25.0, 0.1)
    r = _renderer(sd)
    b = O.lego_batch(800, 800)
    full = r.render({k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in b.items()})
    ro, rd = ops.raygen(b["pose"].to(DEV), b["intrinsics"].to(DEV), 800, 800)
    sel = torch.randperm(640000, generator=torch.Generator().manual_seed(7))[:2048].to(DEV)
    sub = r.render_rays(ro[sel], rd[sel])
    for k in ("rgb_map_0", "acc_map_0", "depth_map_0", "rgb_map", "acc_map", "depth_map", "disp_map"):
        a = full[k].reshape(640000, -1)[sel]
        assert torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(sub[k].reshape(2048, -1), nan=-7.0)), k
    # permutation of the rays permutes the outputs, nothing else
    perm = torch.randperm(2048, generator=torch.Generator().manual_seed(8)).to(DEV)
    sub_p = r.render_rays(ro[sel][perm], rd[sel][perm])
    assert torch.equal(sub_p["rgb_map"], sub["rgb_map"][perm])
    with torch.no_grad():
        ref, aux = O.render_rays(sd, ro[sel].cpu(), rd[sel].cpu(), return_aux=True)
    q99 = lambda e: float(e.flatten().kthvalue(int(0.99 * e.numel()))[0])
    # (a) coarse pass: no dependence on sample placement
    for k in ("rgb_map_0", "acc_map_0", "depth_map_0"):
        err = (sub[k].cpu() - ref[k]).abs()
        scale = 6.0 if "depth" in k else 1.0
        print("800x800 subset %-11s abs err/scale p99 %.2e max %.2e" % (k, q99(err) / scale, float(err.max()) / scale))
        assert q99(err) <= 1e-3 * scale, k      # max can hit a last-sample sign flip (SURVEY 8c' item 3)
    # (b) fine pass at the REFERENCE's sample positions: the kernels' own arithmetic error
    packed = r.packed("fine")
    raw_f = ops.mlp_forward(packed, ro[sel], rd[sel], aux["z_all"].to(DEV))
    rgb, _, acc, _, depth = ops.composite_forward(raw_f, aux["z_all"].to(DEV), rd[sel])
    for k, v in (("rgb_map", rgb), ("acc_map", acc), ("depth_map", depth)):
        err = (v.cpu() - ref[k]).abs()
        scale = 6.0 if "depth" in k else 1.0
        print("800x800 subset %-11s (reference z) abs err/scale p99 %.2e max %.2e" % (k, q99(err) / scale, float(err.max()) / scale))
        assert q99(err) <= 1e-3 * scale, k
    # (c) end to end: bf16 coarse weights move the importance samples, and a random-init field with
    # PE up to 2^9 is rough at that scale, so the fine quadrature itself changes; reported, loosely bounded
    for k in ("rgb_map", "acc_map", "depth_map"):
        err = (sub[k].cpu() - ref[k]).abs()
        scale = 6.0 if "depth" in k else 1.0
        print("800x800 subset %-11s (end to end)  abs err/scale median %.2e p99 %.2e" % (k, float(err.median()) / scale, q99(err) / scale))
        assert float(err.median()) <= 2e-3 * scale and q99(err) <= 3e-2 * scale, k


def test_bf16_psnr_against_ground_truth_within_0p05_db():
    """north_star: "< 0.05 dB PSNR in bf16 mode".  No trained checkpoint or ground-truth image exists in the
    reference tree (latest.pth is missing), so the criterion is exercised on a synthetic ground truth: the fp32
    oracle render of a teacher field.  A perturbed copy of the teacher plays the model under evaluation; its PSNR
    against the ground truth must not move by more than 0.05 dB when the reference's renderer (fp32 oracle) is
    swapped for ours (bf16 mode), for the coarse and the fine image."""
    teacher = O.make_state_dict(11, 300.0, 19.4)          # sigma_raw spread to +-2.6 around zero: structured, opaque scene
    gsd = torch.Generator().manual_seed(3)
    # weights perturbed by 60 % of their spread: PSNR 32-36 dB, the range of a trained lego model
    student = {k: v + 0.6 * v.std() * torch.randn(v.shape, generator=gsd) if v.dim() == 2 else v.clone() for k, v in teacher.items()}
    b = O.lego_batch(40, 40)
    with torch.no_grad():
        gt = O.render(teacher, b)
        ref = O.render(student, b)
    ours = _renderer(student).render({k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in b.items()})
    for k in ("rgb_map_0", "rgb_map"):
        p_ref, p_ours = _psnr(ref[k], gt[k]), _psnr(ours[k].cpu(), gt[k])
        print("%-9s PSNR vs ground truth: reference renderer %.3f dB, ours (bf16) %.3f dB, difference %.4f dB" % (k, p_ref, p_ours, p_ours - p_ref))
        assert 15.0 < p_ref < 60.0                          # a meaningful, non-degenerate operating point
        assert abs(p_ours - p_ref) < 0.05, (k, p_ref, p_ours)
