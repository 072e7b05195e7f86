"""SURVEY 8(f) "next" rows: reference-format checkpoints (CPU), and on the GPU the ray-batched trainer,
device PSNR and the density-query occupancy-grid builder."""
import os

import pytest
import torch

from oracle import nerf_oracle as O


def test_checkpoint_format_roundtrip(tmp_path):
    """net_utils.py:323-380: file names, dict keys, retention of five numbered files, latest.pth preference."""
    from nerf_rep_for_test_b200 import Network
    from nerf_rep_for_test_b200 import extras as X
    net = Network(device=torch.device("cpu"))
    net.load_state_dict(O.make_state_dict(1))
    opt = torch.optim.Adam(net.parameters(), lr=5e-4, eps=1e-8)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.99)
    d = str(tmp_path / "ckpt")
    for e in range(7):
        X.save_model(net, opt, sched, None, d, e)
    assert sorted(os.listdir(d)) == ["%d.pth" % e for e in range(2, 7)]          # at most five numbered files
    X.save_model(net, opt, sched, None, d, 9, last=True)
    ck = torch.load(os.path.join(d, "latest.pth"))
    assert set(ck) == {"net", "optim", "scheduler", "recorder", "epoch"} and ck["epoch"] == 9
    assert list(ck["net"]) == list(O.make_state_dict(1))                          # the 48 reference state_dict keys
    net2 = Network(device=torch.device("cpu"))
    assert X.load_network(net2, d) == 10                                          # latest.pth wins, epoch + 1
    for k, v in net.state_dict().items():
        assert torch.equal(v, net2.state_dict()[k])
    opt2 = torch.optim.Adam(net2.parameters(), lr=1.0)
    sched2 = torch.optim.lr_scheduler.ExponentialLR(opt2, gamma=0.5)
    assert X.load_model(net2, opt2, sched2, None, d, epoch=6) == 7
    assert opt2.param_groups[0]["lr"] == opt.param_groups[0]["lr"]
    assert X.load_network(net2, str(tmp_path / "missing")) == 0
    # a checkpoint written by the reference's own save_model is a plain dict with the same keys: load a hand-made one
    torch.save({"net": O.make_state_dict(2)}, os.path.join(d, "plain.pth"))
    assert X.load_network(net2, os.path.join(d, "plain.pth")) == 0
    assert torch.equal(net2.state_dict()["model.alpha_linear.weight"], O.make_state_dict(2)["model.alpha_linear.weight"])


gpu = pytest.mark.gpu


def _net(sd, dev):
    from nerf_rep_for_test_b200 import Network
    net = Network(device=dev)
    net.load_state_dict(sd)
    return net.to(dev)


@gpu
def test_occupancy_grid_builder_vs_reference_algorithm():
    """build_occupancy_grid == the 27-probe rule of volume_renderer.py:875-961 evaluated with the fp32 oracle MLP
    (cells whose maximum density sits within the bf16 error of the threshold may differ)."""
    from nerf_rep_for_test_b200 import RenderConfig, Renderer
    from nerf_rep_for_test_b200 import extras as X
    dev = torch.device("cuda:0")
    sd = O.make_state_dict(7, 40.0, -1.2)
    r = Renderer(_net(sd, dev).eval(), RenderConfig(perturb=0, enable_ess=False, enable_ert=False), mode="bf16")
    res = 10
    grid = X.build_occupancy_grid(r, density_threshold=0.01, res=res).cpu()
    lo, cell = -2.0, 4.0 / res
    idx = torch.stack(torch.meshgrid([torch.arange(res)] * 3, indexing="ij"), -1).reshape(-1, 3).float()       # (x,y,z)
    off = torch.stack(torch.meshgrid([torch.arange(3)] * 3, indexing="ij"), -1).reshape(-1, 3).float() / 2.0
    pts = lo + (idx[:, None, :] + off[None, :, :]) * cell                                                    # [R^3,27,3]
    emb = O.pos_enc(pts.reshape(-1, 3), 10)
    raw = O.nerf_mlp(sd, "model.", torch.cat([emb, torch.zeros(emb.shape[0], 27)], -1))
    dens = torch.relu(raw[:, 3]).view(-1, 27).max(1)[0]
    want = (dens > 0.01).view(res, res, res)
    margin = (dens - 0.01).abs().view(res, res, res) < 5e-3
    assert 0.05 < float(want.float().mean()) < 0.95
    assert bool(((grid == want) | margin).all())
    assert float((grid != want).float().mean()) < 0.02


@gpu
def test_ray_batch_trainer_improves_psnr_and_checkpoints(tmp_path):
    """f1 + f2 + f3 together: fit a student to views rendered from a teacher, PSNR (device) goes up, and the
    checkpoint restores the renderer's output exactly (packed-weight cache follows the new parameters)."""
    from nerf_rep_for_test_b200 import RenderConfig, Renderer
    from nerf_rep_for_test_b200 import extras as X
    dev = torch.device("cuda:0")
    Himg = 40
    # random-init sigma_raw is -0.065 +- 0.009 for seed 11: gain 300 / bias 19.4 spreads it to +-2.6 around zero
    teacher = Renderer(_net(O.make_state_dict(11, 300.0, 19.4), dev).eval(),
                       RenderConfig(perturb=0, enable_ess=False, enable_ert=False), mode="bf16")
    b = O.lego_batch(Himg, Himg)
    K = b["intrinsics"][0]
    import math
    poses = []
    for i in range(4):
        th = 2 * math.pi * i / 4
        rot = torch.tensor([[math.cos(th), -math.sin(th), 0, 0], [math.sin(th), math.cos(th), 0, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]])
        poses.append(rot @ b["pose"][0])
    images = torch.stack([teacher.render({"pose": p[None].to(dev), "intrinsics": K[None].to(dev), "H": Himg, "W": Himg})["rgb_map"]
                          for p in poses])
    student = _net(O.make_state_dict(12, 1.0, 0.15), dev)       # small positive density everywhere: relu alive
    r = Renderer(student, RenderConfig(perturb=1, enable_ess=False, enable_ert=False), mode="bf16")
    tr = X.RayBatchTrainer(r, images, poses, K, n_rays=512, precrop_iters=20, precrop_frac=0.5, seed=1)
    p0 = float(tr.evaluate(0))
    losses = [float(l) for l in tr.fit(150)]
    p1 = float(tr.evaluate(0))
    print("PSNR view 0: %.2f dB -> %.2f dB; loss %.4f -> %.4f" % (p0, p1, sum(losses[:10]) / 10, sum(losses[-10:]) / 10))
    assert p1 > p0 + 2.0
    assert sum(losses[-10:]) < 0.6 * sum(losses[:10])
    d = str(tmp_path / "ck")
    X.save_model(student, tr.step_fn.opt, None, None, d, 0, last=True)
    student.eval()
    r.perturb = 0
    ro, rd = tr.rays[1]
    mid = Himg * (Himg // 2)                       # rows through the image centre (the corners only see background)
    ro, rd = ro[mid - 128:mid + 128].contiguous(), rd[mid - 128:mid + 128].contiguous()
    before = r.render_rays(ro, rd)["rgb_map"].clone()
    with torch.no_grad():
        for p in student.parameters():
            p.add_(0.05)
        student.model_fine.alpha_linear.bias.add_(5.0)      # make sure the perturbed field is not empty
    assert float((r.render_rays(ro, rd)["rgb_map"] - before).abs().max()) > 1e-3
    assert X.load_network(student, d) == 1
    assert torch.equal(r.render_rays(ro, rd)["rgb_map"], before)
