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


def test_checkpoint_written_by_the_reference_itself(tmp_path):
    """a checkpoint written by the reference's OWN net_utils.save_model (:323-343) from its own Network,
    make_optimizer (optimizer.py:8-28: one param group per named parameter) and scheduler/recorder loads into our
    Network (extras.load_network / load_model) and -- through merge_adam_state, what TrainStep.load_state_dict does --
    into the flat single-tensor Adam of the training path; and the other way round, a state split from the flat Adam
    loads into the reference's optimizer via the reference's own load_model."""
    from oracle import ref_loader
    if not ref_loader.reference_available():
        pytest.skip("/root/reference not mounted (GPU box)")
    import os as _os
    import sys as _sys
    from nerf_rep_for_test_b200 import Network
    from nerf_rep_for_test_b200 import extras as X
    from nerf_rep_for_test_b200 import training as T
    cfg, ref_net, _ = ref_loader.build_reference(O.make_state_dict(4))
    import importlib.util
    import types
    if "termcolor" not in _sys.modules:                  # not installed here; net_utils only colours a log line with it
        tc = types.ModuleType("termcolor")
        tc.colored = lambda s, *a, **k: s
        _sys.modules["termcolor"] = tc

    def _load(name, rel):                                # by file path: src/train/__init__.py pulls in the dataset stack
        spec = importlib.util.spec_from_file_location(name, _os.path.join(ref_loader.REFERENCE_ROOT, rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    make_optimizer = _load("ref_optimizer", "src/train/optimizer.py").make_optimizer
    net_utils = _load("ref_net_utils", "src/utils/net_utils.py")
    ref_net.train()
    opt = make_optimizer(cfg, ref_net)
    for it in range(2):                                   # two Adam steps so that the state is populated
        for p in ref_net.parameters():
            p.grad = torch.randn(p.shape, generator=torch.Generator().manual_seed(it)) * 1e-3
        opt.step()

    class _SD:                                            # scheduler / recorder stand-ins with the state_dict protocol
        def __init__(self):
            self.d = {"x": 1}

        def state_dict(self):
            return dict(self.d)

        def load_state_dict(self, d):
            self.d = dict(d)
    d = str(tmp_path / "refckpt")
    net_utils.save_model(ref_net, opt, _SD(), _SD(), d, 7, last=True)          # the reference's writer
    ours = Network(device=torch.device("cpu"))
    assert X.load_network(ours, d) == 8
    for k, v in ref_net.state_dict().items():
        assert torch.equal(v, ours.state_dict()[k]), k
    ck = torch.load(_os.path.join(d, "latest.pth"))
    assert len(ck["optim"]["param_groups"]) == 48
    # -> flat Adam (what TrainStep.load_state_dict does), in TrainStep's own storage order
    params = [p for m in (ours.model, ours.model_fine) for p in T.model_params(m)]
    offsets, off = {}, 0
    for p in params:
        offsets[id(p)] = off
        off += p.numel()
    flat_sd = T.merge_adam_state(ck["optim"], list(ours.named_parameters()), offsets, off, torch.zeros(1))
    flat = torch.nn.Parameter(torch.zeros(off))
    flat_opt = torch.optim.Adam([flat], lr=1.0)
    flat_opt.load_state_dict(flat_sd)
    ref_state = opt.state_dict()["state"]
    for i, (name, p) in enumerate(ours.named_parameters()):
        o = offsets[id(p)]
        assert torch.equal(flat_opt.state_dict()["state"][0]["exp_avg"][o:o + p.numel()].view_as(p), ref_state[i]["exp_avg"]), name
    # <- and back: split, save with OUR writer, load with the REFERENCE's load_model into a fresh reference optimizer
    split = T.split_adam_state(flat_opt.state_dict(), list(ours.named_parameters()), offsets)
    d2 = str(tmp_path / "ourckpt")

    class _Opt:
        def state_dict(self):
            return split
    X.save_model(ours, _Opt(), _SD(), _SD(), d2, 3, last=True)
    _, ref_net2, _ = ref_loader.build_reference(O.make_state_dict(5))
    opt2 = make_optimizer(cfg, ref_net2)
    assert net_utils.load_model(ref_net2, opt2, _SD(), _SD(), d2) == 4
    for i in range(48):
        assert torch.equal(opt2.state_dict()["state"][i]["exp_avg_sq"], ref_state[i]["exp_avg_sq"]), i
    for k, v in ref_net.state_dict().items():
        assert torch.equal(v, ref_net2.state_dict()[k]), k


def test_flat_adam_state_speaks_the_reference_optimizer_layout():
    """TrainStep runs Adam over ONE flat tensor; the reference's optimizer (src/train/optimizer.py:14-19) has
    one param group per named parameter.  split_adam_state / merge_adam_state convert both ways: after identical steps
    the split state equals the state of a stock per-parameter Adam, loads into it, and merges back bit for bit."""
    from nerf_rep_for_test_b200 import training as T
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    named = list(net.named_parameters())
    n_total = sum(p.numel() for _, p in named)
    # a different storage order than named_parameters() (TrainStep orders by model_params, not by registration)
    storage = [named[2][1], named[3][1], named[0][1], named[1][1]]
    offsets, off = {}, 0
    flat = torch.nn.Parameter(torch.zeros(n_total))
    for p in storage:
        offsets[id(p)] = off
        flat.data[off:off + p.numel()] = p.data.reshape(-1)
        off += p.numel()
    ref_opt = torch.optim.Adam([{"params": [p], "lr": 5e-4, "weight_decay": 0.0, "eps": 1e-8} for _, p in named], 5e-4, eps=1e-8)
    flat_opt = torch.optim.Adam([flat], lr=5e-4, eps=1e-8)
    for it in range(3):
        g = torch.randn(n_total, generator=torch.Generator().manual_seed(it))
        flat.grad = g.clone()
        for p in storage:
            p.grad = g[offsets[id(p)]:offsets[id(p)] + p.numel()].view_as(p).clone()
        ref_opt.step()
        flat_opt.step()
    split = T.split_adam_state(flat_opt.state_dict(), named, offsets)
    ref_sd = ref_opt.state_dict()
    assert len(split["param_groups"]) == len(ref_sd["param_groups"]) == 4
    assert [g["params"] for g in split["param_groups"]] == [g["params"] for g in ref_sd["param_groups"]]
    for i in range(4):
        for k in ("exp_avg", "exp_avg_sq"):
            assert torch.allclose(split["state"][i][k], ref_sd["state"][i][k], rtol=0, atol=0), (i, k)
        assert float(split["state"][i]["step"]) == float(ref_sd["state"][i]["step"]) == 3.0
    fresh = torch.optim.Adam([{"params": [p]} for _, p in named], 1.0)
    fresh.load_state_dict(split)                                 # what the reference's load_model does (:312)
    assert fresh.param_groups[0]["lr"] == 5e-4
    merged = T.merge_adam_state(ref_sd, named, offsets, n_total, flat.data)
    for k in ("exp_avg", "exp_avg_sq"):
        assert torch.equal(merged["state"][0][k], flat_opt.state_dict()["state"][0][k])
    flat2 = torch.optim.Adam([torch.nn.Parameter(torch.zeros(n_total))], lr=1.0)
    flat2.load_state_dict(merged)
    assert flat2.param_groups[0]["lr"] == 5e-4 and float(flat2.state_dict()["state"][0]["step"]) == 3.0
    assert T.merge_adam_state(flat_opt.state_dict(), named, offsets, n_total, flat.data) is not None   # flat layout passes through
    with pytest.raises(ValueError):
        T.merge_adam_state({"state": {}, "param_groups": ref_sd["param_groups"][:3]}, named, offsets, n_total, flat.data)


def test_occupancy_builder_uses_the_lookup_geometry():
    """the grid must be built in the geometry of the lookup (cell = 4/(R-1), volume_renderer.py:992-1007),
    otherwise ess_mode='skip' cuts geometry on the +x/+y/+z side.  Analytic density: every point with density above the
    threshold must be found occupied by the reference's own lookup, for off-centre blobs and at the box boundary."""
    from nerf_rep_for_test_b200 import extras as X
    g = torch.Generator().manual_seed(0)
    for centre, radius in (((1.3, 1.1, 1.4), 0.45), ((-1.2, 0.3, 1.7), 0.3), ((1.9, 1.9, 1.9), 0.4)):
        c = torch.tensor(centre)
        dens = lambda p: torch.relu(radius - (p - c).norm(dim=-1))
        for res in (16, 33):
            grid = X.occupancy_from_density(dens, res, density_threshold=0.01)
            pts = c + (torch.rand(20000, 3, generator=g) * 2 - 1) * radius
            pts = pts[(dens(pts) > 0.01) & (pts.abs().max(-1)[0] <= 2.0)]
            assert pts.shape[0] > 100
            assert not bool(O.is_empty_space(grid, pts).any()), (centre, res)
            # and it is not trivially full: far from the blob the lookup says empty
            far = -c.sign() * 1.5 + (torch.rand(1000, 3, generator=g) - 0.5) * 0.2
            assert bool(O.is_empty_space(grid, far).all())


gpu = pytest.mark.gpu


def _net(sd, dev):
    from nerf_rep_for_test_b200 import Network
    net = Network(device=dev)
    net.load_state_dict(sd)
    return net.to(dev)


@gpu
def test_occupancy_grid_builder_vs_reference_algorithm():
    """build_occupancy_grid == the 27-probe rule of volume_renderer.py:875-961 evaluated with the fp32 oracle MLP on the
    cells of the LOOKUP geometry (cells whose maximum density sits within the bf16 error of the threshold may differ)."""
    from nerf_rep_for_test_b200 import RenderConfig, Renderer
    from nerf_rep_for_test_b200 import extras as X
    dev = torch.device("cuda:0")
    sd = O.make_state_dict(7, 40.0, -1.2)
    r = Renderer(_net(sd, dev).eval(), RenderConfig(perturb=0, enable_ess=False, enable_ert=False), mode="bf16")
    res = 11
    grid = X.build_occupancy_grid(r, density_threshold=0.01, res=res).cpu()
    lo, cell = -2.0, 4.0 / (res - 1)
    idx = torch.stack(torch.meshgrid([torch.arange(res - 1)] * 3, indexing="ij"), -1).reshape(-1, 3).float()   # (x,y,z)
    off = torch.stack(torch.meshgrid([torch.arange(3)] * 3, indexing="ij"), -1).reshape(-1, 3).float() / 2.0
    pts = lo + (idx[:, None, :] + off[None, :, :]) * cell                                                    # [(R-1)^3,27,3]
    emb = O.pos_enc(pts.reshape(-1, 3), 10)
    raw = O.nerf_mlp(sd, "model.", torch.cat([emb, torch.zeros(emb.shape[0], 27)], -1))
    dens = torch.relu(raw[:, 3]).view(-1, 27).max(1)[0]
    want = (dens > 0.01).view(res - 1, res - 1, res - 1)
    margin = (dens - 0.01).abs().view(res - 1, res - 1, res - 1) < 5e-3
    got = grid[:res - 1, :res - 1, :res - 1]
    assert 0.05 < float(want.float().mean()) < 0.95
    assert bool(((got == want) | margin).all())
    assert float((got != want).float().mean()) < 0.02
    assert torch.equal(grid[res - 1], grid[res - 2]) and torch.equal(grid[:, :, res - 1], grid[:, :, res - 2])


@gpu
def test_skip_render_with_a_built_grid_stays_close_to_the_dense_render():
    """ess_mode='skip' with a grid produced by build_occupancy_grid.  A skipped sample's density is forced to
    0, so the grid must cover every sample whose density matters: the skip render may differ from the dense render only by
    what densities below the build threshold (0.01) contribute -- bounded here on an off-centre blob that reaches the +x
    side of the box, where the old builder (cells of 4/R instead of the lookup's 4/(R-1)) cut geometry."""
    from nerf_rep_for_test_b200 import RenderConfig, Renderer
    from nerf_rep_for_test_b200 import extras as X
    dev = torch.device("cuda:0")
    sd = O.make_state_dict(6, 300.0, 6.0)
    for k in list(sd):
        if k.startswith("model_fine."):
            sd[k] = sd["model." + k[len("model_fine."):]].clone()
    net = _net(sd, dev).eval()
    dense = Renderer(net, RenderConfig(perturb=0, enable_ess=False, enable_ert=False), mode="bf16")
    skip = Renderer(net, RenderConfig(perturb=0, enable_ess=True, enable_ert=False), mode="bf16")
    skip.occupancy_grid = X.build_occupancy_grid(skip, density_threshold=0.01, res=64)
    skip.ess_mode = "skip"
    occ = float(skip.occupancy_grid.float().mean())
    b = O.lego_batch(48, 48)
    gb = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}
    a, c = dense.render(gb), skip.render(gb)
    for k in ("rgb_map_0", "acc_map_0", "rgb_map", "acc_map"):
        err = (a[k] - c[k]).abs()
        print("skip vs dense %-10s max %.2e mean %.2e (grid %.0f %% occupied, %.1f MLP rows per ray)" % (
            k, float(err.max()), float(err.mean()), 100 * occ, float(skip.eval_counts.sum()) / (48 * 48)))
        # coarse pass: same sample positions, only sub-threshold densities removed (64 samples x 0.01 x 0.0635 ~ 4e-2 worst case)
        assert float(err.max()) <= (5e-2 if k.endswith("_0") else 1.5e-1), k
        assert float(err.mean()) <= 1e-2, k


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
    X.save_model(student, tr.step_fn, None, None, d, 0, last=True)          # TrainStep speaks the reference's optimizer layout
    ck = torch.load(os.path.join(d, "latest.pth"))
    assert len(ck["optim"]["param_groups"]) == 48 and len(ck["optim"]["state"]) == 48
    ref_opt = torch.optim.Adam([{"params": [p]} for _, p in student.named_parameters()], 5e-4)
    ref_opt.load_state_dict(ck["optim"])                                     # what the reference's load_model does (:312)
    m0 = tr.step_fn.opt.state_dict()["state"][0]["exp_avg"].clone()
    tr.step_fn.load_state_dict(ck["optim"])
    assert torch.equal(tr.step_fn.opt.state_dict()["state"][0]["exp_avg"], m0)
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
