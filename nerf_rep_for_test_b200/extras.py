"""Host-side pieces for the "next" rows of SURVEY 8(f): callers and data formats either side of the hot path.

f1  RayBatchTrainer      -- a ray-batched training loop around TrainStep (the reference's loop feeds whole
                            800x800 images and cannot complete an iteration); uses the otherwise
                            unused task_arg.N_rays / precrop_iters / precrop_frac keys of lego.yaml:14,26-27.
f2  psnr / mse_to_psnr   -- the evaluator's image metric on the device (src/evaluators/nerf.py:389-392,465-474
                            computes it on the CPU after a synchronising copy).
f3  save_model / load_model / load_network -- the reference's checkpoint format (src/utils/net_utils.py:288-380):
                            <dir>/{epoch}.pth or latest.pth holding {'net','optim','scheduler','recorder','epoch'},
                            at most five numbered files kept.  Pass the TrainStep itself as `optim`: its state_dict() /
                            load_state_dict() speak the reference optimizer's per-parameter layout (optimizer.py:14-19).
f4  build_occupancy_grid -- a working version of _populate_occupancy_grid_kilonerf_method
                            (volume_renderer.py:875-961: 3x3x3 density probes per cell, occupied if any > 0.01),
                            as batched density queries through the tcgen05 MLP kernel.
None of this is on the hot path; it is plumbing so that a user of the reference finds the callers they need.
"""
import os

import torch

from . import lib as L
from . import ops


# ----------------------------------------------------------------------------------------------- f2
def mse_to_psnr(mse):
    """src/train/trainers/nerf.py:74: psnr = -10 log10(mse)."""
    return -10.0 * torch.log10(mse)


def psnr(rgb_map, target_rgb):
    """PSNR of a rendered map against the ground truth, on the device, no host synchronisation."""
    return mse_to_psnr(torch.mean((rgb_map.reshape(-1, 3) - target_rgb.reshape(-1, 3)[:, :3].to(rgb_map.device)) ** 2))


# ----------------------------------------------------------------------------------------------- f3
def _numbered(model_dir):
    return [int(p.split(".")[0]) for p in os.listdir(model_dir) if p != "latest.pth" and p.split(".")[0].isdigit()]


def save_model(net, optim, scheduler, recorder, model_dir, epoch, last=False):
    """net_utils.py:323-343 (same file names, same dict keys, same retention of five numbered files)."""
    os.makedirs(model_dir, exist_ok=True)
    sd = lambda o: o.state_dict() if o is not None and hasattr(o, "state_dict") else {}
    model = {"net": net.state_dict(), "optim": sd(optim), "scheduler": sd(scheduler), "recorder": sd(recorder),
             "epoch": epoch}
    torch.save(model, os.path.join(model_dir, "latest.pth" if last else "%d.pth" % epoch))
    pths = _numbered(model_dir)
    if len(pths) > 5:
        os.remove(os.path.join(model_dir, "%d.pth" % min(pths)))


def _pick(model_dir, epoch):
    names = os.listdir(model_dir)
    pths = _numbered(model_dir)
    if not pths and "latest.pth" not in names:
        return None
    if epoch == -1:
        return "latest" if "latest.pth" in names else str(max(pths))
    return str(epoch)


def load_model(net, optim, scheduler, recorder, model_dir, resume=True, epoch=-1):
    """net_utils.py:288-320: resume training; returns the epoch to continue from (0 when nothing is found)."""
    if not resume or not os.path.exists(model_dir):
        return 0
    pth = _pick(model_dir, epoch)
    if pth is None:
        return 0
    ckpt = torch.load(os.path.join(model_dir, pth + ".pth"), map_location="cpu")
    net.load_state_dict(ckpt["net"])
    if "optim" not in ckpt:
        return 0
    if optim is not None:
        optim.load_state_dict(ckpt["optim"])
    if scheduler is not None and ckpt.get("scheduler"):
        scheduler.load_state_dict(ckpt["scheduler"])
    if recorder is not None and ckpt.get("recorder") and hasattr(recorder, "load_state_dict"):
        recorder.load_state_dict(ckpt["recorder"])
    return ckpt["epoch"] + 1


def load_network(net, model_dir, resume=True, epoch=-1, strict=True):
    """net_utils.py:346-380: weights only; model_dir may be a directory or a file.  The renderer's packed
    weight cache notices the new parameter versions by itself (renderer.py: _packed_weights)."""
    if not resume or not os.path.exists(model_dir):
        return 0
    if os.path.isdir(model_dir):
        pth = _pick(model_dir, epoch)
        if pth is None:
            return 0
        path = os.path.join(model_dir, pth + ".pth")
    else:
        path = model_dir
    ckpt = torch.load(path, map_location="cpu")
    net.load_state_dict(ckpt["net"], strict=strict)
    return ckpt["epoch"] + 1 if "epoch" in ckpt else 0


# ----------------------------------------------------------------------------------------------- f4
def query_density(renderer, points, which="coarse", batch=1 << 21):
    """relu(sigma_raw) of arbitrary points [n,3] through the bf16 MLP kernel (one row per point: o = p, z = 0)."""
    dev = renderer.device
    pk = renderer.packed(which, "bf16")
    out = torch.empty(points.shape[0], device=dev)
    d = torch.zeros((1, 3), device=dev)
    d[0, 2] = 1.0
    for i in range(0, points.shape[0], batch):
        p = points[i:i + batch].to(dev, torch.float32).contiguous()
        raw = ops.mlp_forward(pk, p, d.expand(p.shape[0], 3).contiguous(), torch.zeros((p.shape[0], 1), device=dev))
        out[i:i + batch] = torch.relu(raw[:, 0, 3])
    return out


def occupancy_from_density(density_fn, res, density_threshold=0.01, bbox_min=(-2.0, -2.0, -2.0), bbox_max=(2.0, 2.0, 2.0),
                           device="cpu"):
    """Occupancy grid bool [R,R,R] in the geometry of the LOOKUP (volume_renderer.py:992-1007 and ess.cu grid_index map
    p to cell clamp(int((p - lo) / (hi - lo) * (R - 1)), 0, R - 1)): cell i < R-1 covers [lo + i h, lo + (i+1) h) with
    h = (hi - lo) / (R - 1), and cell R-1 is only reached by p >= hi (and by the clamp).  A cell is occupied when any of
    its 3x3x3 probes at offsets {0, 1/2, 1} of the cell has density > threshold (the 27-probe rule of :875-961);
    neighbouring cells share their face probes, so the (2R-1)^3 lattice is evaluated once and max-pooled 3x3x3 with
    stride 2.  The boundary cell R-1 copies its neighbour R-2 (its own extent is the face p = hi).
    The reference builds cell i from [lo + i (hi-lo)/R, ...) (:905-915), one cell size off the lookup at the far end of
    the box; with ess_mode='skip' an "empty" lookup forces density 0, so the builder has to agree with the lookup.
    density_fn(points [n,3]) -> [n] density."""
    res = int(res)
    lo = torch.tensor(bbox_min, device=device, dtype=torch.float32)
    hi = torch.tensor(bbox_max, device=device, dtype=torch.float32)
    m = 2 * (res - 1) + 1
    ax = [lo[k] + (hi[k] - lo[k]) / (res - 1) * (torch.arange(m, device=device, dtype=torch.float32) * 0.5) for k in range(3)]
    dens = torch.empty((m, m, m), device=device)
    for ix in range(m):      # one x-slab at a time bounds the temporary to m^2 points
        pts = torch.stack(torch.meshgrid(ax[0][ix:ix + 1], ax[1], ax[2], indexing="ij"), -1).reshape(-1, 3)
        dens[ix] = density_fn(pts).view(m, m)
    pooled = torch.nn.functional.max_pool3d(dens[None, None], kernel_size=3, stride=2)[0, 0] > density_threshold   # [R-1]^3
    grid = torch.zeros((res, res, res), dtype=torch.bool, device=device)
    grid[:res - 1, :res - 1, :res - 1] = pooled
    grid[res - 1, :, :] = grid[res - 2, :, :]
    grid[:, res - 1, :] = grid[:, res - 2, :]
    grid[:, :, res - 1] = grid[:, :, res - 2]
    return grid


def build_occupancy_grid(renderer, density_threshold=0.01, res=None, bbox_min=(-2.0, -2.0, -2.0), bbox_max=(2.0, 2.0, 2.0)):
    """occupancy_from_density with relu(sigma) of the coarse network evaluated by the tensor-core MLP kernel; indexed
    [x,y,z] like _is_empty_space.  Safe to use with ess_mode='skip': a point whose cell the lookup calls empty lies in a
    cell none of whose 27 probes exceeded the threshold."""
    res = int(res or renderer.occupancy_grid_resolution)
    return occupancy_from_density(lambda pts: query_density(renderer, pts), res, density_threshold, bbox_min, bbox_max,
                                  device=renderer.device)


# ----------------------------------------------------------------------------------------------- f1
class RayBatchTrainer:
    """Original-NeRF style loop: every step draws N_rays pixels from one training view, renders them with the
    differentiable path and takes a TrainStep (fwd + bwd + all-reduce + clip + Adam).  images [V,H,W,3 or 4]
    (alpha is composited onto white when white_bkgd, as blender.py:101-117 does), poses [V,4,4], intrinsics [3,3]."""

    def __init__(self, renderer, images, poses, intrinsics, n_rays=1024, precrop_iters=0, precrop_frac=0.5, lr=5e-4,
                 lr_decay_steps=0, lr_decay_gamma=0.1, seed=0, precision="bf16", ref_compat_sampler=False):
        from .training import TrainStep
        self.r = renderer
        dev = renderer.device
        imgs = torch.as_tensor(images, dtype=torch.float32)
        if imgs.shape[-1] == 4:
            a = imgs[..., 3:4]
            imgs = imgs[..., :3] * a + (1.0 - a) if renderer.white_bkgd else imgs[..., :3]
        self.images = imgs.to(dev)
        self.V, self.H, self.W = imgs.shape[:3]
        K = torch.as_tensor(intrinsics, dtype=torch.float32).reshape(3, 3).to(dev)
        self.rays = [ops.raygen(torch.as_tensor(p, dtype=torch.float32).to(dev), K, self.H, self.W) for p in poses]
        self.n_rays, self.precrop_iters, self.precrop_frac = n_rays, precrop_iters, precrop_frac
        self.step_fn = TrainStep(renderer, lr=lr, precision=precision, ref_compat_sampler=ref_compat_sampler)
        self.lr0, self.decay_steps, self.gamma = lr, lr_decay_steps, lr_decay_gamma
        self.gen = torch.Generator(device=dev).manual_seed(seed)
        self.iteration = 0
        renderer.net.train()
        if float(renderer.perturb) <= 0:
            raise L.NerfB200Error("training expects stratified jitter (perturb > 0, lego.yaml:22)")

    def _pixels(self):
        H, W = self.H, self.W
        if self.iteration < self.precrop_iters:      # centre crop during the first iterations (lego.yaml:26-27)
            dh, dw = int(H // 2 * self.precrop_frac), int(W // 2 * self.precrop_frac)
            ys = torch.randint(H // 2 - dh, H // 2 + dh, (self.n_rays,), device=self.r.device, generator=self.gen)
            xs = torch.randint(W // 2 - dw, W // 2 + dw, (self.n_rays,), device=self.r.device, generator=self.gen)
        else:
            ys = torch.randint(0, H, (self.n_rays,), device=self.r.device, generator=self.gen)
            xs = torch.randint(0, W, (self.n_rays,), device=self.r.device, generator=self.gen)
        return ys * W + xs

    def step(self):
        v = int(torch.randint(0, self.V, (1,), generator=self.gen, device=self.r.device))
        pix = self._pixels()
        ro, rd = self.rays[v]
        target = self.images[v].reshape(-1, 3)[pix]
        if self.decay_steps:
            lr = self.lr0 * self.gamma ** (self.iteration / self.decay_steps)
            self.step_fn.set_lr(lr)
        loss = self.step_fn(ro[pix].contiguous(), rd[pix].contiguous(), target.contiguous())
        self.iteration += 1
        return loss

    def fit(self, n_steps):
        return [self.step() for _ in range(n_steps)]

    @torch.no_grad()
    def evaluate(self, view):
        """PSNR of a full training view with the inference path (perturb off for the render)."""
        r = self.r
        was, tr = r.perturb, r.net.training
        r.perturb = 0
        r.net.eval()
        ro, rd = self.rays[view]
        out = r.render_rays(ro, rd)
        r.perturb = was
        r.net.train(tr)
        return psnr(out["rgb_map"], self.images[view])
