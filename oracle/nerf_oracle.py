"""CPU oracle: a restatement of the reference NeRF volume-rendering hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs may import this module; the product
path (nerf_rep_for_test_b200/) never does and fails loudly without its CUDA
library.

Every function restates one piece of /root/reference (YuhhhZhao/NeRF-rep_for_test)
with the SAME torch CPU ops in the SAME order, so that on one torch build the
restatement is bit-identical to the real reference (asserted by
tests/test_oracle.py::test_oracle_equals_live_reference whenever /root/reference is mounted, and
pinned by the golden vectors under tests/golden/ that
oracle/gen_golden.py produced from the real reference).

Parity status: the reference's own tests hold NO golden vectors for this path
 -- parity is pinned by outputs of the reference
itself run in the build container (tests/golden/*.npz, generator committed).

The arithmetic primitives (nn.Linear, sin/cos, cumsum, searchsorted, sort,
cumprod, exp, sigmoid) live in the third-party dependency `torch` (reference
README.md:16 names torch==2.3.1; this image has torch 2.11.0), not under
/root/reference; the call sites are cited per function below.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
# synthetic inputs (weights, camera) live in fixtures.py so that the product benchmark never needs this module
from fixtures import (CH_DIR, CH_XYZ, D_LAYERS, FAR, L_DIR, L_XYZ, LEGO_CAMERA_ANGLE_X, LEGO_TEST_POSE0, N_IMPORTANCE,  # noqa: E402,F401
                      N_SAMPLES, NEAR, SKIP, W_HID, lego_batch, make_state_dict, state_dict_shapes)


# ---------------------------------------------------------------------------
# a1  ray generation -- volume_renderer.py:115-147
# ---------------------------------------------------------------------------
def get_rays(H, W, pose, intrinsics):
    """pose [4,4] c2w, intrinsics [3,3]; returns rays_o, rays_d [H*W,3] (normalised)."""
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing="ij")
    i = i.t()
    j = j.t()
    dirs = torch.stack([(i - intrinsics[0, 2]) / intrinsics[0, 0],
                        -(j - intrinsics[1, 2]) / intrinsics[1, 1],
                        -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * pose[:3, :3], -1)
    rays_o = pose[:3, 3].expand(rays_d.shape)
    rays_o = rays_o.reshape(-1, 3)
    rays_d = rays_d.reshape(-1, 3)
    rays_d = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)   # :140
    return rays_o.contiguous(), rays_d.contiguous()


# ---------------------------------------------------------------------------
# a2  stratified coarse sampling -- volume_renderer.py:218-237
# ---------------------------------------------------------------------------
def coarse_t_table(n_samples=N_SAMPLES, near=NEAR, far=FAR, lindisp=False):
    t_vals = torch.linspace(0., 1., steps=n_samples)
    if not lindisp:
        return near * (1. - t_vals) + far * t_vals
    return 1. / (1. / near * (1. - t_vals) + 1. / far * t_vals)


def sample_coarse(n_rays, n_samples=N_SAMPLES, near=NEAR, far=FAR, lindisp=False, t_rand=None):
    z_vals = coarse_t_table(n_samples, near, far, lindisp).expand([n_rays, n_samples])
    if t_rand is not None:   # perturb>0 branch, :228-235, with the jitter injected
        mids = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
        upper = torch.cat([mids, z_vals[..., -1:]], -1)
        lower = torch.cat([z_vals[..., :1], mids], -1)
        z_vals = lower + (upper - lower) * t_rand
    return z_vals


# ---------------------------------------------------------------------------
# a3  positional encoding + MLP -- freq.py:3-32, encoding/__init__.py:7-18,
#     network.py:49-74, volume_renderer.py:270-284
# ---------------------------------------------------------------------------
def pos_enc(x, n_freqs):
    freq_bands = 2. ** torch.linspace(0., n_freqs - 1, steps=n_freqs)     # freq.py:19
    out = [x]
    for freq in freq_bands:
        for p_fn in (torch.sin, torch.cos):                                # sin before cos
            out.append(p_fn(x * freq))                                     # freq.py:25
    return torch.cat(out, -1)


def nerf_mlp(sd, prefix, x, return_hidden=False):
    """network.py:49-74 with use_viewdirs=True; x [M, 90] -> [M, 4] = (rgb_raw, sigma_raw)."""
    input_pts, input_views = torch.split(x, [CH_XYZ, CH_DIR], dim=-1)
    h = input_pts
    hidden = []
    for i in range(D_LAYERS):
        h = F.linear(h, sd[prefix + "pts_linears.%d.weight" % i], sd[prefix + "pts_linears.%d.bias" % i])
        h = F.relu(h)
        hidden.append(h)
        if i == SKIP:
            h = torch.cat([input_pts, h], -1)
    alpha = F.linear(h, sd[prefix + "alpha_linear.weight"], sd[prefix + "alpha_linear.bias"])
    feature = F.linear(h, sd[prefix + "feature_linear.weight"], sd[prefix + "feature_linear.bias"])
    h = torch.cat([feature, input_views], -1)
    h = F.relu(F.linear(h, sd[prefix + "views_linears.0.weight"], sd[prefix + "views_linears.0.bias"]))
    rgb = F.linear(h, sd[prefix + "rgb_linear.weight"], sd[prefix + "rgb_linear.bias"])
    out = torch.cat([rgb, alpha], -1)
    if return_hidden:
        return out, hidden
    return out


def query_network(sd, prefix, pts, view_dirs, chunk=4096):
    """volume_renderer.py:270-284."""
    pts_flat = torch.reshape(pts, [-1, pts.shape[-1]])
    embedded = pos_enc(pts_flat, L_XYZ)
    vd = view_dirs[:, None].expand(pts.shape)
    vd_flat = torch.reshape(vd, [-1, vd.shape[-1]])
    embedded = torch.cat([embedded, pos_enc(vd_flat, L_DIR)], -1)
    raw = torch.cat([nerf_mlp(sd, prefix, embedded[i:i + chunk])
                     for i in range(0, embedded.shape[0], chunk)], 0)
    return torch.reshape(raw, list(pts.shape[:-1]) + [raw.shape[-1]])


# ---------------------------------------------------------------------------
# a5 / a6  compositing -- volume_renderer.py:286-357 and :1089-1157
# ---------------------------------------------------------------------------
def _alpha_rgb(raw, z_vals, rays_d):
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.Tensor([1e10]).expand(dists[..., :1].shape)], -1)
    dists = dists * torch.norm(rays_d[..., None, :], dim=-1)
    rgb = torch.sigmoid(raw[..., :3])
    alpha = 1. - torch.exp(-F.relu(raw[..., 3]) * dists)
    return alpha, rgb


def _maps(weights, rgb, z_vals, white_bkgd):
    rgb_map = torch.sum(weights[..., None] * rgb, -2)
    depth_map = torch.sum(weights * z_vals, -1)
    disp_map = 1. / torch.max(1e-10 * torch.ones_like(depth_map), depth_map / torch.sum(weights, -1))
    acc_map = torch.sum(weights, -1)
    if white_bkgd:
        rgb_map = rgb_map + (1. - acc_map[..., None])
    return rgb_map, disp_map, acc_map, depth_map


def raw2outputs(raw, z_vals, rays_d, white_bkgd=True):
    """volume_renderer.py:286-357 (raw_noise_std == 0)."""
    alpha, rgb = _alpha_rgb(raw, z_vals, rays_d)
    weights = alpha * torch.cumprod(
        torch.cat([torch.ones((alpha.shape[0], 1)), 1. - alpha + 1e-10], -1), -1)[:, :-1]
    rgb_map, disp_map, acc_map, depth_map = _maps(weights, rgb, z_vals, white_bkgd)
    return rgb_map, disp_map, acc_map, weights, depth_map


def raw2outputs_ert(raw, z_vals, rays_d, ert_threshold=0.01, white_bkgd=True, ref_compat=True):
    """volume_renderer.py:1089-1157 (the variant lego.yaml selects, enable_ert=True).

    ref_compat=True restates the reference literally, including the argmax quirk
    (:1115-1123): when ANY ray of the call has T<threshold somewhere, rays that
    never drop below it get first_termination=argmax(all False)=0 and lose ALL
    weights.  ref_compat=False is the intended semantics (truncate only the rays
    that terminate); the two agree whenever no ray or every ray terminates.
    The occupancy-grid side effect (:1147-1155) is restated in ess_update().
    """
    alpha, rgb = _alpha_rgb(raw, z_vals, rays_d)
    n_rays, n_samples = raw.shape[:2]
    alpha_shifted = torch.cat([torch.zeros((n_rays, 1)), alpha[:, :-1]], dim=1)
    transmittance = torch.cumprod(1.0 - alpha_shifted, dim=1)
    weights = alpha * transmittance
    low = transmittance < ert_threshold
    if low.any():
        first = low.float().argmax(dim=1)
        if not ref_compat:
            first = torch.where(low.any(dim=1), first, torch.full_like(first, n_samples))
        idx = torch.arange(n_samples).expand(n_rays, -1)
        weights = weights * (~(idx >= first.unsqueeze(1))).float()
    rgb_map, disp_map, acc_map, depth_map = _maps(weights, rgb, z_vals, white_bkgd)
    return rgb_map, disp_map, acc_map, weights, depth_map


# ---------------------------------------------------------------------------
# a4  hierarchical sampling -- volume_renderer.py:239-268 and call site :181-183
# ---------------------------------------------------------------------------
def fine_u_table(n_importance=N_IMPORTANCE):
    return torch.linspace(0., 1., steps=n_importance)


def pdf_to_cdf(weights):
    """:241-244 -- weights is weights[..., 1:-1] of the coarse pass ([N, 62])."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    return torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)


def sample_from_cdf(t_mids, cdf, u):
    """:253-268 given cdf [N,63], bins t_mids [N,63] and u [N,n_imp]; returns samples, inds."""
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.max(torch.zeros_like(inds - 1), inds - 1)
    above = torch.min((cdf.shape[-1] - 1) * torch.ones_like(inds), inds)
    inds_g = torch.stack([below, above], -1)
    matched_shape = [inds_g.shape[0], inds_g.shape[1], cdf.shape[-1]]
    cdf_g = torch.gather(cdf.unsqueeze(1).expand(matched_shape), 2, inds_g)
    bins_g = torch.gather(t_mids.unsqueeze(1).expand(matched_shape), 2, inds_g)
    denom = (cdf_g[..., 1] - cdf_g[..., 0])
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_g[..., 0]) / denom
    samples = bins_g[..., 0] + t * (bins_g[..., 1] - bins_g[..., 0])
    return samples, inds


def sample_fine(t_mids, weights, u=None, n_importance=N_IMPORTANCE):
    cdf = pdf_to_cdf(weights)
    if u is None:
        u = fine_u_table(n_importance).expand(list(cdf.shape[:-1]) + [n_importance])
    samples, inds = sample_from_cdf(t_mids, cdf, u)
    return samples, inds, cdf


# ---------------------------------------------------------------------------
# a8  occupancy grid (ESS) -- volume_renderer.py:830-873, :963-1007
# ---------------------------------------------------------------------------
BBOX_MIN, BBOX_MAX = -2.0, 2.0


def init_occupancy_grid(res=128, random_mask=None):
    """:830-873; the reference ORs an UNSEEDED rand<0.1 mask (:861) -- injected here."""
    g = torch.stack(torch.meshgrid([torch.arange(res)] * 3, indexing="ij"), -1).float()
    g = (g / (res - 1)) * 2.0 - 1.0
    grid = torch.norm(g, dim=-1) <= 1.2
    if random_mask is not None:
        grid = grid | random_mask
    return grid


def grid_coords(pts, res=128):
    """:992-1003 index arithmetic (truncation toward zero, clamp to faces)."""
    bmin = torch.tensor([BBOX_MIN] * 3)
    bmax = torch.tensor([BBOX_MAX] * 3)
    n = torch.clamp((pts - bmin) / (bmax - bmin), 0, 1)
    c = (n * (res - 1)).long()
    return torch.clamp(c, 0, res - 1)


def is_empty_space(grid, pts):
    c = grid_coords(pts, grid.shape[0])
    return ~grid[c[:, 0], c[:, 1], c[:, 2]]


def ess_update(grid, rays_d, z_vals, raw, weights):
    """Side effect of _raw2outputs_with_ert every grid_update_interval-th call
    (:1147-1155 -> :963-985).  NOTE the reference omits the ray origin (uses
    rays_d*z only); restated literally."""
    eff = weights > 1e-4
    if eff.any():
        pts = (rays_d[..., None, :] * z_vals[..., :, None])[eff]
        dens = F.relu(raw[..., 3])[eff]
        c = grid_coords(pts, grid.shape[0])
        occ = dens > 0.01
        if occ.any():
            oc = c[occ]
            grid[oc[:, 0], oc[:, 1], oc[:, 2]] = True
    return grid


def sample_coarse_ess(grid, rays_o, rays_d, n_samples=N_SAMPLES, near=NEAR, far=FAR, ref_compat=False):
    """_sample_coarse_with_ess (:1009-1087), perturb=0.

    ref_compat=False = intended per-ray semantics: a ray whose empty ratio is
    > 0.5 keeps its occupied z's and adds linspace(min_occ, max_occ, 64-n_keep),
    sorted.  ref_compat=True additionally restates the stride-0 aliasing of the
    reference (:1020,1077: z_vals is an expand()ed view, so each write lands in
    the single shared row and later rays read the already-overwritten row).
    """
    n_rays = rays_o.shape[0]
    base = coarse_t_table(n_samples, near, far)
    z_vals = base.expand([n_rays, n_samples])
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
    is_empty = is_empty_space(grid, pts.reshape(-1, 3)).reshape(n_rays, n_samples)
    empty_ratios = is_empty.float().mean(dim=1)
    highly = empty_ratios > 0.5
    if not highly.any():
        return z_vals.contiguous()
    if ref_compat:
        shared = base.clone()
        for i in range(n_rays):
            if highly[i]:
                occ = shared[~is_empty[i]]
                if len(occ) > 0:
                    n_add = max(0, n_samples - len(occ))
                    comb = torch.cat([occ, torch.linspace(occ.min(), occ.max(), n_add)]) if n_add > 0 else occ
                    shared, _ = torch.sort(comb)
        return shared.expand([n_rays, n_samples]).contiguous()
    out = z_vals.clone()
    for i in range(n_rays):
        if highly[i]:
            occ = base[~is_empty[i]]
            if len(occ) > 0:
                n_add = max(0, n_samples - len(occ))
                comb = torch.cat([occ, torch.linspace(occ.min(), occ.max(), n_add)]) if n_add > 0 else occ
                out[i], _ = torch.sort(comb)
    return out


# ---------------------------------------------------------------------------
# whole path -- volume_renderer.py:109-216 (per 2048-ray chunk :147-205)
# ---------------------------------------------------------------------------
def render_rays(sd, rays_o, rays_d, white_bkgd=True, use_ert=False, ert_threshold=0.01,
                ref_compat=True, ray_chunk=2048, mlp_chunk=4096, return_aux=False,
                n_samples=N_SAMPLES, n_importance=N_IMPORTANCE, near=NEAR, far=FAR, detach_sampler=False):
    """perturb=0, net.eval() (deterministic u), ESS off.  Returns dict of [N,...] maps."""
    outs = {}
    aux = {}

    def put(d, k, v):
        d.setdefault(k, []).append(v)

    comp = (lambda raw, z, d: raw2outputs_ert(raw, z, d, ert_threshold, white_bkgd, ref_compat)) \
        if use_ert else (lambda raw, z, d: raw2outputs(raw, z, d, white_bkgd))
    for s in range(0, rays_o.shape[0], ray_chunk):
        ro, rd = rays_o[s:s + ray_chunk], rays_d[s:s + ray_chunk]
        t_vals = sample_coarse(ro.shape[0], n_samples, near, far)
        pts = ro[..., None, :] + rd[..., None, :] * t_vals[..., :, None]
        raw = query_network(sd, "model.", pts, rd, mlp_chunk)
        rgb0, disp0, acc0, weights, depth0 = comp(raw, t_vals, rd)
        put(outs, "rgb_map_0", rgb0), put(outs, "disp_map_0", disp0)
        put(outs, "acc_map_0", acc0), put(outs, "depth_map_0", depth0)
        t_mid = .5 * (t_vals[..., 1:] + t_vals[..., :-1])
        t_fine, inds, cdf = sample_fine(t_mid, weights[..., 1:-1], None, n_importance)
        if detach_sampler:      # original-NeRF semantics; the reference does NOT detach (:181-183)
            t_fine = t_fine.detach()
        z_all, _ = torch.sort(torch.cat([t_vals, t_fine], -1), -1)
        pts_f = ro[..., None, :] + rd[..., None, :] * z_all[..., :, None]
        raw_f = query_network(sd, "model_fine.", pts_f, rd, mlp_chunk)
        rgb, disp, acc, w_f, depth = comp(raw_f, z_all, rd)
        put(outs, "rgb_map", rgb), put(outs, "disp_map", disp)
        put(outs, "acc_map", acc), put(outs, "depth_map", depth)
        if return_aux:
            for k, v in (("z_coarse", t_vals), ("raw_coarse", raw), ("weights_coarse", weights),
                         ("cdf", cdf), ("inds", inds), ("z_fine_samples", t_fine), ("z_all", z_all),
                         ("raw_fine", raw_f), ("weights_fine", w_f)):
                put(aux, k, v)
    outs = {k: torch.cat(v, 0) for k, v in outs.items()}
    if return_aux:
        return outs, {k: torch.cat(v, 0) for k, v in aux.items()}
    return outs


def render(sd, batch, **kw):
    """Renderer.render(batch) restated: dict of [H,W,(3)] maps (:207-214)."""
    H, W = int(batch["H"]), int(batch["W"])
    pose = batch["pose"].squeeze(0)
    K = batch["intrinsics"].squeeze(0)
    rays_o, rays_d = get_rays(H, W, pose, K)
    res = render_rays(sd, rays_o, rays_d, **kw)
    if isinstance(res, tuple):
        outs, aux = res
    else:
        outs, aux = res, None
    for k in outs:
        outs[k] = outs[k].view(H, W, 3) if k in ("rgb_map", "rgb_map_0") else outs[k].view(H, W)
    return (outs, aux) if aux is not None else outs
