"""CPU oracle of the KiloNeRF-style path (BASELINE configs[4] part ii).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs, never by
the product path (nerf_rep_for_test_b200/kilo.py calls libnerfb200.so and nothing else).

Restates, in numpy float32 with the operation order of the CUDA sources, the kernels of the reference's
`cuda/` extension that define this path:
  get_rays_d        cuda/generate_inputs.cu:11-35
  march             cuda/generate_inputs.cu:60-126   (generate_query_indices_on_ray_kernel_0)
  network_eval      cuda/network_eval.cu:24-254      (network_eval_query_index_kernel_0<32>)
  integrate         cuda/integrate.cu:9-57
  background        cuda/integrate.cu:84-97

PARITY UNPINNED by the reference: the extension is never built or loaded there (it needs MAGMA + GL,
cuda/setup.py:7,18-19), its Python driver is a stub (volume_renderer.py:1159-1413), it ships no trained
micro-MLP weights, no network-assignment grid and no test vectors.  The anchor is the kernel sources
themselves; `oracle/build_kilo_ref.py` additionally compiles those few .cu files where they lie into
oracle/_ref/ so the GPU tests can run the reference's own kernels next to ours on the same inputs.
Known, tolerated difference: the reference evaluates cos/sin with the fast __cosf/__sinf intrinsics on
arguments up to 512 rad (network_eval.cu:136-139); this oracle uses exact float32 cos/sin.
"""
import os as _os
import sys as _sys

import numpy as np

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))

F = np.float32
HIDDEN = 32
POS_F, DIR_F = 10, 4
POS_EMB, DIR_EMB = 3 * (2 * POS_F + 1), 3 * (2 * DIR_F + 1)          # 63, 27
PARAM_SIZE = (POS_EMB + 1) * HIDDEN + (HIDDEN + 1) * HIDDEN + (HIDDEN + 1) * (HIDDEN + 1) + \
    (HIDDEN + DIR_EMB + 1) * HIDDEN + (HIDDEN + 1) * 3                # 6212 (network_eval.cu:48-52)


def get_rays_d(H, W, cx, cy, fx, fy, c2w):
    """generate_inputs.cu:11-35: rays_d[y,x,:] = c2w(3x3) . [(x-cx)/fx, -(y-cy)/fy, -1], NOT normalised."""
    c2w = np.asarray(c2w, F).reshape(3, 3)
    x = np.arange(W, dtype=F)[None, :]
    y = np.arange(H, dtype=F)[:, None]
    v = [np.broadcast_to((x - F(cx)) / F(fx), (H, W)), np.broadcast_to(-(y - F(cy)) / F(fy), (H, W)),
         np.full((H, W), -1, F)]
    out = np.zeros((H, W, 3), F)
    for j in range(3):          # accumulation order of network_eval.cu:87-92 (j outer)
        for i in range(3):
            out[..., i] = out[..., i] + v[j] * c2w[i, j]
    return out.reshape(-1, 3)


def march(origin, directions, grid, active, depth_idx, gmin, gmax, dbp, max_samples, max_depth, min_distance,
          initial):
    """generate_inputs.cu:60-126.  grid: int16 [R0,R1,R2] (network id or -1).  Returns query_indices
    [n,max_samples] int32 (unfilled = -1 here; the reference leaves them uninitialised), assigned
    [n,max_samples] int16 (-1 = unfilled) and the updated (active, depth_idx)."""
    origin, gmin, gmax = np.asarray(origin, F), np.asarray(gmin, F), np.asarray(gmax, F)
    directions = np.asarray(directions, F)
    n = directions.shape[0]
    res = np.array(grid.shape, np.int64)
    voxel = ((gmax - gmin) / res.astype(F)).astype(F)
    strides = np.array([res[1] * res[2], res[2], 1], np.int64)
    flat = grid.reshape(-1)
    q = np.full((n, max_samples), -1, np.int32)
    a = np.full((n, max_samples), -1, np.int16)
    active = np.ones(n, bool) if initial else active.copy()
    depth_idx = np.zeros(n, np.int32) if initial else depth_idx.astype(np.int32).copy()
    out_pos = np.zeros(n, np.int64)
    depth = depth_idx.copy()
    dist = (F(min_distance) + depth.astype(F) * F(dbp)).astype(F)           # :78
    running = active.copy()
    eps = F(0.001)
    while True:
        live = running & (depth < max_depth) & (out_pos < max_samples)       # :79
        if not live.any():
            break
        idx = np.nonzero(live)[0]
        p = (origin[None, :] + dist[idx, None] * directions[idx]).astype(F)  # :83
        vi = ((p - gmin[None, :]) / voxel[None, :]).astype(F).astype(np.int64)   # :84 float -> int truncation
        inside = np.all((gmin[None, :] + eps < p) & (p < gmax[None, :] - eps), axis=1)   # :87
        fl = (vi * strides[None, :]).sum(1)
        net = np.where(inside, flat[np.clip(fl, 0, flat.size - 1)], -1)      # :90
        hit = net != -1
        hi = idx[hit]
        a[hi, out_pos[hi]] = net[hit]
        q[hi, out_pos[hi]] = (hi * max_depth + depth[hi]).astype(np.int32)   # :95
        out_pos[hi] += 1
        depth[idx] += 1
        dist[idx] = (dist[idx] + F(dbp)).astype(F)                           # :108 accumulated, not recomputed
    full = out_pos >= max_samples
    new_active = active.copy()
    new_active[active & ~full] = False                                       # :111-113
    new_depth = depth_idx.copy()
    new_depth[active & full] = depth[active & full]                          # :114-117
    return q, a, new_active, new_depth


def fourier(x, n_freq):
    """per-scalar grouping, cos BEFORE sin (network_eval.cu:127-140): [x, cos(f0 x..f9 x), sin(f0 x..f9 x)]."""
    f = (2.0 ** np.arange(n_freq)).astype(F)
    arg = (x[..., None] * f).astype(F)
    return np.concatenate([x[..., None], np.cos(arg).astype(F), np.sin(arg).astype(F)], -1)   # [..., 2n+1]


def network_eval(query_indices, networks, params, domain_mins, domain_maxs, H, W, cx, cy, fx, fy, c2w, origin,
                 max_depth, min_distance, dbp):
    """network_eval.cu:24-254 for a flat list of queries with their network ids -> [n,4] (sigmoid rgb,
    relu sigma).  params: [num_networks, 6212] packed per layer as [bias(out) | W(in-major, out fastest)]."""
    qi = np.asarray(query_indices, np.int64)
    nets = np.asarray(networks, np.int64)
    c2w = np.asarray(c2w, F).reshape(3, 3)
    origin = np.asarray(origin, F)
    depth = qi % max_depth
    pix = qi // max_depth
    x, y = (pix % W).astype(F), (pix // W).astype(F)
    v = [(x - F(cx)) / F(fx), -(y - F(cy)) / F(fy), np.full(x.shape, -1, F)]
    d = np.zeros(qi.shape + (3,), F)
    for j in range(3):
        for i in range(3):
            d[:, i] = d[:, i] + v[j] * c2w[i, j]
    dist = (F(min_distance) + depth.astype(F) * F(dbp)).astype(F)            # :96
    pos = (origin[None, :] + dist[:, None] * d).astype(F)
    norm = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]).astype(F)).astype(F)
    d = (d / norm[:, None]).astype(F)
    dmin, dmax = np.asarray(domain_mins, F)[nets], np.asarray(domain_maxs, F)[nets]
    local = (F(2) * (pos - dmin) / (dmax - dmin) - F(1)).astype(F)           # :123
    P = np.asarray(params, F)[nets]                                          # [n, 6212]
    o = 0

    def take(nin, nout):
        nonlocal o
        b = P[:, o:o + nout]
        w = P[:, o + nout:o + nout + nin * nout].reshape(-1, nin, nout)
        o += nout + nin * nout
        return b, w

    relu = lambda t: np.maximum(t, F(0))
    emb = fourier(local, POS_F).reshape(-1, POS_EMB)                         # j-major, e-minor
    b, w = take(POS_EMB, HIDDEN)
    h0 = b + np.einsum("nk,nko->no", emb, w)
    b, w = take(HIDDEN, HIDDEN)
    h1 = b + np.einsum("nk,nko->no", relu(h0), w)
    b, w = take(HIDDEN, HIDDEN + 1)
    h2 = b + np.einsum("nk,nko->no", relu(h1), w)                            # [:,0] = density, rest = feature (no relu)
    demb = fourier(d, DIR_F).reshape(-1, DIR_EMB)
    b, w = take(HIDDEN + DIR_EMB, HIDDEN)
    h3 = b + np.einsum("nk,nko->no", np.concatenate([h2[:, 1:], demb], 1), w)
    b, w = take(HIDDEN, 3)
    rgb = b + np.einsum("nk,nko->no", relu(h3), w)
    assert o == PARAM_SIZE
    out = np.concatenate([1.0 / (1.0 + np.exp(-rgb.astype(np.float64))), relu(h2[:, :1])], 1)
    return out.astype(F)


def integrate(rgb_sigma, n_filled, dists, rgb_map, acc_map, T, active, thr, initial):
    """integrate.cu:9-57 on the first n_filled[ray] slots of each ray (the reference truncates the unfilled
    tail beforehand).  In-place on copies; returns (rgb_map, acc_map, T, active)."""
    n, spp, _ = rgb_sigma.shape
    rgb_map, acc_map, T, active = rgb_map.copy(), acc_map.copy(), T.copy(), active.copy()
    if initial:
        T[:] = 1
    proc = T > F(thr)
    if initial:
        rgb_map[:] = 0
        acc_map[:] = 0
    for s in range(spp):
        m = proc & (s < n_filled)
        if not m.any():
            continue
        rs = rgb_sigma[m, s]
        alpha = (F(1) - np.exp(-(rs[:, 3] * dists[m]).astype(F)).astype(F)).astype(F)
        w = (alpha * T[m]).astype(F)
        T[m] = (T[m] * (F(1) - alpha + F(1e-10)).astype(F)).astype(F)
        rgb_map[m] = (rgb_map[m] + rs[:, :3] * w[:, None]).astype(F)
        acc_map[m] = (acc_map[m] + w).astype(F)
    active[proc & (T <= F(thr))] = False
    return rgb_map, acc_map, T, active


def render(H, W, cx, cy, fx, fy, c2w, origin, grid, params, domain_mins, domain_maxs, gmin, gmax, dbp, max_depth,
           min_distance, spp, thr, white_bkgd=True, max_passes=64):
    """multi-pass driver in the spirit of KiloNeRF's renderer (no driver exists in the reference repo)."""
    d = get_rays_d(H, W, cx, cy, fx, fy, c2w)
    n = d.shape[0]
    dists = (F(dbp) * np.sqrt((d * d).sum(1).astype(F))).astype(F)
    rgb_map, acc_map, T = np.zeros((n, 3), F), np.zeros(n, F), np.ones(n, F)
    active, depth_idx = np.ones(n, bool), np.zeros(n, np.int32)
    evaluated = 0
    for p in range(max_passes):
        q, a, active, depth_idx = march(origin, d, grid, active, depth_idx, gmin, gmax, dbp, spp, max_depth,
                                        min_distance, p == 0)
        filled = a >= 0
        n_filled = filled.sum(1)
        evaluated += int(filled.sum())
        rs = np.zeros((n, spp, 4), F)
        if filled.any():
            rs[filled] = network_eval(q[filled], a[filled], params, domain_mins, domain_maxs, H, W, cx, cy, fx, fy,
                                      c2w, origin, max_depth, min_distance, dbp)
        rgb_map, acc_map, T, active = integrate(rs, n_filled, dists, rgb_map, acc_map, T, active, thr, p == 0)
        if not active.any():
            break
    if white_bkgd:
        rgb_map = (rgb_map + (F(1) - acc_map)[:, None]).astype(F)           # integrate.cu:84-97 with bg = 1
    return rgb_map, acc_map, evaluated


from fixtures import make_kilo_scene as make_scene  # noqa: E402,F401  (synthetic scene: data only)
