"""Host side of the KiloNeRF-style path (BASELINE configs[4] ii).

Mirrors the reference's `kilonerf_cuda` bindings (cuda/pybind.cu:11-39) for the functions on this path --
get_rays_d, generate_query_indices_on_ray, network_eval_query_index, integrate,
replace_transparency_by_background_color -- with the same argument meaning, plus `KiloRenderer.render`, the
whole-frame driver the reference never had (its render_cuda_parallel is a stub, volume_renderer.py:1159-1413).
Everything runs in libnerfb200.so; there is no CPU path.
"""
import ctypes as C

import torch

from . import lib as L


def _camera(H, W, cx, cy, fx, fy, c2w, origin):
    cam = L.KiloCamera()
    cam.H, cam.W, cam.cx, cam.cy, cam.fx, cam.fy = int(H), int(W), float(cx), float(cy), float(fx), float(fy)
    c = [float(v) for v in torch.as_tensor(c2w, dtype=torch.float32).reshape(-1).tolist()]
    o = [float(v) for v in torch.as_tensor(origin, dtype=torch.float32).reshape(-1).tolist()]
    for i in range(9):
        cam.c2w[i] = c[i]
    for i in range(3):
        cam.origin[i] = o[i]
    return cam


def _grid(grid, gmin, gmax):
    g = L.KiloGrid()
    for i in range(3):
        g.res[i] = int(grid.shape[i])
        g.gmin[i] = float(gmin[i])
        g.gmax[i] = float(gmax[i])
    return g


def _march_params(dbp, spp, max_depth, min_distance, thr=0.01, white_bkgd=True, max_passes=1):
    mp = L.KiloMarchParams()
    mp.distance_between_points, mp.max_samples_per_ray, mp.max_depth_index = float(dbp), int(spp), int(max_depth)
    mp.min_distance, mp.transmittance_threshold = float(min_distance), float(thr)
    mp.white_bkgd, mp.max_passes = int(bool(white_bkgd)), int(max_passes)
    return mp


def get_rays_d(H, W, cx, cy, fx, fy, c2w, device):
    """kilonerf_cuda.get_rays_d (cuda/generate_inputs.cu:37-52): [H*W,3] unnormalised directions."""
    cam = _camera(H, W, cx, cy, fx, fy, c2w, [0, 0, 0])
    dirs = torch.empty((H * W, 3), device=device)
    L.check(L.load().nerfb200_kilo_rays_d(C.byref(cam), L.dev(dirs), L.stream_ptr()), "kilo_rays_d")
    return dirs


def generate_query_indices_on_ray(origin, directions, occupancy_grid, active_ray_mask, depth_indices, gmin, gmax,
                                  distance_between_points, max_samples_per_ray, max_depth_index, min_distance,
                                  is_initial_query):
    """kilonerf_cuda.generate_query_indices_on_ray (cuda/generate_inputs.cu:129-193).  active_ray_mask (uint8)
    and depth_indices (int32) are updated in place; returns (query_indices int32, assigned_networks int16)."""
    n = directions.shape[0]
    dev = directions.device
    q = torch.empty((n, max_samples_per_ray), dtype=torch.int32, device=dev)
    a = torch.empty((n, max_samples_per_ray), dtype=torch.int16, device=dev)
    g = _grid(occupancy_grid, gmin, gmax)
    origin = torch.as_tensor(origin, dtype=torch.float32, device=dev).contiguous()
    L.check(L.load().nerfb200_kilo_march(C.byref(g), L.dev(origin), L.dev(directions, torch.float32),
                                         L.dev(occupancy_grid, torch.int16), n, float(distance_between_points),
                                         int(max_samples_per_ray), int(max_depth_index), float(min_distance),
                                         int(bool(is_initial_query)), L.dev(q), L.dev(a),
                                         L.dev(active_ray_mask, torch.uint8), L.dev(depth_indices, torch.int32),
                                         L.stream_ptr()), "kilo_march")
    return q, a


def network_eval_query_index(query_indices, assigned_networks, params, domain_mins, domain_maxs, H, W, cx, cy, fx, fy,
                             c2w, origin, max_depth_index, min_distance, distance_between_samples):
    """kilonerf_cuda.network_eval_query_index (cuda/network_eval.cu:256-297) on ray-major query slots; the
    grouping by network that the reference does with thrust beforehand happens inside."""
    n, spp = query_indices.shape
    dev = query_indices.device
    lib = L.load()
    nn = params.shape[0]
    cam = _camera(H, W, cx, cy, fx, fy, c2w, origin)
    mp = _march_params(distance_between_samples, spp, max_depth_index, min_distance)
    ws = torch.empty(lib.nerfb200_kilo_workspace_bytes(n, spp, nn) + 256, dtype=torch.uint8, device=dev)
    base = (ws.data_ptr() + 255) & ~255
    out = torch.empty((n, spp, 4), device=dev)
    L.check(lib.nerfb200_kilo_network_eval(C.byref(cam), C.byref(mp), L.dev(query_indices, torch.int32),
                                           L.dev(assigned_networks, torch.int16), n, L.dev(params, torch.float32),
                                           L.dev(domain_mins, torch.float32), L.dev(domain_maxs, torch.float32), nn,
                                           C.c_void_p(base), ws.numel() - 256, L.dev(out), L.stream_ptr()),
            "kilo_network_eval")
    return out


def integrate(rgb_sigma, assigned_networks, dists, rgb_map, acc_map, transmittance, active_ray_mask,
              transmittance_threshold, is_initial_query):
    """kilonerf_cuda.integrate (cuda/integrate.cu:59-81); rgb_map/acc_map/transmittance/active updated in place."""
    n, spp = assigned_networks.shape
    L.check(L.load().nerfb200_kilo_integrate(L.dev(rgb_sigma, torch.float32), L.dev(assigned_networks, torch.int16),
                                             L.dev(dists, torch.float32), n, spp, float(transmittance_threshold),
                                             int(bool(is_initial_query)), L.dev(rgb_map, torch.float32),
                                             L.dev(acc_map, torch.float32), L.dev(transmittance, torch.float32),
                                             L.dev(active_ray_mask, torch.uint8), L.stream_ptr()), "kilo_integrate")


class KiloRenderer:
    """Whole-frame renderer over a grid of micro-MLPs.

    scene: occupancy grid int16 [R,R,R] of network ids (-1 = empty), params [N,6212], per-network domains,
    global domain.  render(batch) takes the reference's batch dict ('pose','intrinsics','H','W') and
    returns {'rgb_map' [H,W,3], 'acc_map' [H,W]} (the maps the reference's CUDA path produced,
    volume_renderer.py:1215-1228)."""

    def __init__(self, grid, params, domain_mins, domain_maxs, gmin, gmax, distance_between_points, max_depth_index,
                 min_distance, max_samples_per_ray=16, transmittance_threshold=0.01, white_bkgd=True, max_passes=None,
                 device="cuda:0"):
        self.device = torch.device(device)
        self.lib = L.load()
        t = lambda x, dt: torch.as_tensor(x).to(self.device, dt).contiguous()
        self.grid = t(grid, torch.int16)
        self.params = t(params, torch.float32)
        if self.params.shape[1] != self.lib.nerfb200_kilo_param_size():
            raise L.NerfB200Error("micro-MLP parameter block must have %d floats" % self.lib.nerfb200_kilo_param_size())
        self.domain_mins, self.domain_maxs = t(domain_mins, torch.float32), t(domain_maxs, torch.float32)
        self.gmin, self.gmax = [float(v) for v in gmin], [float(v) for v in gmax]
        self.dbp, self.max_depth, self.min_distance = distance_between_points, max_depth_index, min_distance
        self.spp, self.thr, self.white_bkgd = max_samples_per_ray, transmittance_threshold, white_bkgd
        self.max_passes = max_passes or -(-max_depth_index // max_samples_per_ray)
        self.stats = torch.zeros(2, dtype=torch.int64, device=self.device)
        self._ws = None

    def render(self, batch):
        H, W = int(batch["H"]), int(batch["W"])
        pose = torch.as_tensor(batch["pose"]).reshape(4, 4).float().cpu()
        K = torch.as_tensor(batch["intrinsics"]).reshape(3, 3).float().cpu()
        cam = _camera(H, W, K[0, 2], K[1, 2], K[0, 0], K[1, 1], pose[:3, :3].contiguous(), pose[:3, 3])
        g = _grid(self.grid, self.gmin, self.gmax)
        mp = _march_params(self.dbp, self.spp, self.max_depth, self.min_distance, self.thr, self.white_bkgd,
                           self.max_passes)
        nn = self.params.shape[0]
        need = self.lib.nerfb200_kilo_workspace_bytes(H * W, self.spp, nn) + 256
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        base = (self._ws.data_ptr() + 255) & ~255
        rgb = torch.empty((H * W, 3), device=self.device)
        acc = torch.empty(H * W, device=self.device)
        L.check(self.lib.nerfb200_kilo_render(C.byref(cam), C.byref(g), C.byref(mp), L.dev(self.grid),
                                              L.dev(self.params), L.dev(self.domain_mins), L.dev(self.domain_maxs), nn,
                                              C.c_void_p(base), self._ws.numel() - 256, L.dev(rgb), L.dev(acc),
                                              L.dev(self.stats), L.stream_ptr()), "kilo_render")
        return {"rgb_map": rgb.view(H, W, 3), "acc_map": acc.view(H, W)}
