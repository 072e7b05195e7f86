"""Kernel-level Python wrappers over the C ABI (one function per entry point).

Inputs/outputs are CUDA tensors; every wrapper allocates its outputs with torch and
enqueues the kernel on torch's current stream.  No wrapper has a non-CUDA code path.
"""
import ctypes as C

import torch

from . import lib as L

F32 = torch.float32


def _f(t, device=None):
    if t is None:
        return None
    if not t.is_cuda:
        raise L.NerfB200Error("libnerfb200 needs CUDA tensors (got %s); there is no CPU path" % t.device)
    return t.to(F32).contiguous()


def raygen(pose, intrinsics, H, W):
    pose, intrinsics = _f(pose.reshape(4, 4)), _f(intrinsics.reshape(3, 3))
    ro = torch.empty((H * W, 3), device=pose.device)
    rd = torch.empty((H * W, 3), device=pose.device)
    L.check(L.load().nerfb200_raygen(L.dev(pose), L.dev(intrinsics), H, W, L.dev(ro), L.dev(rd), L.stream_ptr()),
            "raygen")
    return ro, rd


def sample_coarse(z_table, n_rays, perturb=False, seed=0):
    z_table = _f(z_table)
    S = z_table.numel()
    z = torch.empty((n_rays, S), device=z_table.device)
    L.check(L.load().nerfb200_sample_coarse(L.dev(z_table), n_rays, S, int(perturb), seed, L.dev(z),
                                           L.stream_ptr()), "sample_coarse")
    return z


class PackedWeights:
    """1024-byte aligned device buffer holding one model's weights in a kernel layout."""

    def __init__(self, tensors, mode, device):
        """tensors: dict name -> tensor with the 24 per-model names of network.py (no prefix)."""
        lib = L.load()
        self.mode = mode
        nbytes = lib.nerfb200_packed_weights_bytes(mode)
        self.buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
        self.base = (self.buf.data_ptr() + 1023) & ~1023
        keep = {k: _f(v.detach().to(device)) for k, v in tensors.items()}
        w = L.MlpWeights()
        for i in range(8):
            w.pts_w[i] = keep["pts_linears.%d.weight" % i].data_ptr()
            w.pts_b[i] = keep["pts_linears.%d.bias" % i].data_ptr()
        w.views_w, w.views_b = keep["views_linears.0.weight"].data_ptr(), keep["views_linears.0.bias"].data_ptr()
        w.feature_w, w.feature_b = keep["feature_linear.weight"].data_ptr(), keep["feature_linear.bias"].data_ptr()
        w.alpha_w, w.alpha_b = keep["alpha_linear.weight"].data_ptr(), keep["alpha_linear.bias"].data_ptr()
        w.rgb_w, w.rgb_b = keep["rgb_linear.weight"].data_ptr(), keep["rgb_linear.bias"].data_ptr()
        L.check(lib.nerfb200_pack_weights(C.byref(w), mode, C.c_void_p(self.base), L.stream_ptr()), "pack_weights")
        torch.cuda.current_stream().synchronize()  # `keep` may be freed after this returns

    @property
    def ptr(self):
        return C.c_void_p(self.base)


def pack_from_state_dict(sd, prefix, mode, device):
    return PackedWeights({k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}, mode, device)


def mlp_forward(packed, rays_o, rays_d, z_vals):
    rays_o, rays_d, z_vals = _f(rays_o), _f(rays_d), _f(z_vals)
    n, S = z_vals.shape
    raw = torch.empty((n, S, 4), device=z_vals.device)
    L.check(L.load().nerfb200_mlp_forward(packed.ptr, packed.mode, L.dev(rays_o), L.dev(rays_d), L.dev(z_vals), n, S,
                                         L.dev(raw), L.stream_ptr()), "mlp_forward")
    return raw


class TrainStore:
    """What the training forward keeps for nerfb200_mlp_backward: activation tile images + relu sign bits
    (layout: csrc/train_layout.cuh)."""

    # block index of each plane inside a 128-row tile of `acts`, and its width in 64-column blocks
    PLANES = {"pe": (0, 1), "dpe": (1, 1), "hv": (34, 2)}
    PLANES.update({"h%d" % i: (2 + 4 * i, 4) for i in range(8)})
    BLOCKS, BLOCK_BYTES = 36, 16384
    DACT_PLANES = {"d9": (32, 4)}
    DACT_PLANES.update({"dpre%d" % i: (4 * i, 4) for i in range(8)})
    DACT_BLOCKS = 36

    def __init__(self, n_rows, device):
        lib = L.load()
        self.n_rows = n_rows
        self.acts = torch.empty(lib.nerfb200_train_acts_bytes(n_rows), dtype=torch.uint8, device=device)
        self.masks = torch.empty(lib.nerfb200_train_masks_bytes(n_rows), dtype=torch.uint8, device=device)

    def plane(self, name):
        """[n_rows, 64*width] bf16 copy of one saved plane (tests / debugging only)."""
        blk, width = self.PLANES[name]
        return untile(self.acts, self.n_rows, self.BLOCKS, blk, width)


def untile(buf, n_rows, blocks_per_tile, first_block, n_blocks):
    """Undo the SWIZZLE_128B tile-image layout: uint8 buffer -> [n_rows, 64*n_blocks] bf16."""
    n_tiles = (n_rows + 127) // 128
    t = buf.view(torch.bfloat16).view(n_tiles, blocks_per_tile, 128, 8, 8)[:, first_block:first_block + n_blocks]
    r = torch.arange(128, device=buf.device)
    u = torch.arange(8, device=buf.device)
    phys = (u[None, :] ^ (r[:, None] & 7))                       # [128, 8]: physical 16-byte unit of logical unit u
    idx = phys[None, None, :, :, None].expand(n_tiles, n_blocks, 128, 8, 8)
    t = torch.gather(t, 3, idx)                                  # logical unit order
    return t.permute(0, 2, 1, 3, 4).reshape(n_tiles * 128, n_blocks * 64)[:n_rows].contiguous()


def mlp_forward_train(packed, rays_o, rays_d, z_vals):
    """BF16 mode: (raw [n,S,4] fp32, TrainStore) -- stage outputs and relu sign bits kept for backward."""
    rays_o, rays_d, z_vals = _f(rays_o), _f(rays_d), _f(z_vals)
    n, S = z_vals.shape
    raw = torch.empty((n, S, 4), device=z_vals.device)
    store = TrainStore(n * S, z_vals.device)
    L.check(L.load().nerfb200_mlp_forward_train(packed.ptr, packed.mode, L.dev(rays_o), L.dev(rays_d), L.dev(z_vals),
                                               n, S, L.dev(raw), L.dev(store.acts), L.dev(store.masks),
                                               L.stream_ptr()), "mlp_forward_train")
    return raw, store


GRAD_SHAPES = [(256, 63), (256,)] + [(256, 256), (256,)] * 4 + [(256, 319), (256,)] + [(256, 256), (256,)] * 2 + \
    [(128, 283), (128,), (256, 256), (256,), (1, 256), (1,), (3, 128), (3,)]


def mlp_backward(bwd, g_raw, store, keep_workspace=None, grads=None):
    """nerfb200_mlp_backward: the 24 gradients (order of training._NAMES) of one model given dL/draw.
    bwd = Renderer.packed_bwd(which): (packed W^T image pointer, nerfb200_mlp_weights struct of the fp32 tensors).
    `grads`: optional list of 24 preallocated fp32 tensors (e.g. views of one flat gradient buffer) to overwrite.
    `keep_workspace`: optional dict that receives the workspace tensor (tests read the dgrad planes)."""
    packed_bwd_ptr, wstruct = bwd
    g_raw = _f(g_raw).reshape(-1, 4)
    lib = L.load()
    n_rows = store.n_rows
    if g_raw.shape[0] != n_rows:
        raise L.NerfB200Error("mlp_backward: g_raw has %d rows, the forward saved %d" % (g_raw.shape[0], n_rows))
    dev = g_raw.device
    ws_bytes = lib.nerfb200_mlp_backward_workspace_bytes(n_rows)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    if grads is None:
        grads = [torch.empty(sh, device=dev) for sh in GRAD_SHAPES]
    else:
        for t, sh in zip(grads, GRAD_SHAPES):
            if tuple(t.shape) != tuple(sh) or t.dtype != torch.float32 or not t.is_contiguous() or t.device != dev:
                raise L.NerfB200Error("mlp_backward: gradient buffer %s does not match %s fp32 contiguous" % (tuple(t.shape), sh))
    g = L.MlpGrads()
    for i in range(8):
        g.pts_w[i], g.pts_b[i] = grads[2 * i].data_ptr(), grads[2 * i + 1].data_ptr()
    (g.views_w, g.views_b, g.feature_w, g.feature_b, g.alpha_w, g.alpha_b, g.rgb_w, g.rgb_b) = \
        [t.data_ptr() for t in grads[16:24]]
    L.check(lib.nerfb200_mlp_backward(packed_bwd_ptr, C.byref(wstruct), L.dev(g_raw), L.dev(store.acts),
                                      L.dev(store.masks), n_rows, L.dev(ws), ws_bytes, C.byref(g), L.stream_ptr()),
            "mlp_backward")
    if keep_workspace is not None:
        keep_workspace["ws"] = ws
    return grads


def mlp_backward_input(bwd, ws, rays_o, rays_d, z_vals):
    """g_z [n,S]: dL/d z_vals through the MLP input (bf16 path, reference graph).  `ws` is the workspace tensor of the
    preceding mlp_backward call on the same rows (keep_workspace)."""
    _, wstruct = bwd
    rays_o, rays_d, z_vals = _f(rays_o), _f(rays_d), _f(z_vals)
    n, S = z_vals.shape
    g_z = torch.empty((n, S), device=z_vals.device)
    L.check(L.load().nerfb200_mlp_backward_input(C.byref(wstruct), L.dev(ws), L.dev(rays_o), L.dev(rays_d), L.dev(z_vals),
                                                n, S, L.dev(g_z), L.stream_ptr()), "mlp_backward_input")
    return g_z


def mse_pair_grad(rgb0, rgb, target):
    """(loss [1], g_rgb0, g_rgb) of mse(rgb0, t) + mse(rgb, t) in one kernel."""
    rgb0, rgb, target = _f(rgb0), _f(rgb), _f(target)
    n = rgb0.shape[0]
    g0, g1 = torch.empty_like(rgb0), torch.empty_like(rgb)
    loss = torch.empty(1, device=rgb0.device)
    L.check(L.load().nerfb200_mse_pair_grad(L.dev(rgb0), L.dev(rgb), L.dev(target), n, L.dev(g0), L.dev(g1), L.dev(loss),
                                           L.stream_ptr()), "mse_pair_grad")
    return loss, g0, g1


def adam_clip_step(params, grads, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step, clip_value=0.0, grad_scale=1.0):
    """clip_grad_value_ + Adam over flat fp32 CUDA buffers, in place (one kernel)."""
    for t in (params, grads, exp_avg, exp_avg_sq):
        if not (t.is_cuda and t.dtype == F32 and t.is_contiguous() and t.numel() == params.numel()):
            raise L.NerfB200Error("adam_clip_step: flat contiguous fp32 CUDA buffers of equal size expected")
    L.check(L.load().nerfb200_adam_clip_step(L.dev(params), L.dev(grads), L.dev(exp_avg), L.dev(exp_avg_sq), params.numel(),
                                            float(lr), float(beta1), float(beta2), float(eps), int(step), float(clip_value),
                                            float(grad_scale), L.stream_ptr()), "adam_clip_step")


def mlp_forward_stages(packed, rays_o, rays_d, z_vals):
    """BF16 mode diagnostic: (raw, stage_dump [10,128,256]) -- fp32 stage outputs of rows 0..127."""
    rays_o, rays_d, z_vals = _f(rays_o), _f(rays_d), _f(z_vals)
    n, S = z_vals.shape
    raw = torch.empty((n, S, 4), device=z_vals.device)
    dump = torch.zeros((10, 128, 256), device=z_vals.device)
    L.check(L.load().nerfb200_mlp_forward_stages(packed.ptr, packed.mode, L.dev(rays_o), L.dev(rays_d), L.dev(z_vals),
                                                n, S, L.dev(raw), L.dev(dump), L.stream_ptr()), "mlp_forward_stages")
    return raw, dump


def composite_forward(raw, z_vals, rays_d, variant=L.COMPOSITE_PLAIN, ert_threshold=0.01, white_bkgd=True,
                      compat_chunk=2048, want_weights=True, keep_bits=None):
    """keep_bits (int32/uint32 words from ess_compact, optional): rows with a cleared bit have zero density and
    their raw entries are never read (nerfb200_composite_forward_masked)."""
    raw, z_vals, rays_d = _f(raw), _f(z_vals), _f(rays_d)
    n, S = z_vals.shape
    d = raw.device
    rgb, disp, acc, depth = (torch.empty((n, 3), device=d), torch.empty(n, device=d), torch.empty(n, device=d),
                             torch.empty(n, device=d))
    w = torch.empty((n, S), device=d) if want_weights else None
    if keep_bits is not None:
        if keep_bits.numel() * 32 < n * S:
            raise L.NerfB200Error("composite_forward: keep_bits has %d words, need %d" % (keep_bits.numel(), (n * S + 31) // 32))
        L.check(L.load().nerfb200_composite_forward_masked(
            L.dev(raw), L.dev(z_vals), L.dev(rays_d), L.dev(keep_bits, torch.int32), n, S, variant, ert_threshold,
            int(white_bkgd), compat_chunk, L.dev(rgb), L.dev(disp), L.dev(acc), L.dev(depth), L.dev(w), L.stream_ptr()),
            "composite_forward_masked")
        return rgb, disp, acc, w, depth
    L.check(L.load().nerfb200_composite_forward(L.dev(raw), L.dev(z_vals), L.dev(rays_d), n, S, variant,
                                               ert_threshold, int(white_bkgd), compat_chunk, L.dev(rgb), L.dev(disp),
                                               L.dev(acc), L.dev(depth), L.dev(w), L.stream_ptr()),
            "composite_forward")
    return rgb, disp, acc, w, depth   # same order as _raw2outputs (:357)


def sigma_noise(raw, std, seed=0):
    """raw[..., 3] += N(0,1) * std in place (raw_noise_std, volume_renderer.py:310-314); returns raw."""
    if raw.dtype != torch.float32 or not raw.is_contiguous() or raw.shape[-1] != 4:
        raise L.NerfB200Error("sigma_noise: raw must be a contiguous fp32 [..., 4] tensor")
    L.check(L.load().nerfb200_sigma_noise(L.dev(raw), raw.numel() // 4, float(std), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                         L.stream_ptr()), "sigma_noise")
    return raw


def composite_backward(raw, z_vals, rays_d, g_rgb=None, g_acc=None, g_depth=None, g_weights=None, white_bkgd=True):
    raw, z_vals, rays_d = _f(raw), _f(z_vals), _f(rays_d)
    n, S = z_vals.shape
    g_raw = torch.empty_like(raw)
    L.check(L.load().nerfb200_composite_backward(L.dev(raw), L.dev(z_vals), L.dev(rays_d), n, S, int(white_bkgd),
                                                L.dev(_f(g_rgb)), L.dev(_f(g_acc)), L.dev(_f(g_depth)),
                                                L.dev(_f(g_weights)), L.dev(g_raw), L.stream_ptr()),
            "composite_backward")
    return g_raw


def composite_backward_z(raw, z_vals, rays_d, g_rgb=None, g_acc=None, g_depth=None, g_weights=None, white_bkgd=True):
    """(g_raw, g_z): composite_backward plus dL/d z_vals through the interval lengths and depth_map."""
    raw, z_vals, rays_d = _f(raw), _f(z_vals), _f(rays_d)
    n, S = z_vals.shape
    g_raw = torch.empty_like(raw)
    g_z = torch.empty_like(z_vals)
    L.check(L.load().nerfb200_composite_backward_z(L.dev(raw), L.dev(z_vals), L.dev(rays_d), n, S, int(white_bkgd),
                                                  L.dev(_f(g_rgb)), L.dev(_f(g_acc)), L.dev(_f(g_depth)),
                                                  L.dev(_f(g_weights)), L.dev(g_raw), L.dev(g_z), L.stream_ptr()),
            "composite_backward_z")
    return g_raw, g_z


def sample_pdf_backward(z_coarse, weights, u, g_z_all):
    """g_weights [n,S] of the coarse pass given dL/d z_all [n,S+n_u] (reference graph: sampler not detached)."""
    z_coarse, weights, u, g_z_all = _f(z_coarse), _f(weights), _f(u), _f(g_z_all)
    n, S = z_coarse.shape
    n_u = u.shape[-1]
    if tuple(g_z_all.shape) != (n, S + n_u):
        raise L.NerfB200Error("sample_pdf_backward: g_z_all must be [%d,%d]" % (n, S + n_u))
    g_w = torch.empty((n, S), device=z_coarse.device)
    L.check(L.load().nerfb200_sample_pdf_backward(L.dev(z_coarse), L.dev(weights), L.dev(u), int(u.dim() == 2), n, S, n_u,
                                                 L.dev(g_z_all), L.dev(g_w), L.stream_ptr()), "sample_pdf_backward")
    return g_w


def weights_struct(tensors):
    """nerfb200_mlp_weights over a list of the 24 fp32 CUDA tensors of one model (order of training._NAMES)."""
    w = L.MlpWeights()
    ts = [_f(t.detach()) for t in tensors]
    for i in range(8):
        w.pts_w[i], w.pts_b[i] = ts[2 * i].data_ptr(), ts[2 * i + 1].data_ptr()
    (w.views_w, w.views_b, w.feature_w, w.feature_b, w.alpha_w, w.alpha_b, w.rgb_w, w.rgb_b) = [t.data_ptr() for t in ts[16:24]]
    return w, ts


def mlp_forward_train_fp32(tensors, rays_o, rays_d, z_vals):
    """fp32-accurate training forward: (raw [n,S,4], acts buffer).  tensors: the 24 fp32 tensors of one model."""
    rays_o, rays_d, z_vals = _f(rays_o), _f(rays_d), _f(z_vals)
    n, S = z_vals.shape
    lib = L.load()
    raw = torch.empty((n, S, 4), device=z_vals.device)
    acts = torch.empty(lib.nerfb200_train_fp32_acts_bytes(n * S), dtype=torch.uint8, device=z_vals.device)
    w, keep = weights_struct(tensors)
    L.check(lib.nerfb200_mlp_forward_train_fp32(C.byref(w), L.dev(rays_o), L.dev(rays_d), L.dev(z_vals), n, S, L.dev(raw),
                                                L.dev(acts), L.stream_ptr()), "mlp_forward_train_fp32")
    return raw, acts


def mlp_backward_fp32(tensors, g_raw, acts, rays_d, n, S, grads=None, want_g_z=False):
    """(24 gradients, g_z [n,S] or None) of nerfb200_mlp_backward_fp32."""
    lib = L.load()
    g_raw = _f(g_raw).reshape(-1, 4)
    rays_d = _f(rays_d)
    dev = g_raw.device
    if g_raw.shape[0] != n * S:
        raise L.NerfB200Error("mlp_backward_fp32: g_raw has %d rows, expected %d" % (g_raw.shape[0], n * S))
    ws_bytes = lib.nerfb200_train_fp32_workspace_bytes(n * S)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    if grads is None:
        grads = [torch.empty(sh, device=dev) for sh in GRAD_SHAPES]
    g = L.MlpGrads()
    for i in range(8):
        g.pts_w[i], g.pts_b[i] = grads[2 * i].data_ptr(), grads[2 * i + 1].data_ptr()
    (g.views_w, g.views_b, g.feature_w, g.feature_b, g.alpha_w, g.alpha_b, g.rgb_w, g.rgb_b) = \
        [t.data_ptr() for t in grads[16:24]]
    g_z = torch.empty((n, S), device=dev) if want_g_z else None
    w, keep = weights_struct(tensors)
    L.check(lib.nerfb200_mlp_backward_fp32(C.byref(w), L.dev(g_raw), L.dev(acts), L.dev(rays_d), n, S, L.dev(ws), ws_bytes,
                                           C.byref(g), L.dev(g_z), L.stream_ptr()), "mlp_backward_fp32")
    return grads, g_z


def sample_from_cdf(cdf, bins, u):
    cdf, bins, u = _f(cdf), _f(bins), _f(u)
    n, nb = cdf.shape
    per_ray = int(u.dim() == 2)
    n_u = u.shape[-1]
    samples = torch.empty((n, n_u), device=cdf.device)
    inds = torch.empty((n, n_u), dtype=torch.int32, device=cdf.device)
    L.check(L.load().nerfb200_sample_from_cdf(L.dev(cdf), L.dev(bins), L.dev(u), per_ray, n, nb, n_u, L.dev(samples),
                                             L.dev(inds), L.stream_ptr()), "sample_from_cdf")
    return samples, inds


def sample_pdf_merge(z_coarse, weights, u, want_aux=True):
    z_coarse, weights, u = _f(z_coarse), _f(weights), _f(u)
    n, S = z_coarse.shape
    per_ray = int(u.dim() == 2)
    n_u = u.shape[-1]
    d = z_coarse.device
    z_all = torch.empty((n, S + n_u), device=d)
    zs = torch.empty((n, n_u), device=d) if want_aux else None
    inds = torch.empty((n, n_u), dtype=torch.int32, device=d) if want_aux else None
    cdf = torch.empty((n, S - 1), device=d) if want_aux else None
    L.check(L.load().nerfb200_sample_pdf_merge(L.dev(z_coarse), L.dev(weights), L.dev(u), per_ray, n, S, n_u,
                                              L.dev(z_all), L.dev(zs), L.dev(inds), L.dev(cdf), L.stream_ptr()),
            "sample_pdf_merge")
    return z_all, zs, inds, cdf


def ess_resample_compat(grid_u8, rays_o, rays_d, z_table, chunk=2048):
    """The reference's literal ESS resampling (one shared row per `chunk` rays): z_vals [n,S]."""
    rays_o, rays_d, z_table = _f(rays_o), _f(rays_d), _f(z_table)
    n, S = rays_o.shape[0], z_table.numel()
    z = torch.empty((n, S), device=rays_o.device)
    scratch = torch.empty(max(n, 1), dtype=torch.int64, device=rays_o.device)
    L.check(L.load().nerfb200_ess_resample_compat(L.dev(grid_u8, torch.uint8), grid_u8.shape[0], L.dev(rays_o), L.dev(rays_d),
                                                 n, S, chunk, L.dev(z_table), L.dev(z), L.dev(scratch), L.stream_ptr()),
            "ess_resample_compat")
    return z


def jitter_rows(z_vals, seed=0):
    """Stratified jitter of placed per-ray depths from each row's own mid-points (:1079-1085); returns a new tensor."""
    z = _f(z_vals).clone()
    n, S = z.shape
    L.check(L.load().nerfb200_jitter_rows(L.dev(z), n, S, seed, L.stream_ptr()), "jitter_rows")
    return z


def ess_resample(grid_u8, rays_o, rays_d, z_vals):
    rays_o, rays_d = _f(rays_o), _f(rays_d)
    z = _f(z_vals).clone()
    n, S = z.shape
    n_empty = torch.empty(n, dtype=torch.int32, device=z.device)
    L.check(L.load().nerfb200_ess_resample(L.dev(grid_u8, torch.uint8), grid_u8.shape[0], L.dev(rays_o), L.dev(rays_d),
                                          n, S, L.dev(z), L.dev(n_empty), L.stream_ptr()), "ess_resample")
    return z, n_empty


def ess_update(grid_u8, rays_o, rays_d, z_vals, raw, weights, use_origin=False):
    n, S = z_vals.shape
    L.check(L.load().nerfb200_ess_update(L.dev(grid_u8, torch.uint8), grid_u8.shape[0], L.dev(_f(rays_o)),
                                        L.dev(_f(rays_d)), L.dev(_f(z_vals)), L.dev(_f(raw)), L.dev(_f(weights)), n, S,
                                        int(use_origin), L.stream_ptr()), "ess_update")
    return grid_u8


def ess_compact(grid_u8, rays_o, rays_d, z_vals, z_term=None, want_bits=False):
    """(row_ids int32 [n*S] with the first n_active entries valid, n_active int32 [1] on the device); with
    want_bits also the keep-bit words (int32 [ceil(n*S/32)], bit m&31 of word m>>5 = row m listed)."""
    rays_o, rays_d, z_vals = _f(rays_o), _f(rays_d), _f(z_vals)
    n, S = z_vals.shape
    row_ids = torch.empty(n * S, dtype=torch.int32, device=z_vals.device)
    n_active = torch.zeros(1, dtype=torch.int32, device=z_vals.device)
    bits = torch.empty((n * S + 31) // 32, dtype=torch.int32, device=z_vals.device) if want_bits else None
    L.check(L.load().nerfb200_ess_compact(L.dev(grid_u8, torch.uint8), grid_u8.shape[0], L.dev(rays_o), L.dev(rays_d),
                                         L.dev(z_vals), L.dev(_f(z_term)), n, S, L.dev(row_ids), L.dev(n_active),
                                         L.dev(bits, torch.int32) if bits is not None else None,
                                         L.stream_ptr()), "ess_compact")
    return (row_ids, n_active, bits) if want_bits else (row_ids, n_active)


def mlp_forward_sparse(packed, rays_o, rays_d, z_vals, row_ids, n_active):
    rays_o, rays_d, z_vals = _f(rays_o), _f(rays_d), _f(z_vals)
    n, S = z_vals.shape
    raw = torch.empty((n, S, 4), device=z_vals.device)
    L.check(L.load().nerfb200_mlp_forward_sparse(packed.ptr, packed.mode, L.dev(rays_o), L.dev(rays_d), L.dev(z_vals),
                                                n, S, L.dev(row_ids), L.dev(n_active), L.dev(raw), L.stream_ptr()),
            "mlp_forward_sparse")
    return raw


def ray_cull(rays_o, rays_d, z_table, box_lo, box_hi):
    """uint8 [n]: 1 where the ray segment over [z_table[0], z_table[-1]] meets the box (nerfb200_ray_cull)."""
    import ctypes as C
    rays_o, rays_d, z_table = _f(rays_o), _f(rays_d), _f(z_table)
    n = rays_o.shape[0]
    active = torch.empty(n, dtype=torch.uint8, device=rays_o.device)
    lo, hi = (C.c_float * 3)(*[float(v) for v in box_lo]), (C.c_float * 3)(*[float(v) for v in box_hi])
    L.check(L.load().nerfb200_ray_cull(L.dev(rays_o), L.dev(rays_d), n, L.dev(z_table), z_table.numel(), lo, hi,
                                      L.dev(active, torch.uint8), L.stream_ptr()), "ray_cull")
    return active


def ert_depth(weights, z_vals, thr):
    weights, z_vals = _f(weights), _f(z_vals)
    n, S = z_vals.shape
    zt = torch.empty(n, device=z_vals.device)
    L.check(L.load().nerfb200_ert_depth(L.dev(weights), L.dev(z_vals), n, S, float(thr), L.dev(zt), L.stream_ptr()),
            "ert_depth")
    return zt


# ---- every wrapper runs on the device that owns its tensors (the C ABI works on the CURRENT device): a network on
# cuda:1 with cuda:0 current would otherwise launch on the wrong GPU
def _device_of(args, kwargs):
    for a in list(args) + list(kwargs.values()):
        if torch.is_tensor(a) and a.is_cuda:
            return a.device
        if isinstance(a, (list, tuple)) and a and torch.is_tensor(a[0]) and a[0].is_cuda:
            return a[0].device
        if isinstance(a, PackedWeights):
            return a.buf.device
        if isinstance(a, TrainStore):
            return a.acts.device
    return None


def _on_tensor_device(fn):
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = _device_of(args, kwargs)
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapped


for _name, _fn in list(globals().items()):
    if callable(_fn) and not _name.startswith("_") and getattr(_fn, "__module__", None) == __name__ and \
            not isinstance(_fn, type) and _name not in ("untile",):
        globals()[_name] = _on_tensor_device(_fn)
del _name, _fn
