"""Weight holder with the reference's frozen state_dict layout.

Mirrors /root/reference/src/models/nerf/network.py:9-47 (NeRF) and :126-159
(Network): same submodule names, shapes and default nn.Linear init, so
`Network().state_dict()` has exactly the reference's 48 keys (`model.*`,
`model_fine.*`) and checkpoints are interchangeable (net_utils.py:323-380 stores
them under 'net').  The Renderer never calls these modules' forward(): it reads
the parameter tensors and runs the CUDA kernels.  When the reference scaffold is
importable, its own `Network` can be passed to `Renderer` instead.
"""
import torch
import torch.nn as nn

CH_XYZ, CH_DIR = 63, 27


class NeRF(nn.Module):
    def __init__(self, D=8, W=256, input_ch=CH_XYZ, input_ch_views=CH_DIR, skips=(4,), use_viewdirs=True):
        super().__init__()
        if not use_viewdirs:
            raise NotImplementedError("the B200 path implements use_viewdirs=True (lego.yaml:20)")
        self.D, self.W, self.input_ch, self.input_ch_views = D, W, input_ch, input_ch_views
        self.skips, self.use_viewdirs = list(skips), use_viewdirs
        self.pts_linears = nn.ModuleList(
            [nn.Linear(input_ch, W)]
            + [nn.Linear(W, W) if i not in self.skips else nn.Linear(W + input_ch, W) for i in range(D - 1)])
        self.views_linears = nn.ModuleList([nn.Linear(input_ch_views + W, W // 2)])
        self.feature_linear = nn.Linear(W, W)
        self.alpha_linear = nn.Linear(W, 1)
        self.rgb_linear = nn.Linear(W // 2, 3)

    def forward(self, x):
        raise RuntimeError("NeRF.forward is not part of the B200 path: query it through Renderer "
                           "(libnerfb200 mlp_forward); there is no PyTorch fallback")


class Network(nn.Module):
    """Same attributes the reference Renderer captures (volume_renderer.py:41-45)."""

    def __init__(self, device=None):
        super().__init__()
        self.N_samples, self.N_importance, self.chunk = 64, 128, 4096
        self.white_bkgd, self.use_viewdirs = 1, True
        self.device = torch.device(device) if device is not None else torch.device(
            "cuda" if torch.cuda.is_available() else "cpu")
        self.input_ch, self.input_ch_views = CH_XYZ, CH_DIR
        self.embed_fn = self.embeddirs_fn = None   # positional encoding is fused into the MLP kernel
        self.model = NeRF()
        self.model_fine = NeRF()
