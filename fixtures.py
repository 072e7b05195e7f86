"""Synthetic inputs shared by bench.py, the tests and the oracles: deterministic random-init weights in the
reference's state_dict layout, the lego test camera, and the synthetic micro-MLP scene of the a9 path.

Pure data generation -- no rendering arithmetic lives here (that is oracle/ on the CPU side and
nerf_rep_for_test_b200/ on the product side), so the product benchmark can build its inputs without touching
the oracle.
"""
import math

import numpy as np
import torch

# lego.yaml / volume_renderer.py:31-54 defaults
N_SAMPLES = 64
N_IMPORTANCE = 128
NEAR, FAR = 2.0, 6.0
L_XYZ, L_DIR = 10, 4
W_HID, D_LAYERS, SKIP = 256, 8, 4
CH_XYZ = 3 + 3 * 2 * L_XYZ   # 63
CH_DIR = 3 + 3 * 2 * L_DIR   # 27


# ---------------------------------------------------------------------------
# weights: same tensors / names / shapes as network.py:22-47 (state_dict frozen)
# ---------------------------------------------------------------------------
def state_dict_shapes():
    shapes = {}
    for prefix in ("model.", "model_fine."):
        for i in range(D_LAYERS):
            fin = CH_XYZ if i == 0 else (W_HID + CH_XYZ if i == SKIP + 1 else W_HID)
            shapes[prefix + "pts_linears.%d.weight" % i] = (W_HID, fin)
            shapes[prefix + "pts_linears.%d.bias" % i] = (W_HID,)
        shapes[prefix + "views_linears.0.weight"] = (W_HID // 2, W_HID + CH_DIR)
        shapes[prefix + "views_linears.0.bias"] = (W_HID // 2,)
        shapes[prefix + "feature_linear.weight"] = (W_HID, W_HID)
        shapes[prefix + "feature_linear.bias"] = (W_HID,)
        shapes[prefix + "alpha_linear.weight"] = (1, W_HID)
        shapes[prefix + "alpha_linear.bias"] = (1,)
        shapes[prefix + "rgb_linear.weight"] = (3, W_HID // 2)
        shapes[prefix + "rgb_linear.bias"] = (3,)
    return shapes


def make_state_dict(seed=0, sigma_gain=1.0, sigma_bias=0.0):
    """Deterministic random-init weights, independent of torch's RNG.

    Same distribution as nn.Linear's default init (U(-1/sqrt(fan_in), +), used by
    network.py:22-47) but drawn from numpy's frozen legacy MT19937 stream so the
    fixture is reproducible on any box.  `sigma_gain/sigma_bias` rescale
    alpha_linear so density is large enough to exercise compositing / ERT /
    a peaked sample_pdf (random init alone gives sigma ~ 0.02, SURVEY 8c').
    """
    rs = np.random.RandomState(seed)
    sd = {}
    for name, shape in state_dict_shapes().items():
        fan_in = shape[1] if len(shape) == 2 else None
        if fan_in is None:
            wname = name.replace(".bias", ".weight")
            fan_in = state_dict_shapes()[wname][1]
        bound = 1.0 / math.sqrt(fan_in)
        arr = rs.uniform(-bound, bound, size=shape).astype(np.float32)
        if "alpha_linear" in name:
            arr = arr * np.float32(sigma_gain)
            if name.endswith("bias"):
                arr = arr + np.float32(sigma_bias)
        sd[name] = torch.from_numpy(arr)
    return sd


# lego transforms_test.json frame 0 (data/nerf_synthetic/lego/transforms_test.json),
# camera_angle_x = 0.6911112070083618 -- the pose BASELINE.md section 3 names.
LEGO_TEST_POSE0 = [
    [-0.9999999403953552, 0.0, 0.0, 0.0],
    [0.0, -0.7341099977493286, 0.6790305972099304, 2.737260103225708],
    [0.0, 0.6790306568145752, 0.7341098785400391, 2.959291696548462],
    [0.0, 0.0, 0.0, 1.0],
]
LEGO_CAMERA_ANGLE_X = 0.6911112070083618


def lego_batch(H, W, pose=None):
    """Batch dict as blender.py:42,120-124 builds it."""
    pose = torch.tensor(LEGO_TEST_POSE0 if pose is None else pose, dtype=torch.float32)
    focal = 0.5 * W / np.tan(0.5 * LEGO_CAMERA_ANGLE_X)
    K = torch.tensor([[focal, 0, W / 2], [0, focal, H / 2], [0, 0, 1]], dtype=torch.float32)
    return {"pose": pose[None], "intrinsics": K[None], "H": H, "W": W}


# ---------------------------------------------------------------------------
# a9: synthetic scene for the KiloNeRF-style path (parameter packing of cuda/network_eval.cu:48-52)
# ---------------------------------------------------------------------------
KILO_HIDDEN, KILO_POS_EMB, KILO_DIR_EMB = 32, 63, 27
KILO_PARAM_SIZE = (KILO_POS_EMB + 1) * KILO_HIDDEN + (KILO_HIDDEN + 1) * KILO_HIDDEN + (KILO_HIDDEN + 1) * (KILO_HIDDEN + 1) + \
    (KILO_HIDDEN + KILO_DIR_EMB + 1) * KILO_HIDDEN + (KILO_HIDDEN + 1) * 3          # 6212


def make_kilo_scene(seed=0, net_res=16, grid_res=128, blob_radius=0.7, sigma_scale=40.0):
    """Synthetic config-5 scene: [-1.5,1.5]^3 domain, net_res^3 micro-MLPs with random weights, occupancy =
    cells whose centre lies in a sphere (network id = the net_res^3 cell containing it), -1 elsewhere."""
    rs = np.random.RandomState(seed)
    gmin, gmax = np.array([-1.5, -1.5, -1.5], np.float32), np.array([1.5, 1.5, 1.5], np.float32)
    nn = net_res ** 3
    params = np.zeros((nn, KILO_PARAM_SIZE), np.float32)
    o = 0
    for nin, nout in ((KILO_POS_EMB, KILO_HIDDEN), (KILO_HIDDEN, KILO_HIDDEN), (KILO_HIDDEN, KILO_HIDDEN + 1), (KILO_HIDDEN + KILO_DIR_EMB, KILO_HIDDEN), (KILO_HIDDEN, 3)):
        bound = 1.0 / np.sqrt(nin)
        size = nout + nin * nout
        params[:, o:o + size] = rs.uniform(-bound, bound, size=(nn, size)).astype(np.float32)
        if nout == KILO_HIDDEN + 1:      # density output: positive and large so early termination happens
            params[:, o] = np.float32(0.5)
            params[:, o + nout:o + size:nout] *= np.float32(0.1)
            params[:, o] *= np.float32(sigma_scale)
        o += size
    c = (np.arange(grid_res, dtype=np.float32) + np.float32(0.5)) / np.float32(grid_res)
    g = np.stack(np.meshgrid(c, c, c, indexing="ij"), -1) * (gmax - gmin) + gmin
    occ = np.sqrt((g * g).sum(-1)) <= blob_radius
    cell = np.arange(grid_res) // (grid_res // net_res)
    net_id = (cell[:, None, None] * net_res + cell[None, :, None]) * net_res + cell[None, None, :]
    grid = np.where(occ, net_id, -1).astype(np.int16)
    k = np.arange(net_res, dtype=np.float32)
    lo = np.stack(np.meshgrid(k, k, k, indexing="ij"), -1).reshape(-1, 3) / np.float32(net_res) * (gmax - gmin) + gmin
    hi = lo + (gmax - gmin) / np.float32(net_res)
    return dict(grid=grid, params=params, domain_mins=lo.astype(np.float32), domain_maxs=hi.astype(np.float32), gmin=gmin, gmax=gmax)
