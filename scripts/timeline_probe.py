import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import nerf_oracle as O
from nerf_rep_for_test_b200 import lib as L, ops
dev = torch.device("cuda:0")
sd = O.make_state_dict(0)
n, S = 8192, 192
b = O.lego_batch(800, 800)
ro, rd = ops.raygen(b["pose"].to(dev), b["intrinsics"].to(dev), 800, 800)
z = ops.sample_coarse(torch.linspace(2, 6, S, device=dev), n)
packed = ops.pack_from_state_dict(sd, "model_fine.", L.MODE_BF16, dev)
os.environ.pop("NERFB200_TIMELINE", None)
for _ in range(2):
    ops.mlp_forward(packed, ro[:n], rd[:n], z)
torch.cuda.synchronize()
os.environ["NERFB200_TIMELINE"] = sys.argv[1]
ops.mlp_forward(packed, ro[:n], rd[:n], z)
torch.cuda.synchronize()
