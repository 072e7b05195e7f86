"""One coarse-sized and one fine-sized launch of the fp32-accurate tensor-core MLP kernel (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fixtures as FX
from nerf_rep_for_test_b200 import lib as L, ops
dev = torch.device("cuda:0")
sd = FX.make_state_dict(0)
b = FX.lego_batch(800, 800)
ro, rd = ops.raygen(b["pose"][0].to(dev), b["intrinsics"][0].to(dev), 800, 800)
ro, rd = ro[:32560].contiguous(), rd[:32560].contiguous()
packed = ops.pack_from_state_dict(sd, "model_fine.", L.MODE_FP32_TC, dev)
z = torch.sort(torch.rand(32560, 192, device=dev) * 4 + 2, -1)[0]
for _ in range(2):
    ops.mlp_forward(packed, ro, rd, z)
torch.cuda.synchronize()
