"""One launch of sample_pdf_merge at full-frame size (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerf_rep_for_test_b200 import ops
dev = torch.device("cuda:0")
N, S, U = 640000, 64, 128
g = torch.Generator(device=dev).manual_seed(0)
ztab = torch.linspace(2, 6, S, device=dev)
z_c = ops.sample_coarse(ztab, N)
raw_c = torch.randn(N, S, 4, device=dev, generator=g) * 0.3
rd = torch.nn.functional.normalize(torch.randn(N, 3, device=dev, generator=g), dim=-1)
w_c = ops.composite_forward(raw_c, z_c, rd)[3]
u = torch.linspace(0, 1, U, device=dev)
for _ in range(2):
    ops.sample_pdf_merge(z_c, w_c, u, want_aux=False)
torch.cuda.synchronize()
