"""Render the bench's ESS-skip + ERT configuration (synthetic opaque blob) a few times (timing breakdowns)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fixtures as FX
from nerf_rep_for_test_b200 import Network, RenderConfig, Renderer
dev = torch.device("cuda:0")
sd5 = FX.make_state_dict(6, 300.0, 6.0)
for k in list(sd5):
    if k.startswith("model_fine."):
        sd5[k] = sd5["model." + k[len("model_fine."):]].clone()
net5 = Network(device=dev); net5.load_state_dict(sd5); net5.to(dev).eval()
res = 128
gc = torch.stack(torch.meshgrid([torch.arange(res, device=dev)] * 3, indexing="ij"), -1).float() / (res - 1) * 2 - 1
blob = torch.norm(gc, dim=-1) <= 0.35
r = Renderer(net5, RenderConfig(perturb=0, enable_ess=True, enable_ert=True), mode="bf16")
r.occupancy_grid = blob
r.ess_mode = "skip"
b = FX.lego_batch(800, 800)
b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(n):
    r.render(b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); r.render(b); e1.record(); torch.cuda.synchronize()
print("skip render ms", e0.elapsed_time(e1))
