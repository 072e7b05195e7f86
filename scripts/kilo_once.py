"""Render the bench's a9 scene a few times (for ncu / timing breakdowns of the KiloNeRF-style path)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import kilo_oracle as KO, nerf_oracle as O
from nerf_rep_for_test_b200 import kilo
dev = torch.device("cuda:0")
sc = KO.make_scene(seed=0, net_res=16, grid_res=128, blob_radius=1.0)
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
kr = kilo.KiloRenderer(sc["grid"], sc["params"], sc["domain_mins"], sc["domain_maxs"], sc["gmin"], sc["gmax"], 4.0 / 384, 384, 2.0,
                       max_samples_per_ray=spp, device=dev)
b = O.lego_batch(800, 800)
b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(n):
    kr.render(b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); kr.render(b); e1.record(); torch.cuda.synchronize()
print("spp", spp, "ms", e0.elapsed_time(e1), "passes", kr.max_passes)
