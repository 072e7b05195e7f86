"""Frame time of the a9 path on the bench's scene (median of 5) and a dump of the rgb map, used for the A/B record of the
micro-MLP kernels in profiles/r02_kilo_ab.txt (the variants were separate builds; the CUDA-core kernel is gone)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import kilo_oracle as KO, nerf_oracle as O
from nerf_rep_for_test_b200 import kilo
dev = torch.device("cuda:0")
sc = KO.make_scene(seed=0, net_res=16, grid_res=128, blob_radius=1.0)
spp = 32
kr = kilo.KiloRenderer(sc["grid"], sc["params"], sc["domain_mins"], sc["domain_maxs"], sc["gmin"], sc["gmax"], 4.0 / 384, 384, 2.0,
                       max_samples_per_ray=spp, device=dev)
b = O.lego_batch(800, 800)
b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}
for _ in range(3):
    out = kr.render(b)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = kr.render(b); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("frame ms", np.median(ts), "samples", int(kr.stats[0]))
np.save(sys.argv[1], out["rgb_map"].cpu().numpy())
