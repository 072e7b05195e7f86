"""Frame time of every arithmetic mode on one 800x800 lego view (random-init weights), CUDA events, L2 flushed.
Usage: python scripts/mode_probe.py [modes...]   (default: bf16 mixed fp32tc)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fixtures as FX
from nerf_rep_for_test_b200 import Network, RenderConfig, Renderer, lib as L

dev = torch.device("cuda:0")
modes = sys.argv[1:] or ["bf16", "fp16", "mixed", "mixed16", "fp32tc"]
sd = FX.make_state_dict(0)
net = Network(device=dev)
net.load_state_dict(sd)
net.to(dev).eval()
H = W = 800
b = FX.lego_batch(H, W)
gb = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {}
for mode in modes:
    r = Renderer(net, RenderConfig(perturb=0, enable_ess=False, enable_ert=False), mode=mode)
    r.render(gb)
    torch.cuda.synchronize()
    times = []
    L.profile_enable(True)
    for _ in range(3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r.render(gb)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    mlp_ms, n, rows = L.profile_read()
    L.profile_enable(False)
    res[mode] = {"ms_per_frame": sorted(times)[1], "all_ms": times, "mlp_ms_per_frame": mlp_ms / 3,
                 "mlp_tflops_algorithmic": rows * 1186816 / (mlp_ms * 1e-3) / 1e12}
    print(mode, json.dumps(res[mode]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/mode_probe.json", "w"), indent=1)
