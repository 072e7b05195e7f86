"""Frame time in the reference's DEFAULT configuration (lego.yaml: enable_ess=True, enable_ert=True -> occupancy-grid
resampling + the literal ERT compositor with its 2048-ray chunk quirk) next to the benchmark configuration (both off)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fixtures as FX
from nerf_rep_for_test_b200 import Network, RenderConfig, Renderer
dev = torch.device("cuda:0")
net = Network(device=dev); net.load_state_dict(FX.make_state_dict(0)); net.to(dev).eval()
b = FX.lego_batch(800, 800)
gb = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, cfg, kw in (("benchmark config (ESS/ERT off)", dict(enable_ess=False, enable_ert=False), {}),
                      ("lego.yaml default (ESS resample + ERT, reference-compatible quirk)", dict(enable_ess=True, enable_ert=True), {}),
                      ("ESS resample + ERT, intended per-ray truncation (ref_compat=False)", dict(enable_ess=True, enable_ert=True), dict(ref_compat=False))):
    r = Renderer(net, RenderConfig(perturb=0, **cfg), mode="bf16", **kw)
    for _ in range(2):
        r.render(gb)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        r.render(gb)
    e1.record(); torch.cuda.synchronize()
    print("%-75s %.2f ms/frame" % (name, e0.elapsed_time(e1) / 3))
