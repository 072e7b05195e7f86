"""Tiny end-to-end exercise of every kernel (for compute-sanitizer memcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import nerf_oracle as O
from nerf_rep_for_test_b200 import Network, RenderConfig, Renderer, lib as L, ops, training as T
dev = torch.device("cuda:0")
sd = O.make_state_dict(0, 30.0, 0.3)
net = Network(device=dev); net.load_state_dict(sd); net.to(dev).eval()
b = O.lego_batch(9, 7)
gb = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}
for mode in ("fp32", "bf16"):
    for ess, ert in ((False, False), (True, True)):
        r = Renderer(net, RenderConfig(perturb=0, enable_ess=ess, enable_ert=ert), mode=mode)
        out = r.render(gb)
        r.render_host(b)
r = Renderer(net, RenderConfig(perturb=0, enable_ess=True, enable_ert=True), mode="bf16")
r.ess_mode = "skip"
r.render(gb)
r2 = Renderer(net, RenderConfig(perturb=1, enable_ess=False, enable_ert=False), mode="bf16")
net.train()
ro, rd = ops.raygen(gb["pose"], gb["intrinsics"], 9, 7)
step = T.TrainStep(r2)
step(ro, rd, torch.rand(63, 3, device=dev))
raw, dump = ops.mlp_forward_stages(r2.packed("coarse"), ro, rd, ops.sample_coarse(r2._table("z"), 63))
torch.cuda.synchronize()
print("sanitize smoke ok")
