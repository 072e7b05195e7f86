"""Quick device-time probe of the whole render path and of the MLP kernel alone."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import nerf_oracle as O
from nerf_rep_for_test_b200 import Network, RenderConfig, Renderer, lib as L, ops

dev = torch.device("cuda:0")
mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
HW = int(sys.argv[2]) if len(sys.argv) > 2 else 800
sd = O.make_state_dict(0)
net = Network(device=dev); net.load_state_dict(sd); net.to(dev).eval()
r = Renderer(net, RenderConfig(perturb=0, enable_ess=False, enable_ert=False), mode=mode)
b = O.lego_batch(HW, HW)
gb = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}
for _ in range(2):
    out = r.render(gb)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
e0.record()
for _ in range(reps):
    out = r.render(gb)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
rays = HW * HW
print("%s render %dx%d: %.2f ms/frame  %.3f Mrays/s  %.1f TFLOP/s algorithmic" % (mode, HW, HW, ms, rays / ms / 1e3, rays * 303824896 / ms / 1e9))
# MLP kernel alone on 8192 rays x 192
n, S = 8192, 192
ro, rd = ops.raygen(gb["pose"], gb["intrinsics"], HW, HW)
z = ops.sample_coarse(torch.linspace(2, 6, S, device=dev), n)
packed = ops.pack_from_state_dict(sd, "model_fine.", r.MODES[mode], dev)
for _ in range(2):
    raw = ops.mlp_forward(packed, ro[:n], rd[:n], z)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    raw = ops.mlp_forward(packed, ro[:n], rd[:n], z)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("%s mlp_forward %d rows: %.3f ms  %.1f TFLOP/s algorithmic (1186816 FLOP/row)" % (mode, n * S, ms, n * S * 1186816 / ms / 1e9))
