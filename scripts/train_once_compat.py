"""Training steps for the launch list / ncu: bf16 detached (default), then bf16 with the reference graph, then fp32."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fixtures as FX
from nerf_rep_for_test_b200 import Network, RenderConfig, Renderer, ops, training as T
dev = torch.device("cuda:0")
variants = sys.argv[1:] or ["bf16", "bf16+compat", "fp32+compat"]
for v in variants:
    net = Network(device=dev); net.load_state_dict(FX.make_state_dict(0)); net.to(dev).train()
    r = Renderer(net, RenderConfig(perturb=1, enable_ess=False, enable_ert=False), mode="bf16")
    step = T.TrainStep(r, precision=v.split("+")[0], ref_compat_sampler=v.endswith("compat"))
    b = FX.lego_batch(800, 800)
    ro, rd = ops.raygen(b["pose"][0].to(dev), b["intrinsics"][0].to(dev), 800, 800)
    sel = torch.randint(0, 640000, (4096,), device=dev)
    ro, rd = ro[sel].contiguous(), rd[sel].contiguous()
    tgt = torch.rand(4096, 3, device=dev)
    for _ in range(3):
        step(ro, rd, tgt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10 if v.startswith("bf16") else 3
    e0.record()
    for _ in range(n):
        step(ro, rd, tgt)
    e1.record()
    torch.cuda.synchronize()
    print("train %-12s %.3f ms per step (%d steps)" % (v, e0.elapsed_time(e1) / n, n), flush=True)
