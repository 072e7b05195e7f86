"""Where does the end-to-end (host-buffer) frame spend its time beyond the device render?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fixtures as FX
from nerf_rep_for_test_b200 import Network, RenderConfig, Renderer
dev = torch.device("cuda:0")
net = Network(device=dev); net.load_state_dict(FX.make_state_dict(0)); net.to(dev).eval()
r = Renderer(net, RenderConfig(perturb=0, enable_ess=False, enable_ert=False), mode="bf16")
b = FX.lego_batch(800, 800)
gb = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}
for _ in range(2):
    r.render(gb); r.render_host(b)
torch.cuda.synchronize()
def wall(fn, n=5):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return sorted(ts)[len(ts) // 2], ts
def dev_render():
    r.render(gb)
print("render (device maps), wall per frame incl. sync: %.2f ms %s" % wall(dev_render))
print("render_host (pinned host maps), wall per frame:  %.2f ms %s" % wall(lambda: r.render_host(b)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(5):
    torch.cuda.synchronize(); e0.record(); r.render(gb); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print("render, CUDA events: %.2f ms" % sorted(ts)[2], ts)
ts = []
for _ in range(5):
    torch.cuda.synchronize(); e0.record(); r.render_host(b); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print("render_host, CUDA events on the render stream: %.2f ms" % sorted(ts)[2], ts)
# back-to-back device renders without a sync in between (what bench.py's `value` loop does)
torch.cuda.synchronize(); e0.record()
for _ in range(5):
    r.render(gb)
e1.record(); torch.cuda.synchronize()
print("5 renders back to back, CUDA events: %.2f ms per frame" % (e0.elapsed_time(e1) / 5))
