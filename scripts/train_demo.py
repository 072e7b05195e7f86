"""Convergence demo of the training path: fit a student NeRF to views rendered from a teacher
field with RayBatchTrainer (4096 rays per step) and log PSNR on a held-out view.  Writes a small text log."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fixtures as FX
from nerf_rep_for_test_b200 import Network, RenderConfig, Renderer, extras as X
dev = torch.device("cuda:0")
HW, n_views, steps = 160, 12, int(sys.argv[1]) if len(sys.argv) > 1 else 800
def net(sd):
    n = Network(device=dev); n.load_state_dict(sd); return n.to(dev)
teacher = Renderer(net(FX.make_state_dict(11, 300.0, 19.4)).eval(), RenderConfig(perturb=0, enable_ess=False, enable_ert=False), mode="bf16")
b = FX.lego_batch(HW, HW)
K = b["intrinsics"][0]
poses = []
for i in range(n_views + 1):
    th = 2 * math.pi * i / (n_views + 1)
    rot = torch.tensor([[math.cos(th), -math.sin(th), 0, 0], [math.sin(th), math.cos(th), 0, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]])
    poses.append(rot @ b["pose"][0])
images = torch.stack([teacher.render({"pose": p[None].to(dev), "intrinsics": K[None].to(dev), "H": HW, "W": HW})["rgb_map"] for p in poses])
student = net(FX.make_state_dict(12, 1.0, 0.15))
r = Renderer(student, RenderConfig(perturb=1, enable_ess=False, enable_ert=False), mode="bf16")
tr = X.RayBatchTrainer(r, images, poses, K, n_rays=4096, precrop_iters=0, seed=1)    # the last view is only evaluated, never drawn
tr.V = n_views
lines = ["step  loss      psnr(held-out view)  it/s"]
torch.cuda.synchronize(); t0 = time.perf_counter()
for s in range(0, steps, 100):
    losses = tr.fit(100)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    lines.append("%4d  %.6f  %6.2f dB           %.1f" % (s + 100, float(sum(losses[-20:]) / 20), float(tr.evaluate(n_views)), (s + 100) / dt))
    print(lines[-1], flush=True)
open(os.path.join(ROOT, "gpurun_out", "train_demo.txt"), "w").write("\n".join(lines) + "\n")
