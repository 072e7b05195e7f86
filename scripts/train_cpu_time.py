"""Host-side cost of a training step: enqueue N steps without synchronising and compare with the device time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import nerf_oracle as O
from nerf_rep_for_test_b200 import Network, RenderConfig, Renderer, training as T
dev = torch.device("cuda:0")
net = Network(device=dev); net.load_state_dict(O.make_state_dict(0)); net.to(dev).train()
r = Renderer(net, RenderConfig(perturb=1, enable_ess=False, enable_ert=False), mode="bf16")
step = T.TrainStep(r)
ro, rd = O.get_rays(800, 800, torch.tensor(O.LEGO_TEST_POSE0), O.lego_batch(800, 800)["intrinsics"][0])
sel = torch.randint(0, 640000, (4096,))
ro, rd = ro[sel].to(dev), rd[sel].to(dev)
tgt = torch.rand(4096, 3, device=dev)
for _ in range(5): step(ro, rd, tgt)
torch.cuda.synchronize()
n = 40
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(n): step(ro, rd, tgt)
e1.record(); t1 = time.perf_counter()
torch.cuda.synchronize(); t2 = time.perf_counter()
print("host enqueue %.3f ms/step, device %.3f ms/step, wall %.3f ms/step" % ((t1 - t0) / n * 1e3, e0.elapsed_time(e1) / n, (t2 - t0) / n * 1e3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(n): step(ro, rd, tgt)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
