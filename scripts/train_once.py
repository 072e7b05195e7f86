"""Run a few 4096-ray training steps (for ncu captures of the training kernels)."""
import os, sys
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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(n): step(ro, rd, tgt)
torch.cuda.synchronize()
print("ok")
