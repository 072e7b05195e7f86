"""How the CPU oracle's render time depends on the OpenMP environment torchrun imposes (OMP_NUM_THREADS=1)."""
import os, sys, time
mode = sys.argv[1]
if mode == "fix":
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
print(mode, "env OMP", os.environ.get("OMP_NUM_THREADS"), "cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)),
      "torch default threads", torch.get_num_threads(), flush=True)
torch.set_num_threads(os.cpu_count() or 1)
import fixtures as FX
from oracle import nerf_oracle as O
sd = FX.make_state_dict(0)
b = FX.lego_batch(32, 32)
ro, rd = O.get_rays(32, 32, b["pose"][0], b["intrinsics"][0])
for tag, f in (("fp32", lambda: O.render_rays(sd, ro, rd)),
               ("fp64", lambda: O.render_rays({k: v.double() for k, v in sd.items()}, ro.double(), rd.double()))):
    with torch.no_grad():
        f()
        t = time.perf_counter(); f(); dt = time.perf_counter() - t
    print("   ", tag, "render of 1024 rays: %.2f s" % dt, flush=True)
