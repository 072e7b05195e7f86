"""In-kernel clock64 timeline of the TRAINING forward (NERFB200_TIMELINE): per (quad, stage, slot)
issuer wait start / MMA start / epilogue start / epilogue end."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import nerf_oracle as O
from nerf_rep_for_test_b200 import lib as L, ops
dev = torch.device("cuda:0")
sd = O.make_state_dict(0)
n, S = 4096, 192
b = O.lego_batch(800, 800)
ro, rd = ops.raygen(b["pose"].to(dev), b["intrinsics"].to(dev), 800, 800)
z = ops.sample_coarse(torch.linspace(2, 6, S, device=dev), n)
packed = ops.pack_from_state_dict(sd, "model_fine.", L.MODE_BF16, dev)
os.environ.pop("NERFB200_TIMELINE", None)
for _ in range(2):
    ops.mlp_forward_train(packed, ro[:n], rd[:n], z)
torch.cuda.synchronize()
os.environ["NERFB200_TIMELINE"] = sys.argv[1]
ops.mlp_forward_train(packed, ro[:n], rd[:n], z)
torch.cuda.synchronize()
d = np.loadtxt(sys.argv[1], dtype=np.int64)
d = d[d[:, 2] < 2]
t0 = d[:, 3:].min()
print("DBG", os.environ.get("NERFB200_DBG"))
print("it stage slot | wait_start mma_start epi_start epi_end | mma_pass epi handoff_to_next_mma")
rows = {(int(r[0]), int(r[1]), int(r[2])): r[3:] - t0 for r in d if r[3] > 0}
for it in (1,):
    for st in range(10):
        for sl in (0, 1):
            r = rows.get((it, st, sl))
            if r is None:
                continue
            nxt = rows.get((it, st + 1, sl)) if st < 9 else rows.get((it + 1, 0, sl))
            hand = (nxt[1] - r[3]) if nxt is not None else -1
            print("%d %d %d | %7d %7d %7d %7d | %5d %5d %5d" % (it, st, sl, r[0], r[1], r[2], r[3], r[2] - r[1], r[3] - r[2], hand))
