"""Bring-up diagnostic for the tcgen05 MLP kernel: per-stage comparison against the bf16 emulation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import nerf_oracle as O
from nerf_rep_for_test_b200 import lib as L, ops
from bf16_emul import mlp_bf16_stages

dev = torch.device("cuda:0")
sd = O.make_state_dict(0)
n, S = int(sys.argv[1]) if len(sys.argv) > 1 else 4, 64
b = O.lego_batch(16, 16)
ro, rd = O.get_rays(16, 16, b["pose"][0], b["intrinsics"][0])
ro, rd = ro[:n].contiguous(), rd[:n].contiguous()
z = O.sample_coarse(n, S)
packed = ops.pack_from_state_dict(sd, "model.", L.MODE_BF16, dev)
raw, dump = ops.mlp_forward_stages(packed, ro.to(dev), rd.to(dev), z.to(dev))
torch.cuda.synchronize()
pts = (ro[:, None, :] + rd[:, None, :] * z[..., None]).reshape(-1, 3)
dirs = rd[:, None, :].expand(n, S, 3).reshape(-1, 3)
ref_raw, stages = mlp_bf16_stages(sd, "model.", pts, dirs)
rows = min(128, n * S)
for i, st in enumerate(stages):
    got = dump[i, :rows, : st.shape[1]].cpu()
    err = (got - st[:rows]).abs()
    print("stage %d: max abs err %.3e  (ref absmax %.3f)  worst row %d col %d  nonzero got %d/%d" % (
        i, float(err.max()), float(st[:rows].abs().max()), int(err.max(1)[0].argmax()), int(err.max(0)[0].argmax()),
        int((got != 0).sum()), got.numel()))
    if float(err.max()) > 5e-2 and i < 3:
        print("  got[0,:8]", got[0, :8].tolist()); print("  ref[0,:8]", st[0, :8].tolist())
        print("  got[1,:8]", got[1, :8].tolist()); print("  ref[1,:8]", st[1, :8].tolist())
err = (raw.cpu().reshape(-1, 4) - ref_raw).abs()
print("raw: max abs err vs bf16 emulation %.3e ; vs fp32 oracle %.3e" % (
    float(err.max()), float((raw.cpu().reshape(-1, 4) - O.nerf_mlp(sd, "model.", torch.cat([O.pos_enc(pts, 10), O.pos_enc(dirs, 4)], -1))).abs().max())))
