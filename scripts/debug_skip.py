import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import nerf_oracle as O
from nerf_rep_for_test_b200 import Network, RenderConfig, Renderer, lib as L, ops
DEV = torch.device("cuda:0")
sd = O.make_state_dict(6, 40.0, 0.5)
res = 128
gc = torch.stack(torch.meshgrid([torch.arange(res)] * 3, indexing="ij"), -1).float() / (res - 1) * 2 - 1
grid = torch.norm(gc, dim=-1) <= 0.35
b = O.lego_batch(40, 40)
bc = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in b.items()}
net = Network(device=DEV); net.load_state_dict(sd); net.to(DEV).eval()
skip = Renderer(net, RenderConfig(perturb=0, enable_ess=True, enable_ert=True), mode="bf16")
skip.occupancy_grid = grid.to(DEV); skip.ess_mode = "skip"
out = skip.render(bc)
print("counts", skip.eval_counts.tolist())
for k in ("acc_map_0", "acc_map", "rgb_map_0", "rgb_map"):
    print(k, float(out[k].mean()), float(out[k].max()))
# manual pipeline
ro, rd = ops.raygen(bc["pose"], bc["intrinsics"], 40, 40)
z = ops.sample_coarse(skip._table("z"), 1600)
g8 = grid.to(DEV).to(torch.uint8)
ids, na = ops.ess_compact(g8, ro, rd, z)
raw = ops.mlp_forward_sparse(skip.packed("coarse"), ro, rd, z, ids, na)
rgb, disp, acc, w, depth = ops.composite_forward(raw, z, rd, L.COMPOSITE_ERT, 0.01)
print("manual coarse: active", int(na), "acc mean", float(acc.mean()), "raw sigma max", float(raw[..., 3].max()))
zt = ops.ert_depth(w, z, 0.01)
print("z_term finite", int(torch.isfinite(zt).sum()), "min", float(zt.min()))
z_all = ops.sample_pdf_merge(z, w, skip._table("u"), want_aux=False)[0]
ids2, na2 = ops.ess_compact(g8, ro, rd, z_all, zt)
raw_f = ops.mlp_forward_sparse(skip.packed("fine"), ro, rd, z_all, ids2, na2)
print("manual fine: active", int(na2), "sigma max", float(raw_f[..., 3].max()), "acc", float(ops.composite_forward(raw_f, z_all, rd, L.COMPOSITE_ERT, 0.01)[2].mean()))
