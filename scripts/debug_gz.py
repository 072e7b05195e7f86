"""Diagnostic: dL/dz through the fine MLP input, bf16 path vs fp32 path at identical sample positions."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nerf_oracle as O
from nerf_rep_for_test_b200 import Network, RenderConfig, Renderer, ops, training as T, lib as L
DEV = torch.device("cuda:0")
sd = O.make_state_dict(3, 30.0, 0.2)
net = Network(device=DEV); net.load_state_dict(sd); net.to(DEV).eval()
r = Renderer(net, RenderConfig(perturb=0, enable_ess=False, enable_ert=False), mode="bf16")
b = O.lego_batch(32, 32)
ro, rd = O.get_rays(32, 32, b["pose"][0], b["intrinsics"][0])
sel = torch.randperm(ro.shape[0], generator=torch.Generator().manual_seed(0))[:192]
ro, rd = ro[sel].contiguous().to(DEV), rd[sel].contiguous().to(DEV)
tgt = torch.rand(192, 3, generator=torch.Generator().manual_seed(1)).to(DEV)
cos = lambda a, b: float((a * b).sum() / (a.norm() * b.norm() + 1e-30))
rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-30))
st16 = T._forward_passes(r, ro, rd, "bf16")
st32 = T._forward_passes(r, ro, rd, "fp32", _inject={"w_c": st16["w_c"], "z_all": st16["z_all"]})
out = {}
for name, st in (("bf16", st16), ("fp32", st32)):
    rgb = st["maps"][4]
    g_rgb = (rgb - tgt) * (2.0 / rgb.numel())
    g_raw_f, g_z_comp = ops.composite_backward_z(st["raw_f"], st["z_all"], rd, g_rgb, None, None, None)
    _, g_z_mlp = T._mlp_backward(r, "fine", name, g_raw_f, st["store_f"], ro, rd, st["z_all"], None, want_g_z=True)
    g_w = ops.sample_pdf_backward(st["z_c"], st["w_c"], st["u"], g_z_comp + g_z_mlp)
    out[name] = dict(g_raw_f=g_raw_f, g_z_comp=g_z_comp, g_z_mlp=g_z_mlp, g_w=g_w, raw_f=st["raw_f"])
for k in ("raw_f", "g_raw_f", "g_z_comp", "g_z_mlp", "g_w"):
    a, b_ = out["bf16"][k], out["fp32"][k]
    print("%-9s |fp32| %.3e  cos %.5f rel %.4f" % (k, float(b_.norm()), cos(a, b_), rel(a, b_)))
# same g_raw into both MLP backward paths: isolates the MLP input gradient
_, gz16 = T._mlp_backward(r, "fine", "bf16", out["fp32"]["g_raw_f"], st16["store_f"], ro, rd, st16["z_all"], None, want_g_z=True)
print("g_z_mlp with the SAME g_raw: cos %.5f rel %.4f" % (cos(gz16, out["fp32"]["g_z_mlp"]), rel(gz16, out["fp32"]["g_z_mlp"])))
gz = out["fp32"]["g_z_mlp"]
print("per-ray cos of g_z_mlp:", [round(cos(gz16[i], gz[i]), 3) for i in range(0, 192, 16)])
print("|g_z_mlp| by sample position (fp32):", [float(gz[:, i].abs().mean()) for i in range(0, 192, 24)])
