"""Achieved HBM bandwidth of the streaming kernels (ray generation, coarse sampling, compositing, sample_pdf + merge)
at full-frame size (640 000 rays, HBM-resident operands), against MEASURED_PEAKS.json's copy bandwidth.
Algorithmic bytes per ray as in DESIGN.md section 4.3."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fixtures as FX
from nerf_rep_for_test_b200 import lib as L, ops
dev = torch.device("cuda:0")
peak = 6545.9
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]
N, S, U = 640000, 64, 128
b = FX.lego_batch(800, 800)
pose, K = b["pose"][0].to(dev), b["intrinsics"][0].to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]

g = torch.Generator(device=dev).manual_seed(0)
ro, rd = ops.raygen(pose, K, 800, 800)
ztab = torch.linspace(2, 6, S, device=dev)
z_c = ops.sample_coarse(ztab, N)
raw_c = torch.randn(N, S, 4, device=dev, generator=g) * 0.3
u = torch.linspace(0, 1, U, device=dev)
rows = []
def rec(name, ms, bytes_):
    rows.append({"kernel": name, "ms": ms, "algorithmic_GB": bytes_ / 1e9, "GBps": bytes_ / ms / 1e6, "frac_of_hbm_peak": bytes_ / ms / 1e6 / peak})
rec("raygen_kernel (640k rays)", timeit(lambda: ops.raygen(pose, K, 800, 800)), N * 24)
rec("sample_coarse_kernel (perturb=0)", timeit(lambda: ops.sample_coarse(ztab, N)), N * S * 4)
rec("sample_coarse_kernel (stratified jitter)", timeit(lambda: ops.sample_coarse(ztab, N, perturb=True, seed=1)), N * S * 4)
for variant, tag in ((L.COMPOSITE_PLAIN, "plain, exact"), (L.COMPOSITE_ERT, "ert, exact"), (L.COMPOSITE_PLAIN | L.COMPOSITE_FAST_MATH, "plain, fast math")):
    rec("composite_kernel<%s> coarse S=64" % tag, timeit(lambda: ops.composite_forward(raw_c, z_c, rd, variant)), N * (S * 16 + S * 4 + S * 4 + 12 + 24))
w_c = ops.composite_forward(raw_c, z_c, rd)[3]
rec("sample_pdf_merge_kernel 64 -> 192", timeit(lambda: ops.sample_pdf_merge(z_c, w_c, u, want_aux=False)), N * (S * 4 + S * 4 + (S + U) * 4))
z_all = ops.sample_pdf_merge(z_c, w_c, u, want_aux=False)[0]
raw_f = torch.randn(N, S + U, 4, device=dev, generator=g) * 0.3
rec("composite_kernel<plain, exact> fine S=192 (no weights out)", timeit(lambda: ops.composite_forward(raw_f, z_all, rd, want_weights=False)), N * ((S + U) * 20 + 12 + 24))
rec("composite_kernel<plain, fast math> fine S=192 (no weights out)", timeit(lambda: ops.composite_forward(raw_f, z_all, rd, L.COMPOSITE_PLAIN | L.COMPOSITE_FAST_MATH, want_weights=False)), N * ((S + U) * 20 + 12 + 24))
g_rgb = torch.randn(N, 3, device=dev, generator=g)
rec("composite_backward_kernel S=192", timeit(lambda: ops.composite_backward(raw_f, z_all, rd, g_rgb)), N * ((S + U) * (16 + 4 + 16) + 12 + 12))
out = {"hbm_peak_GBps": peak, "peak_source": "MEASURED_PEAKS.json (copy, read+write)", "rays": N, "timing": "CUDA events, median of 10, L2 flushed (256 MB write) before every launch; output allocation included", "kernels": rows}
print(json.dumps(out, indent=1))
