"""HBM write / read / copy bandwidth with plain torch ops (what bounds the activation store)."""
import torch
dev = torch.device("cuda:0")
n = 1 << 30
a = torch.empty(n, dtype=torch.int32, device=dev)   # 4 GiB
b = torch.empty(n, dtype=torch.int32, device=dev)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: a.zero_());  print("write  %.0f GB/s" % (4 * n / ms / 1e6))
ms = t(lambda: a.sum());    print("read   %.0f GB/s" % (4 * n / ms / 1e6))
ms = t(lambda: b.copy_(a)); print("copy   %.0f GB/s (read+write)" % (8 * n / ms / 1e6))
