"""GPU probe: which device exp matches torch-CPU exp (the reference's arithmetic) most often."""
import torch
g = torch.Generator().manual_seed(0)
for lo, hi in ((-1e-3, 0.0), (-1e-1, 0.0), (-5.0, 0.0), (-5.0, 5.0)):
    x = torch.rand(1 << 20, generator=g) * (hi - lo) + lo
    ref = torch.exp(x)
    xc = x.cuda()
    a = torch.exp(xc).cpu()                             # CUDA expf
    b = torch.exp(xc.double()).float().cpu()            # correctly rounded
    c = torch.exp(x.double()).float()                   # CPU double
    print("range [%g,%g]: cuda expf != cpu: %.4f%%   f64-rounded != cpu: %.4f%%  (cpu f64 vs cuda f64 %.4f%%)" % (
        lo, hi, 100 * (a != ref).float().mean(), 100 * (b != ref).float().mean(), 100 * (b != c).float().mean()))
    s = torch.sigmoid(x); sc = torch.sigmoid(xc).cpu()
    s2 = (1.0 / (1.0 + torch.exp(-xc))).cpu()
    s3 = torch.sigmoid(xc.double()).float().cpu()
    print("   sigmoid: cuda != cpu %.4f%%  1/(1+exp(-x)) != cpu %.4f%%  f64 != cpu %.4f%%" % (
        100 * (sc != s).float().mean(), 100 * (s2 != s).float().mean(), 100 * (s3 != s).float().mean()))
