"""CPU emulation of the BF16-mode MLP arithmetic (test helper): operands rounded to bf16, fp32
accumulation, biases/activations in fp32, alpha/rgb heads on the unrounded fp32 activations --
what nerf_rep_for_test_b200/csrc/mlp_bf16_tc.cu computes, built from the oracle's pieces."""
import torch
import torch.nn.functional as F

from oracle import nerf_oracle as O


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def mlp_bf16_stages(sd, prefix, pts, dirs):
    """pts, dirs [M,3] -> (raw [M,4], list of 10 fp32 stage outputs)."""
    pe, dpe = O.pos_enc(pts, O.L_XYZ), O.pos_enc(dirs, O.L_DIR)
    W = lambda n: bf(sd[prefix + n + ".weight"])
    B = lambda n: sd[prefix + n + ".bias"]
    stages = []
    h = bf(pe)
    x = None
    for i in range(8):
        inp = torch.cat([bf(pe), h], -1) if i == 5 else h
        x = F.relu(F.linear(inp, W("pts_linears.%d" % i), B("pts_linears.%d" % i)))
        stages.append(x)
        h = bf(x)
    sigma = F.linear(x, sd[prefix + "alpha_linear.weight"], sd[prefix + "alpha_linear.bias"])
    feat = F.linear(h, W("feature_linear"), B("feature_linear"))
    stages.append(feat)
    hv = F.relu(F.linear(torch.cat([bf(feat), bf(dpe)], -1), W("views_linears.0"), B("views_linears.0")))
    stages.append(hv)
    rgb = F.linear(hv, sd[prefix + "rgb_linear.weight"], sd[prefix + "rgb_linear.bias"])
    return torch.cat([rgb, sigma], -1), stages
