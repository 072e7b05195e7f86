"""Generate tests/golden/*.npz from the UNMODIFIED reference (test infrastructure).

Run in the build container, where /root/reference is mounted:

    python oracle/gen_golden.py

For every case it (1) runs the real reference Renderer.render(batch) on CPU
with intermediates captured by wrapping its bound methods (no source edit),
(2) runs the in-repo restatement oracle/nerf_oracle.py on the same inputs and
asserts it is BIT-IDENTICAL on every output and intermediate, (3) writes the
reference's tensors as a fixture.  The GPU box has no /root/reference; there
the tests compare against these files and against the restatement.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import nerf_oracle as O          # noqa: E402
from oracle.ref_loader import build_reference  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")

CASES = [
    # name, H, W, seed, sigma_gain, sigma_bias, ert
    ("lego16_randinit", 16, 16, 0, 1.0, 0.0, False),
    ("lego16_randinit_ert", 16, 16, 0, 1.0, 0.0, True),
    ("lego8_dense", 8, 8, 1, 60.0, 0.5, False),
    ("lego8_dense_ert", 8, 8, 1, 60.0, 0.5, True),
    # BASELINE.json configs[0]: 32x32 = 1024 rays, lego test pose 0, seed 0 -- random init as is, and with a dense
    # field (the weights of bench.py's parity block).  "slim" fixtures: maps, bin indices, coarse weights, cdf and
    # the LAST sigma_raw of each pass (the flip-ray criterion); no per-sample raw tensors.
    ("lego32_cfg1", 32, 32, 0, 1.0, 0.0, False),
    ("lego32_dense", 32, 32, 0, 30.0, 0.2, False),
]
SLIM = {"lego32_cfg1", "lego32_dense"}


def capture_reference(r, batch):
    """Run r.render(batch) capturing intermediates of every 2048-ray chunk."""
    cap = {k: [] for k in ("z_coarse", "raw", "weights", "t_fine", "cdf", "inds")}
    orig_q, orig_r2o, orig_r2o_ert, orig_sf = r._query_network, r._raw2outputs, r._raw2outputs_with_ert, r._sample_fine
    orig_ss = torch.searchsorted

    def q(pts, vd, model):
        out = orig_q(pts, vd, model)
        cap["raw"].append(out.detach().clone())
        return out

    def wrap_r2o(fn):
        def f(raw, z, d):
            res = fn(raw, z, d)
            cap["weights"].append(res[3].detach().clone())
            cap["z_coarse"].append(z.detach().clone())
            return res
        return f

    def sf(t_mid, w):
        out = orig_sf(t_mid, w)
        cap["t_fine"].append(out.detach().clone())
        return out

    def ss(cdf, u, right=False, **kw):
        out = orig_ss(cdf, u, right=right, **kw)
        cap["cdf"].append(cdf.detach().clone())
        cap["inds"].append(out.detach().clone())
        return out

    r._query_network, r._raw2outputs, r._raw2outputs_with_ert, r._sample_fine = q, wrap_r2o(orig_r2o), wrap_r2o(orig_r2o_ert), sf
    torch.searchsorted = ss
    try:
        with torch.no_grad():
            out = r.render(batch)
    finally:
        torch.searchsorted = orig_ss
        r._query_network, r._raw2outputs, r._raw2outputs_with_ert, r._sample_fine = orig_q, orig_r2o, orig_r2o_ert, orig_sf
    aux = {
        "z_coarse": torch.cat(cap["z_coarse"][0::2]), "z_all": torch.cat(cap["z_coarse"][1::2]),
        "raw_coarse": torch.cat(cap["raw"][0::2]), "raw_fine": torch.cat(cap["raw"][1::2]),
        "weights_coarse": torch.cat(cap["weights"][0::2]), "weights_fine": torch.cat(cap["weights"][1::2]),
        "z_fine_samples": torch.cat(cap["t_fine"]), "cdf": torch.cat(cap["cdf"]), "inds": torch.cat(cap["inds"]),
    }
    return out, aux


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(8)
    for name, H, W, seed, gain, bias, ert in CASES:
        sd = O.make_state_dict(seed, gain, bias)
        _, net, r = build_reference(sd, enable_ess=False, enable_ert=ert)
        batch = O.lego_batch(H, W)
        ref_out, ref_aux = capture_reference(r, batch)
        with torch.no_grad():
            ora_out, ora_aux = O.render(sd, batch, use_ert=ert, ref_compat=True, return_aux=True)
        for k in ref_out:
            a, b = ref_out[k], ora_out[k]
            assert a.shape == b.shape, (name, k, a.shape, b.shape)
            assert torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0)), \
                "restatement != reference on %s/%s (max diff %g)" % (name, k, (a - b).abs().max())
        for k in ref_aux:
            assert torch.equal(ref_aux[k], ora_aux[k]), "restatement != reference on aux %s/%s" % (name, k)
        rays_o, rays_d = O.get_rays(H, W, batch["pose"][0], batch["intrinsics"][0])
        arrays = {"rays_o": rays_o.numpy(), "rays_d": rays_d.numpy(),
                  "pose": batch["pose"].numpy(), "intrinsics": batch["intrinsics"].numpy(),
                  "meta": np.array([H, W, seed, gain, bias, int(ert)], dtype=np.float64)}
        for k, v in ref_out.items():
            arrays["out_" + k] = v.numpy()
        for k, v in ref_aux.items():
            if name in SLIM:
                if k in ("raw_coarse", "raw_fine"):
                    arrays["aux_sigma_last_" + k[4:]] = v[:, -1, 3].numpy()
                    continue
                if k not in ("inds", "cdf", "weights_coarse"):
                    continue
            arrays["aux_" + k] = v.numpy().astype(np.int16) if k == "inds" else v.numpy()
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **arrays)
        acc = ref_out["acc_map"]
        print("%-22s ok  rays=%d  acc_map mean %.4f  rgb mean %.4f  file %.0f KB" % (
            name, H * W, float(acc.mean()), float(ref_out["rgb_map"].mean()), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
