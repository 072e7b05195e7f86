"""CPU: the built objects contain the instructions the design claims (B200_PROFILING.md's SASS mnemonics): tcgen05.mma /
commit / ld and bulk copies in the dense MLP kernels, mma.sync in the two warp-level tensor-core kernels.  Guards against a
kernel silently degrading to a CUDA-core formulation while its parity tests stay green."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "nerf_rep_for_test_b200", "build")

EXPECT = {
    "mlp_bf16_tc2.o": {"mlp_bf16_tc2_kernel": ["UTCHMMA", "UTCBAR", "LDTM", "UBLKCP"]},
    "mlp_f16x2_tc2.o": {"mlp_f16x2_tc2_kernel": ["UTCHMMA", "UTCBAR", "LDTM", "UBLKCP"]},
    "mlp_bwd_dgrad.o": {"mlp_bwd_dgrad_kernel": ["UTCHMMA", "UTCBAR", "LDTM", "UBLKCP"]},
    "mlp_bwd_wgrad.o": {"mlp_bwd_wgrad_kernel": ["UTCHMMA", "UTCBAR", "LDTM", "UBLKCP"]},
    "mlp_bwd_input.o": {"mlp_bwd_input_kernel": ["HMMA.16816.F32.BF16", "LDSM"]},
    "kilo.o": {"eval_tc_kernel": ["HMMA.16816.F32"]},
}


@pytest.mark.parametrize("obj", sorted(EXPECT))
def test_tensor_core_mnemonics_present(built_lib, obj):
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    path = os.path.join(OBJ, obj)
    assert os.path.exists(path), "object not built: " + path
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    # split per function
    funcs = {}
    name = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name is not None:
            funcs[name].append(line)
    for kernel, mnemonics in EXPECT[obj].items():
        bodies = ["\n".join(v) for k, v in funcs.items() if kernel in k]
        assert bodies, (obj, kernel, list(funcs))
        for body in bodies:                       # every instantiation of the kernel
            for mn in mnemonics:
                assert mn in body, "%s: no %s in %s" % (obj, mn, kernel)
    if obj == "kilo.o":                            # the micro-MLP is 312 warp-level MMAs per 32 samples
        body = [("\n".join(v)) for k, v in funcs.items() if "eval_tc_kernel" in k][0]
        assert body.count("HMMA.16816.F32") == 312
