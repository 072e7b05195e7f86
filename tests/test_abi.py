"""CPU: the C-ABI library builds, loads and exports every symbol include/nerfb200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "nerfb200.h")).read()
    return sorted(set(re.findall(r"NERFB200_API[^;]*?\b(nerfb200_\w+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    names = _declared()
    for must in ("nerfb200_raygen", "nerfb200_sample_coarse", "nerfb200_mlp_forward", "nerfb200_composite_forward",
                 "nerfb200_composite_backward", "nerfb200_sample_pdf_merge", "nerfb200_sample_from_cdf",
                 "nerfb200_pack_weights", "nerfb200_ess_resample", "nerfb200_render_rays",
                 "nerfb200_render_image_host", "nerfb200_get_last_error_string"):
        assert must in names


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for name in _declared():
        assert hasattr(lib, name), "symbol %s declared in nerfb200.h but not exported" % name


def test_binding_covers_header(built_lib):
    from nerf_rep_for_test_b200 import lib as L
    assert sorted(L.SIGNATURES) == _declared()
    lib = L.load()
    assert lib.nerfb200_abi_version() == L.ABI_VERSION
    assert lib.nerfb200_packed_weights_bytes(L.MODE_FP32) > 4 * 590000
    assert lib.nerfb200_packed_weights_bytes(L.MODE_BF16) > 2 * 590000
    assert lib.nerfb200_packed_weights_bytes(7) == 0


def test_struct_sizes_match_c_layout():
    from nerf_rep_for_test_b200 import lib as L
    assert ctypes.sizeof(L.MlpWeights) == 24 * 8
    assert ctypes.sizeof(L.Maps) == 32
    assert ctypes.sizeof(L.RenderParams) == 104  # 8 ints, 2 floats, u64, ptr, 2 ints, ptr, int, 6 floats, pad (static_assert in api.cu)


def test_argument_errors_are_reported_not_fatal(built_lib):
    """Bad arguments return non-zero + message (never exit()); no GPU needed for these checks."""
    from nerf_rep_for_test_b200 import lib as L
    lib = L.load()
    rc = lib.nerfb200_raygen(None, None, 4, 4, None, None, None)
    assert rc != 0 and b"null" in lib.nerfb200_get_last_error_string()
    rc = lib.nerfb200_sample_pdf_merge(ctypes.c_void_p(8), ctypes.c_void_p(8), ctypes.c_void_p(8), 0, 1, 9999, 128,
                                       ctypes.c_void_p(8), None, None, None, None)
    assert rc != 0 and b"n_samples" in lib.nerfb200_get_last_error_string()
    with pytest.raises(L.NerfB200Error):
        L.check(rc, "sample_pdf_merge")


def test_renderer_fails_loudly_without_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nerf_rep_for_test_b200 import NerfB200Error, Network, Renderer
    with pytest.raises(NerfB200Error):
        Renderer(Network(device="cpu"))


def test_argument_errors_of_the_skipping_and_noise_entries(built_lib):
    """Entry points added for empty-space skipping (keep bits, ray culling) and raw_noise_std reject bad arguments
    before touching the device."""
    from nerf_rep_for_test_b200 import lib as L
    lib = L.load()
    p8 = ctypes.c_void_p(8)
    assert lib.nerfb200_sigma_noise(None, 5, 1.0, 0, None) != 0 and b"null" in lib.nerfb200_get_last_error_string()
    assert lib.nerfb200_sigma_noise(p8, 5, -1.0, 0, None) != 0 and b"std" in lib.nerfb200_get_last_error_string()
    assert lib.nerfb200_sigma_noise(p8, 0, 1.0, 0, None) == 0                      # empty batch: nothing to do
    box = (ctypes.c_float * 3)(0, 0, 0)
    assert lib.nerfb200_ray_cull(p8, p8, 4, p8, 64, box, box, None, None) != 0 and b"null" in lib.nerfb200_get_last_error_string()
    assert lib.nerfb200_ray_cull(p8, p8, 4, p8, 64, None, box, p8, None) != 0 and b"box" in lib.nerfb200_get_last_error_string()
    assert lib.nerfb200_ray_cull(p8, p8, 0, p8, 64, box, box, p8, None) == 0
    # masked compositor: the literal ERT_COMPAT variant has no masked form
    rc = lib.nerfb200_composite_forward_masked(p8, p8, p8, p8, 4, 64, L.COMPOSITE_ERT_COMPAT, 0.01, 1, 2048, p8, p8, p8, p8,
                                               None, None)
    assert rc != 0 and b"masked" in lib.nerfb200_get_last_error_string()
    # render params: noise cannot be combined with skipping
    prm = L.RenderParams()
    prm.n_samples, prm.n_importance, prm.mode, prm.compat_chunk = 64, 128, L.MODE_BF16, 2048
    prm.raw_noise_std, prm.ess_skip, prm.occupancy_grid, prm.grid_res = 1.0, 1, 8, 128
    maps = L.Maps(p8, p8, p8, p8)
    rc = lib.nerfb200_render_rays(p8, p8, p8, p8, 4, p8, p8, ctypes.byref(prm), p8, 1 << 30, ctypes.byref(maps),
                                  ctypes.byref(maps), None)
    assert rc != 0 and b"raw_noise_std" in lib.nerfb200_get_last_error_string()


def test_ctypes_structs_match_the_c_header_field_by_field(tmp_path):
    """Offsets of every field the Python binding mirrors, taken from the C compiler's view of include/nerfb200.h."""
    import shutil
    import subprocess
    from nerf_rep_for_test_b200 import lib as L
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    structs = {"nerfb200_render_params": L.RenderParams, "nerfb200_maps": L.Maps, "nerfb200_mlp_weights": L.MlpWeights,
               "nerfb200_mlp_grads": L.MlpGrads, "nerfb200_kilo_camera": L.KiloCamera, "nerfb200_kilo_grid": L.KiloGrid,
               "nerfb200_kilo_march_params": L.KiloMarchParams}
    lines = ['#include <stdio.h>', '#include "nerfb200.h"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append('  printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ["  return 0;", "}"]
    src = tmp_path / "offsets.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "offsets"
    subprocess.run([gcc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    c_view = dict(zip(out[0::2], (int(v) for v in out[1::2])))
    for cname, cls in structs.items():
        assert c_view[cname] == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert c_view["%s.%s" % (cname, fname)] == getattr(cls, fname).offset, (cname, fname)


def test_integration_md_stub_matches_the_binding():
    """the ctypes stub in INTEGRATION.md section 3 is what a maintainer copies -- its structures must
    be field for field (name, ctype, size) the ones lib.py binds and the header declares."""
    import ctypes as C
    import os
    import re
    from nerf_rep_for_test_b200 import lib as L
    text = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "INTEGRATION.md")).read()
    block = re.search(r"```python\nimport ctypes as C, torch\n(.*?)```", text, re.S).group(1)
    classes = block[block.index("class MlpWeights"):block.index("def pack(")]
    ns = {"C": C}
    exec(classes, ns)
    for stub, ours in ((ns["MlpWeights"], L.MlpWeights), (ns["RenderParams"], L.RenderParams), (ns["Maps"], L.Maps)):
        assert [f[0] for f in stub._fields_] == [f[0] for f in ours._fields_], stub.__name__
        for (name, a), (_, b) in zip(stub._fields_, ours._fields_):
            assert C.sizeof(a) == C.sizeof(b), (stub.__name__, name)
            assert getattr(stub, name).offset == getattr(ours, name).offset, (stub.__name__, name)
        assert C.sizeof(stub) == C.sizeof(ours), stub.__name__
    assert "ABI %d" % L.ABI_VERSION in classes
