"""ctypes binding of libnerfb200.so (the C ABI declared in include/nerfb200.h).

There is NO fallback: if the library is missing or a call fails this module
raises.  PyTorch only provides device memory and the CUDA stream; every kernel
on the path lives in the shared library.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NERFB200_LIB") or os.path.join(HERE, "libnerfb200.so")   # env override: A/B experiments only

MODE_FP32, MODE_BF16, MODE_FP32_TC, MODE_FP16 = 0, 1, 2, 3


def mode_coarse(m):
    """NERFB200_MODE_COARSE(m): OR-ed into RenderParams.mode to give the coarse pass its own arithmetic mode."""
    return (m + 1) << 8


COMPOSITE_PLAIN, COMPOSITE_ERT, COMPOSITE_ERT_COMPAT = 0, 1, 2
COMPOSITE_FAST_MATH = 0x10          # OR-ed into PLAIN / ERT (include/nerfb200.h)
ABI_VERSION = 4

_f = C.POINTER(C.c_float)
_vp = C.c_void_p


class NerfB200Error(RuntimeError):
    pass


class MlpWeights(C.Structure):
    _fields_ = [("pts_w", _vp * 8), ("pts_b", _vp * 8),
                ("views_w", _vp), ("views_b", _vp), ("feature_w", _vp), ("feature_b", _vp),
                ("alpha_w", _vp), ("alpha_b", _vp), ("rgb_w", _vp), ("rgb_b", _vp)]


class MlpGrads(C.Structure):
    _fields_ = MlpWeights._fields_


class RenderParams(C.Structure):
    _fields_ = [("n_samples", C.c_int), ("n_importance", C.c_int), ("mode", C.c_int),
                ("variant", C.c_int), ("white_bkgd", C.c_int), ("perturb", C.c_int),
                ("u_per_ray", C.c_int), ("compat_chunk", C.c_int), ("ert_threshold", C.c_float),
                ("raw_noise_std", C.c_float), ("seed", C.c_uint64), ("occupancy_grid", _vp), ("grid_res", C.c_int), ("ess_skip", C.c_int),
                ("eval_counts", _vp), ("cull_rays", C.c_int), ("cull_lo", C.c_float * 3), ("cull_hi", C.c_float * 3),
                ("ess_ref_compat", C.c_int)]


class KiloCamera(C.Structure):
    _fields_ = [("H", C.c_int), ("W", C.c_int), ("cx", C.c_float), ("cy", C.c_float), ("fx", C.c_float),
                ("fy", C.c_float), ("c2w", C.c_float * 9), ("origin", C.c_float * 3)]


class KiloGrid(C.Structure):
    _fields_ = [("res", C.c_int * 3), ("gmin", C.c_float * 3), ("gmax", C.c_float * 3)]


class KiloMarchParams(C.Structure):
    _fields_ = [("distance_between_points", C.c_float), ("max_samples_per_ray", C.c_int),
                ("max_depth_index", C.c_int), ("min_distance", C.c_float), ("transmittance_threshold", C.c_float),
                ("white_bkgd", C.c_int), ("max_passes", C.c_int)]


class Maps(C.Structure):
    _fields_ = [("rgb", _vp), ("disp", _vp), ("acc", _vp), ("depth", _vp)]


# name -> (restype, argtypes); must list every symbol include/nerfb200.h declares
SIGNATURES = {
    "nerfb200_abi_version": (C.c_int, []),
    "nerfb200_get_last_error_string": (C.c_char_p, []),
    "nerfb200_launch_count": (C.c_uint64, []),
    "nerfb200_profile_enable": (C.c_int, [C.c_int]),
    "nerfb200_profile_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_double)]),
    "nerfb200_raygen": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "nerfb200_sample_coarse": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_uint64, _vp, _vp]),
    "nerfb200_packed_weights_bytes": (C.c_size_t, [C.c_int]),
    "nerfb200_pack_weights": (C.c_int, [C.POINTER(MlpWeights), C.c_int, _vp, _vp]),
    "nerfb200_mlp_forward": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "nerfb200_train_acts_bytes": (C.c_size_t, [C.c_longlong]),
    "nerfb200_train_masks_bytes": (C.c_size_t, [C.c_longlong]),
    "nerfb200_mlp_backward_workspace_bytes": (C.c_size_t, [C.c_longlong]),
    "nerfb200_mlp_forward_train": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "nerfb200_packed_bwd_bytes": (C.c_size_t, []),
    "nerfb200_pack_weights_bwd": (C.c_int, [C.POINTER(MlpWeights), _vp, _vp]),
    "nerfb200_pack_weights_bwd2": (C.c_int, [C.POINTER(MlpWeights), _vp, _vp, _vp]),
    "nerfb200_mlp_backward": (C.c_int, [_vp, C.POINTER(MlpWeights), _vp, _vp, _vp, C.c_longlong, _vp, C.c_size_t,
                                        C.POINTER(MlpGrads), _vp]),
    "nerfb200_train_fp32_acts_bytes": (C.c_size_t, [C.c_longlong]),
    "nerfb200_train_fp32_workspace_bytes": (C.c_size_t, [C.c_longlong]),
    "nerfb200_mlp_forward_train_fp32": (C.c_int, [C.POINTER(MlpWeights), _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "nerfb200_mlp_backward_fp32": (C.c_int, [C.POINTER(MlpWeights), _vp, _vp, _vp, C.c_int, C.c_int, _vp, C.c_size_t,
                                             C.POINTER(MlpGrads), _vp, _vp]),
    "nerfb200_composite_backward_z": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp,
                                                _vp, _vp, _vp]),
    "nerfb200_sample_pdf_backward": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "nerfb200_mlp_backward_input": (C.c_int, [C.POINTER(MlpWeights), _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "nerfb200_mse_pair_grad": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp]),
    "nerfb200_adam_clip_step": (C.c_int, [_vp, _vp, _vp, _vp, C.c_longlong, C.c_double, C.c_double, C.c_double, C.c_double,
                                          C.c_longlong, C.c_float, C.c_float, _vp]),
    "nerfb200_mlp_forward_stages": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "nerfb200_composite_forward": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                             C.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nerfb200_composite_forward_masked": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                                    C.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nerfb200_sigma_noise": (C.c_int, [_vp, C.c_longlong, C.c_float, C.c_uint64, _vp]),
    "nerfb200_composite_backward": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp,
                                              _vp, _vp]),
    "nerfb200_sample_from_cdf": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "nerfb200_sample_pdf_merge": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp,
                                            _vp, _vp]),
    "nerfb200_ess_resample": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "nerfb200_ess_resample_compat": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "nerfb200_jitter_rows": (C.c_int, [_vp, C.c_int, C.c_int, C.c_uint64, _vp]),
    "nerfb200_ess_update": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp]),
    "nerfb200_ess_compact": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "nerfb200_mlp_forward_sparse": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "nerfb200_ray_cull": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), _vp, _vp]),
    "nerfb200_ert_depth": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_float, _vp, _vp]),
    "nerfb200_accumulate_counts": (C.c_int, [_vp, _vp, _vp]),
    "nerfb200_kilo_param_size": (C.c_int, []),
    "nerfb200_kilo_rays_d": (C.c_int, [C.POINTER(KiloCamera), _vp, _vp]),
    "nerfb200_kilo_march": (C.c_int, [C.POINTER(KiloGrid), _vp, _vp, _vp, C.c_int, C.c_float, C.c_int, C.c_int, C.c_float,
                                      C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "nerfb200_kilo_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "nerfb200_kilo_network_eval": (C.c_int, [C.POINTER(KiloCamera), C.POINTER(KiloMarchParams), _vp, _vp, C.c_int, _vp,
                                             _vp, _vp, C.c_int, _vp, C.c_size_t, _vp, _vp]),
    "nerfb200_kilo_integrate": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_float, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "nerfb200_kilo_render": (C.c_int, [C.POINTER(KiloCamera), C.POINTER(KiloGrid), C.POINTER(KiloMarchParams), _vp, _vp, _vp,
                                       _vp, C.c_int, _vp, C.c_size_t, _vp, _vp, _vp, _vp]),
    "nerfb200_render_workspace_bytes": (C.c_size_t, [C.c_int, C.POINTER(RenderParams)]),
    "nerfb200_render_rays": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _vp, _vp, C.POINTER(RenderParams), _vp,
                                       C.c_size_t, C.POINTER(Maps), C.POINTER(Maps), _vp]),
    "nerfb200_render_image_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.POINTER(RenderParams)]),
    "nerfb200_render_image_host": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp,
                                             C.POINTER(RenderParams), _vp, C.c_size_t, C.POINTER(Maps),
                                             C.POINTER(Maps), _vp]),
}

_LIB = None


def load():
    """Load libnerfb200.so (once).  Raises NerfB200Error when it is missing -- no fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise NerfB200Error(
            "libnerfb200.so not found at %s -- build it with `python -m nerf_rep_for_test_b200.build` "
            "(there is no CPU or PyTorch fallback for this path)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.nerfb200_abi_version() != ABI_VERSION:
        raise NerfB200Error("libnerfb200.so ABI %d != binding ABI %d" % (lib.nerfb200_abi_version(), ABI_VERSION))
    _LIB = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().nerfb200_get_last_error_string().decode("utf-8", "replace")
        raise NerfB200Error("%s failed (code %d): %s" % (what or "libnerfb200 call", rc, msg))


def profile_enable(on):
    check(load().nerfb200_profile_enable(int(bool(on))), "profile_enable")


def profile_read():
    """(total MLP-kernel ms, launches, MLP rows) since profile_enable(True)."""
    ms, n, rows = C.c_double(), C.c_uint64(), C.c_double()
    check(load().nerfb200_profile_read(C.byref(ms), C.byref(n), C.byref(rows)), "profile_read")
    return ms.value, int(n.value), rows.value


def launch_count():
    return int(load().nerfb200_launch_count())


def ptr(t):
    """device (or host) pointer of a tensor, None -> NULL; requires contiguity."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise NerfB200Error("tensor passed to libnerfb200 must be contiguous")
    return C.c_void_p(t.data_ptr())


def dev(t, dtype=None):
    """check that `t` is a contiguous CUDA tensor (of dtype) and return its pointer."""
    import torch
    if t is None:
        return None
    if not t.is_cuda:
        raise NerfB200Error("libnerfb200 needs CUDA tensors (got %s); there is no CPU path" % t.device)
    if dtype is not None and t.dtype != dtype:
        raise NerfB200Error("expected dtype %s, got %s" % (dtype, t.dtype))
    return ptr(t)


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
