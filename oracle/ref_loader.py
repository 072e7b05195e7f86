"""Loader for the UNMODIFIED reference renderer (test infrastructure only).

This file is part of the oracle: it is test infrastructure, not product code.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import it.

It imports /root/reference (read-only tree) exactly as SURVEY.md Appendix A
describes: two sys.modules shims ('imp' was removed in Python 3.12, 'imageio'
is only needed by the reference's video helpers), argv preset because
src/config/config.py:207-217 runs argparse at import time, and cwd set to the
reference root because config.py:180-182 derives relative module paths.

/root/reference does not exist on the GPU box; callers must check
`reference_available()` first.  Nothing here is ever on the product path.
"""
import contextlib
import importlib.machinery
import importlib.util
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("NERFB200_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(
        REFERENCE_ROOT, "src/models/nerf/renderer/volume_renderer.py"))


def _install_shims():
    if "imp" not in sys.modules:
        imp = types.ModuleType("imp")

        def _load_source(name, path):
            loader = importlib.machinery.SourceFileLoader(name, path)
            spec = importlib.util.spec_from_loader(name, loader)
            mod = importlib.util.module_from_spec(spec)
            loader.exec_module(mod)
            return mod

        imp.load_source = _load_source
        sys.modules["imp"] = imp
    if "imageio" not in sys.modules:
        sys.modules["imageio"] = types.ModuleType("imageio")


_CACHE = {}


def load_reference(quiet=True):
    """Return (cfg, make_network, make_renderer) of the unmodified reference."""
    if "mods" in _CACHE:
        return _CACHE["mods"]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    _install_shims()
    old_argv, old_cwd = sys.argv, os.getcwd()
    sys.argv = ["oracle", "--cfg_file", "configs/nerf/lego.yaml"]
    os.chdir(REFERENCE_ROOT)
    sys.path.insert(0, REFERENCE_ROOT)
    sink = io.StringIO()
    try:
        with (contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext()), \
             (contextlib.redirect_stderr(sink) if quiet else contextlib.nullcontext()):
            from src.config import cfg
            from src.models import make_network
            from src.models.nerf.renderer import make_renderer
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)
    _CACHE["mods"] = (cfg, make_network, make_renderer)
    return _CACHE["mods"]


def build_reference(state_dict=None, enable_ess=False, enable_ert=False, quiet=True):
    """Build reference Network + Renderer on CPU, perturb=0, eval()."""
    import torch
    cfg, make_network, make_renderer = load_reference(quiet)
    old_cwd = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    sink = io.StringIO()
    try:
        with (contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext()):
            cfg.defrost() if hasattr(cfg, "defrost") else None
            cfg.enable_ess = bool(enable_ess)
            cfg.enable_ert = bool(enable_ert)
            net = make_network(cfg)
            if state_dict is not None:
                net.load_state_dict({k: torch.as_tensor(v) for k, v in state_dict.items()})
            net.eval()
            # network.py:135 picks cuda when visible; the oracle is a CPU oracle.
            net.device = torch.device("cpu")
            r = make_renderer(cfg, net)
            r.device = torch.device("cpu")
            r.perturb = 0
    finally:
        os.chdir(old_cwd)
    return cfg, net, r
