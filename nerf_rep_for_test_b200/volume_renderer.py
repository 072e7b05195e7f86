"""`renderer_module` entry point.

Select with `renderer_module nerf_rep_for_test_b200.volume_renderer` (configs/nerf/lego.yaml:9 ->
config.py:180-182 -> make_renderer.py:4-8 does imp.load_source(module, path).Renderer(net)), or
import `Renderer` from here in place of src.models.nerf.renderer.volume_renderer (hard imports at
trainers/nerf.py:4, test_ess_ert.py:23, quick_test_ess_ert.py:24).
"""
from nerf_rep_for_test_b200.renderer import Renderer  # noqa: F401
