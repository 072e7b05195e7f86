"""B200-native NeRF volume-rendering hot path (drop-in Renderer + C-ABI CUDA library).

The package only holds what the path needs: csrc/ (sm_100a kernels + the C ABI of
include/nerfb200.h), the ctypes binding, and the host-side mirror of the reference's
Renderer / Network interface.  Importing the package does not load CUDA; constructing a
Renderer or calling an op does, and fails loudly when the library or a GPU is missing.
"""
from .lib import NerfB200Error, MODE_BF16, MODE_FP32  # noqa: F401
from .network import NeRF, Network  # noqa: F401
from .renderer import RenderConfig, Renderer  # noqa: F401
