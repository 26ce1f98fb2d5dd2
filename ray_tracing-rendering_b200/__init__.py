"""B200-native path-tracing core behind the renderer API of
JiGuang283/Ray_Tracing-Rendering.  The product is librtb200.so (hand-written
sm_100a CUDA behind the C-ABI of include/rtb200.h); this package is the thin
Python plumbing used by the tests, bench.py and the multi-GPU launcher.

The directory name contains a hyphen, so import it with
``importlib.import_module("ray_tracing-rendering_b200")``.
"""
from . import abi  # noqa: F401
from .binding import Context, Group, RenderParams, RtbError, load  # noqa: F401
from .renderer import Renderer, INTEGRATOR_NAMES  # noqa: F401
