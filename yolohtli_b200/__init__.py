"""yolohtli_b200 -- B200-native (sm_100a) implementation of Yolohtli's 2D cardiac monodomain
time step behind the reference's own launch API.

The product is the C-ABI shared library ``lib/libyolohtli_b200.so`` (``include/yolohtli_abi.h``)
built from the hand-written CUDA kernels under ``csrc/``.  This package is only the thin host
mirror used by tests and ``bench.py``: ctypes bindings (``_lib``), the parameter block and
headless driver (``host``), synthetic inputs (``synth``) and the row-slab multi-GPU driver
(``slab``).  There is no CPU fallback: importing ``_lib`` without the built library, or
calling a compute entry point without a CUDA device, raises.
"""
from .host import Params, Sim, default_params, check  # noqa: F401
from . import host, io, slab, synth  # noqa: F401
from ._lib import lib, load_library, LIB_PATH, YolohtliError  # noqa: F401

__all__ = ["Params", "Sim", "default_params", "check", "lib", "load_library", "LIB_PATH",
           "YolohtliError"]
