"""cnn_gp on B200: the reference's public API (cnn_gp/__init__.py:1-6) over sm_100a kernels."""
from . import kernels, data, kernel_save_tools
from .kernels import *  # noqa: F401,F403
from .data import *  # noqa: F401,F403
from .kernel_save_tools import *  # noqa: F401,F403

__all__ = kernels.__all__ + data.__all__ + kernel_save_tools.__all__
