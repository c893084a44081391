"""cnn_gp on B200: the reference's public API (cnn_gp/__init__.py:1-6) over sm_100a kernels.

The three API modules keep the reference's names; ``__all__`` is their union, in the
reference's order (kernel algebra, tile iteration, persistence)."""
from . import data, kernel_save_tools, kernels

_api_modules = (kernels, data, kernel_save_tools)
__all__ = tuple(name for mod in _api_modules for name in mod.__all__)
globals().update({name: getattr(mod, name) for mod in _api_modules for name in mod.__all__})
