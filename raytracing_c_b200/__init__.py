"""B200-native path-tracing core behind the raytracer.h / denoiser.h entry points.

Native code lives in csrc/ (sm_100a CUDA + the C ABI) and host/ (plain-C loaders, BVH build,
codecs, CLI).  The Python layer is ctypes plumbing only; see DESIGN.md and INTEGRATION.md.
"""
from . import _ffi, driver  # noqa: F401
from ._ffi import build_native, gpu_lib, host_lib  # noqa: F401

__all__ = ["_ffi", "driver", "build_native", "gpu_lib", "host_lib"]
