"""animatablegaussians_b200 — B200-native (sm_100a) implementation of the AnimatableGaussians
render/train hot path: DualStyleUNet -> per-Gaussian LBS -> differentiable 3D-Gaussian rasterizer
(RGB + depth + alpha).  Host code is Python/PyTorch; all arithmetic is hand-written CUDA behind the
C ABI declared in include/*.h (libagr_b200.so).  There is no CPU or PyTorch fallback.
"""
from . import _lib  # noqa: F401  (defines the ctypes binding; the .so is loaded on first use)

__version__ = "0.1.0"
