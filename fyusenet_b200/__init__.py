"""fyusenet_b200 -- B200-native (sm_100a) backend for the FyuseNet GPU-layer path.

Holds only what the hot path needs: csrc/ (CUDA kernels + the C ABI of include/fyusenet_b200.h),
host/ (C++ host engine mirroring the reference's LayerBuilder / LayerFactory / NeuralNetwork API)
and this thin ctypes binding.  No CPU fallback exists anywhere in the package.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
