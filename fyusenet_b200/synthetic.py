"""Synthetic inputs for benchmarks and smoke runs (the reference's .dat files are git-LFS stubs).

Weight blobs follow the reference's file layout for the StyleNet samples
(samples/samplenetworks/stylenet9x9.cpp:41-56, stylenet3x3.cpp:41-50: per layer `bias[Co]` then `W[Co][Ky][Kx][Ci]`,
layers in the order conv1..3, deconv1..3, res blocks).  The generators are bit-identical to the ones the parity
oracle uses (tests/test_host_cpu.py checks that), but live here so that the product path never imports `oracle/`.
"""
from __future__ import annotations

import numpy as np


def stylenet_file_layers(ksize: int):
    """(name, kernel, cin, cout) in weight-file order."""
    nres = 5 if ksize == 9 else 2
    layers = [("conv1", ksize, 3, 12), ("conv2", 3, 12, 20), ("conv3", 3, 20, 40),
              ("deconv1", 3, 40, 20), ("deconv2", 3, 20, 12), ("deconv3", ksize, 12, 3)]
    for r in range(1, nres + 1):
        layers += [(f"res{r}_1", 3, 40, 40), (f"res{r}_2", 3, 40, 40)]
    return layers


def stylenet_weights(ksize: int, seed: int | None = None) -> np.ndarray:
    """He-normal conv weights, U(-0.05, 0.05) biases, numpy PCG64 (seed 112 for 3x3, 9112 for 9x9); the second conv
    of every residual block is scaled by 0.5 so that the stacked blocks stay inside the fp16 range."""
    if seed is None:
        seed = 9112 if ksize == 9 else 112
    rng = np.random.default_rng(seed)
    parts = []
    for name, k, ci, co in stylenet_file_layers(ksize):
        bias = rng.uniform(-0.05, 0.05, co)
        w = rng.normal(0.0, np.sqrt(2.0 / (k * k * ci)), (co, k, k, ci))
        if name.startswith("res") and name.endswith("_2"):
            w *= 0.5
        parts += [bias.astype(np.float32), w.reshape(-1).astype(np.float32)]
    return np.concatenate(parts)


def image(h: int, w: int, index: int = 0) -> np.ndarray:
    """float32 [H][W][3] in [0,1): the sample's uint8/255 input (samples/desktop/stylenet.cpp:45-47)."""
    rng = np.random.default_rng(1000 + index)
    return rng.random((h, w, 3), dtype=np.float32)
