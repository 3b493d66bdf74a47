"""CPU-only checks of the C-ABI library: it loads, exports every declared symbol, its host-side
geometry agrees with the oracle's layout rules, and it refuses to run without a GPU (no fallback)."""
import re
from pathlib import Path

import numpy as np
import pytest

import fyn_oracle as fo
from fyusenet_b200 import capi

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "fyusenet_b200.h").read_text()
    declared = set(re.findall(r"\b(fyn_[a-z0-9_]+)\s*\(", header))
    declared -= {"fyn_status"}
    assert declared, "no declarations found"
    L = capi.lib()
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, f"library lacks {missing}"
    assert set(capi.EXPORTS) == declared, sorted(set(capi.EXPORTS) ^ declared)
    assert L.fyn_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.FynError, match="no CUDA device|no CPU fallback"):
        capi.Context(0)


@pytest.mark.parametrize("c", list(range(1, 70)) + [128, 256, 512, 1000, 1024, 2048])
def test_deep_geometry_matches_reference_rule(c):
    """fyn_tensor_geometry vs oracle's restatement of cpu/cpubuffershape.cpp:430-447 + deeptiler.cpp:91-94."""
    for (w, h, p) in [(7, 7, 0), (14, 9, 1), (56, 56, 2)]:
        g = capi.tensor_geometry(w, h, c, p, capi.ORDER_DEEP)
        assert (g.tiles_x, g.tiles_y) == fo.deep_tiling(c)
        assert (g.tex_width, g.tex_height) == fo.deep_texture_size(c, w, h, p)
        assert g.bytes == g.tex_width * g.tex_height * 4 * 2


def test_shallow_geometry():
    g = capi.tensor_geometry(10, 6, 23, 1, capi.ORDER_SHALLOW, capi.F32, batch=3)
    assert (g.planes, g.tex_width, g.tex_height, g.packing) == (6, 12, 8, 4)
    assert g.bytes == 3 * 6 * 12 * 8 * 4 * 4
    g = capi.tensor_geometry(10, 6, 3, 0, capi.ORDER_SHALLOW, capi.F32, packing=3)
    assert g.bytes == 10 * 6 * 3 * 4
    with pytest.raises(capi.FynError):
        capi.tensor_geometry(10, 6, 8, 0, capi.ORDER_SHALLOW, capi.F32, packing=3)
    with pytest.raises(capi.FynError):
        capi.tensor_geometry(0, 6, 8)


def test_conv_output_size_rules():
    import ctypes as C
    def size(**kw):
        d = capi.ConvDesc(**kw)
        ow, oh = C.c_int(), C.c_int()
        capi.check(capi.lib().fyn_conv2d_output_size(C.byref(d), C.byref(ow), C.byref(oh)))
        return ow.value, oh.value
    assert size(width=381, height=464, downsample=2, source_step=0.5, fractional=1) == (381, 464)
    assert size(width=381, height=464, downsample=2, source_step=0.25, fractional=1) == (762, 928)
    assert size(width=762, height=928, downsample=1, source_step=0.5, fractional=1) == (1524, 1856)
    assert size(width=1524, height=1856, downsample=2, source_step=1.0, fractional=0) == (762, 928)
