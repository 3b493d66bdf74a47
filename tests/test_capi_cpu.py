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


def test_tcgen05_planner_invariants():
    """Device-free sweep of the tcgen05 conv planner (fyn_conv2d_plan_query): whatever it accepts must respect the
    limits the kernel relies on -- shared memory, ownership of ring slots / stages by loader groups, the lower bound on
    the ring that keeps two MMA-issuing warps from deadlocking, TMEM columns, the epilogue's column split."""
    import itertools
    from fyusenet_b200 import capi
    seen = 0
    for k, ci, co, kind, relu, res, bn in itertools.product([3, 5, 7, 9], [1, 3, 4, 8, 12, 20, 40, 64, 128], [1, 3, 12, 20, 40, 48, 64],
                                                            ["plain", "s2", "f05d2", "f05", "f025d2"], [0, 1], [0, 1], [0, 1]):
        kw = {"plain": {}, "s2": dict(downsample=2), "f05d2": dict(downsample=2, source_step=0.5, fractional=1),
              "f05": dict(source_step=0.5, fractional=1), "f025d2": dict(downsample=2, source_step=0.25, fractional=1)}[kind]
        flags = (capi.FLAG_PRE_RELU if relu else 0) | (capi.FLAG_RESIDUAL_INPUT if res else 0) | (capi.FLAG_POST_BATCHNORM if bn else 0)
        for ys in (1, 2):
            p = capi.conv_plan_query(ys, width=512, height=256, in_channels=ci, out_channels=co, kernel=k, flags=flags, **kw)
            if p is None:
                continue
            seen += 1
            what = f"k{k} {ci}->{co} {kind} ys{ys} flags {flags:#x}"
            assert p.shared_bytes <= 227 * 1024, what
            assert 16 <= p.n <= 64 and p.n % 16 == 0, what
            assert 1 <= p.steps <= 96 and p.row_items <= 1152, what
            assert p.epilogue_warps in (8, 12) and (p.n // 8) % (p.epilogue_warps // 4) == 0, what
            assert 1 <= p.loader_groups <= 16 - p.epilogue_warps and (16 - p.epilogue_warps) % p.loader_groups == 0, what
            assert p.ring_slots % p.loader_groups == 0 and p.staged_rows % p.loader_groups == 0 and p.staged_rows >= p.loader_groups, what
            # progress with two MMA warps needs max(window, 2 * advance) slots; the window must also fit contiguously
            assert p.ring_slots >= max(p.window_rows, 2 * p.row_advance), what
            want_mirror = max(0, p.window_rows - p.row_advance) if p.ring_slots % p.row_advance == 0 else p.window_rows - 1
            assert p.mirror_slots == want_mirror, what
            assert p.phases_y % ys == 0 and p.weight_image_bytes % 16 == 0, what
            assert not (p.bias_folded and bn), what
    assert seen > 500
