"""Pins the CPU oracle against the REFERENCE'S OWN code, compiled here: oracle/_ref/libfyn_ref_kat.so.

oracle/build_ref.py compiles the plain-C++ ground-truth helpers of the reference's unit tests (paddedConvolution, batchnorm,
computeMaxPool / computeAvgPool, stackConvolution) and CPUBufferShape::computeDeepTiling from the sources where they lie under
/root/reference.  These tests run the reference code and oracle/fyn_oracle.c on the same inputs over the reference's own
parameter grids (convlayertests.cpp:426-484, pooltests.cpp:323-355, misctests.cpp:277-286).  The GPU box has no
/root/reference: the prebuilt library travels with the snapshot; if it is missing altogether the module is skipped.
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import pytest

import fyn_oracle as fo

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle"))
import build_ref  # noqa: E402

_LIB = None


def ref():
    global _LIB
    if _LIB is None:
        p = build_ref.build()
        if p is None or not Path(p).exists():
            pytest.skip("oracle/_ref not built and /root/reference not present")
        lib = C.CDLL(str(p))
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
        lib.fynref_padded_convolution.restype = fp
        lib.fynref_padded_convolution.argtypes = [fp, fp] + [C.c_int] * 9
        lib.fynref_batchnorm.restype = fp
        lib.fynref_batchnorm.argtypes = [fp, fp, fp, C.c_int, C.c_int, C.c_int]
        for f in (lib.fynref_max_pool, lib.fynref_avg_pool):
            f.restype = fp
            f.argtypes = [C.c_int] * 4 + [fp] + [C.c_int] * 3
        lib.fynref_stack_convolution.restype = fp
        lib.fynref_stack_convolution.argtypes = [C.c_float, fp] + [C.c_int] * 4
        lib.fynref_deep_tiling.argtypes = [C.c_int, ip, ip]
        lib.fynref_free.argtypes = [fp]
        _LIB = lib
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _take(ptr, shape):
    out = np.ctypeslib.as_array(ptr, shape=(int(np.prod(shape)),)).reshape(shape).copy()
    ref().fynref_free(ptr)
    return out


def ref_padded_conv(x_padded, wb, co, k, down=1, pre_relu=False):
    ci, h, w = x_padded.shape
    x = np.ascontiguousarray(x_padded, np.float32)
    wbc = np.ascontiguousarray(wb, np.float32)
    pad = (k - 1) // 2
    oh, ow = (h - 2 * pad) // down, (w - 2 * pad) // down
    return _take(ref().fynref_padded_convolution(_p(x), _p(wbc), co, k, k, ci, w, h, down, down, int(pre_relu)), (co, oh, ow))


def ref_stack(bias, kern, ci, co):
    kern = np.ascontiguousarray(kern, np.float32)
    ky, kx = kern.shape
    return _take(ref().fynref_stack_convolution(bias, _p(kern), kx, ky, ci, co), (co + ky * kx * ci * co,))


GRID_NXN = [(64, 64, 4, 4), (64, 80, 4, 4), (128, 80, 4, 8), (128, 80, 16, 8), (256, 128, 12, 8)]
GRID_1x1 = [(64, 64, 4, 4), (64, 80, 4, 4), (128, 80, 4, 8), (56, 56, 64, 64), (128, 80, 16, 8), (256, 128, 12, 4)]


def test_library_exports():
    lib = ref()
    for name in ("fynref_padded_convolution", "fynref_batchnorm", "fynref_max_pool", "fynref_avg_pool", "fynref_stack_convolution",
                 "fynref_deep_tiling", "fynref_free"):
        assert hasattr(lib, name)


@pytest.mark.parametrize("channels", list(range(1, 130)) + [256, 512, 1000, 1024, 2048])
def test_deep_tiling_matches_reference(channels):
    """cpu/cpubuffershape.cpp:430-447 compiled from the reference vs the oracle's tiling rule."""
    tx, ty = C.c_int(), C.c_int()
    ref().fynref_deep_tiling(channels, C.byref(tx), C.byref(ty))
    assert fo.deep_tiling(channels) == (tx.value, ty.value)


def test_weight_block_order_matches_reference():
    """layertestbase.cpp:45-60: the [bias | O][Ky][Kx][I] block the oracle and the loaders consume."""
    kern = np.arange(15, dtype=np.float32).reshape(3, 5)
    blk = ref_stack(0.25, kern, 3, 2)
    assert np.all(blk[:2] == 0.25)
    w = blk[2:].reshape(2, 3, 5, 3)
    assert np.array_equal(w, np.broadcast_to(kern[None, :, :, None], (2, 3, 5, 3)))


@pytest.mark.parametrize("pre_relu", [False, True])
@pytest.mark.parametrize("ds", [1, 2])
@pytest.mark.parametrize("k", [3, 5, 7])
@pytest.mark.parametrize("w,h,ci,co", GRID_NXN[:4])
def test_deep_conv_matches_reference_code(w, h, ci, co, k, ds, pre_relu):
    """Deep conv (zero padding in the tile gaps) == the reference's paddedConvolution on a zero-padded tensor, random weights
    and inputs (convlayertests.cpp:277-420 compare the GL layer with exactly this function)."""
    rng = np.random.default_rng(k * 1000 + w + ci * 7 + ds)
    x = rng.uniform(-2, 2, (ci, h, w)).astype(np.float32)
    wb = np.concatenate([rng.uniform(-1, 1, co), rng.normal(0, 0.2, co * k * k * ci)]).astype(np.float32)
    pad = (k - 1) // 2
    xp = np.pad(x, ((0, 0), (pad, pad), (pad, pad)))
    want = ref_padded_conv(xp, wb, co, k, ds, pre_relu)
    got = fo.conv2d(x, wb, co, k, downsample=ds, in_pad=pad, deep=True, act=fo.ACT_RELU if pre_relu else fo.ACT_NONE, prec=fo.FP32)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-5 * max(1.0, float(np.abs(want).max())))


@pytest.mark.parametrize("ds", [1, 2])
@pytest.mark.parametrize("w,h,ci,co", GRID_1x1)
def test_conv1x1_matches_reference_code(w, h, ci, co, ds):
    rng = np.random.default_rng(w + h + ci + co + ds)
    x = rng.uniform(-2, 2, (ci, h, w)).astype(np.float32)
    wb = np.concatenate([rng.uniform(-1, 1, co), rng.normal(0, 0.3, co * ci)]).astype(np.float32)
    want = ref_padded_conv(x, wb, co, 1, ds)
    for deep in (False, True):
        got = fo.conv2d(x, wb, co, 1, downsample=ds, deep=deep, prec=fo.FP32)
        np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-5 * max(1.0, float(np.abs(want).max())))


@pytest.mark.parametrize("k", [3, 5, 7, 9])
def test_shallow_conv_interior_matches_reference_code(k):
    """Shallow convs clamp to edge (no reference KAT beyond 'result is 0'): away from the border they must equal the
    reference's valid convolution.  Covers the 9x9 kernel of StyleNet conv1."""
    rng = np.random.default_rng(k)
    ci, co, h, w = 3, 12, 40, 48
    x = rng.uniform(0, 1, (ci, h, w)).astype(np.float32)
    wb = np.concatenate([rng.uniform(-1, 1, co), rng.normal(0, 0.1, co * k * k * ci)]).astype(np.float32)
    want = ref_padded_conv(x, wb, co, k, 1, True)            # valid area only
    got = fo.conv2d(x, wb, co, k, act=fo.ACT_RELU, prec=fo.FP32)
    m = (k - 1) // 2
    np.testing.assert_allclose(got[:, m:h - m, m:w - m], want, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("w,h,c", [(8, 8, 4), (200, 200, 4), (80, 40, 12), (50, 50, 23), (40, 40, 80)])
def test_pools_match_reference_code(w, h, c):
    """pooltests.cpp:65-117 compiled from the reference vs the oracle's 2x2 stride-2 pools (grid :323-333)."""
    rng = np.random.default_rng(w * h + c)
    x = rng.uniform(-3, 3, (c, h, w)).astype(np.float32)
    for is_max, fn in ((True, ref().fynref_max_pool), (False, ref().fynref_avg_pool)):
        want = _take(fn(2, 2, 2, 2, _p(x), w, h, c), (c, h // 2, w // 2))
        for deep_pad in (0,):
            got = fo.pool2d(x, pool=2, downsample=2, in_pad=deep_pad, is_max=is_max, prec=fo.FP32)
            np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("w,h,c", [(4, 4, 36), (80, 40, 52), (4, 4, 4), (256, 128, 64), (120, 80, 3), (200, 200, 4), (50, 50, 31), (12, 12, 128)])
def test_batchnorm_matches_reference_code(w, h, c):
    """convlayertests.cpp batchnorm() helper vs the oracle's shallow and deep batch-norm (grid misctests.cpp:277-286)."""
    rng = np.random.default_rng(w + 3 * h + 7 * c)
    x = rng.uniform(-3, 3, (c, h, w)).astype(np.float32)
    sc, bi = rng.uniform(0.5, 2, c).astype(np.float32), rng.uniform(-1, 1, c).astype(np.float32)
    want = _take(ref().fynref_batchnorm(_p(x), _p(sc), _p(bi), w, h, c), (c, h, w))
    for deep in (False, True):
        got = fo.batchnorm(x, np.concatenate([sc, bi]), deep=deep, prec=fo.FP32)
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)
