"""Oracle checks for the SURVEY 8f rank-2 layers: the reference's arithmetic known-answer tests
(unit_tests/arithtests.cpp:97-250, re-expressed against the oracle because they need a GL context there) and
closed-form properties of the scaling / concat / swizzle restatements."""
import numpy as np
import pytest

import fyn_oracle as fo

# unit_tests/arithtests.cpp:208-247: (operand1, operand2, width, height, channels)
ARITH_PARAMS = [(3.0, 30.0, 400, 300, 4), (-2.0, 1.0, 200, 200, 5), (10.0, -10.0, 16, 16, 40), (-100.0, 23.0, 55, 57, 30),
                (15.0, -16.0, 99, 52, 47)]
EXPECT = {fo.ARITH_ADD: lambda a, b: a + b, fo.ARITH_SUB: lambda a, b: a - b, fo.ARITH_MUL: lambda a, b: a * b,
          fo.ARITH_DIV: lambda a, b: a / b}


@pytest.mark.parametrize("op", [fo.ARITH_ADD, fo.ARITH_SUB, fo.ARITH_MUL, fo.ARITH_DIV])
@pytest.mark.parametrize("p", ARITH_PARAMS)
def test_singleton_arith_kat(p, op):
    """arithtests.cpp:97-171 (SingletonTestShallow / Deep): constant tensor (op) scalar, ASSERT_NEAR 0.5."""
    a, b, w, h, c = p
    x = np.full((c, h, w), a, np.float32)
    for prec in (fo.FP32, fo.FP16_STORE):
        y = fo.arith(x, b, op, prec=prec)
        assert y.shape == x.shape
        assert np.all(np.abs(y - EXPECT[op](a, b)) <= 0.5)
    np.testing.assert_array_equal(fo.arith(x, b, op), np.float32(EXPECT[op](np.float32(a), np.float32(b))))


@pytest.mark.parametrize("op", [fo.ARITH_ADD, fo.ARITH_SUB])
@pytest.mark.parametrize("p", ARITH_PARAMS)
def test_addsub_kat(p, op):
    """arithtests.cpp:173-206 (ArithTestShallow): two constant tensors, ASSERT_NEAR 0.5."""
    a, b, w, h, c = p
    y = fo.arith(np.full((c, h, w), a, np.float32), np.full((c, h, w), b, np.float32), op)
    np.testing.assert_array_equal(y, np.float32(EXPECT[op](a, b)))


def test_arith_activation_at_fetch_and_errors():
    rng = np.random.default_rng(0)
    a, b = rng.normal(size=(5, 6, 7)).astype(np.float32), rng.normal(size=(5, 6, 7)).astype(np.float32)
    np.testing.assert_array_equal(fo.arith(a, b, fo.ARITH_SUB, act=fo.ACT_RELU), np.maximum(a, 0) - np.maximum(b, 0))
    with pytest.raises(RuntimeError):
        fo.arith(a, b, fo.ARITH_MUL)   # the two-tensor layer only adds / subtracts (gpu/addsublayer.cpp)


@pytest.mark.parametrize("deep", [False, True])
def test_scale_nearest(deep):
    rng = np.random.default_rng(1)
    x = rng.normal(size=(9, 6, 8)).astype(np.float32)
    for pad in (0, 1):
        # integer up-scaling repeats texels; factor 1 is the identity (PADDING2D) / activation (RELU) pseudo-layer
        np.testing.assert_array_equal(fo.scale(x, up=(2, 3), in_pad=pad, deep=deep), x.repeat(3, axis=1).repeat(2, axis=2))
        np.testing.assert_array_equal(fo.scale(x, in_pad=pad, deep=deep), x)
        np.testing.assert_array_equal(fo.scale(x, in_pad=pad, deep=deep, act=fo.ACT_RELU), np.maximum(x, 0))
        # down-scaling by 2 samples at coordinate 2o+1: the second texel of every pair
        np.testing.assert_array_equal(fo.scale(x, down=(2, 2), in_pad=pad, deep=deep), x[:, 1::2, 1::2])
        np.testing.assert_array_equal(fo.scale(x[:, :, :6], down=(3, 1), in_pad=pad, deep=deep), x[:, :, 1:6:3])
    y = fo.scale(x, act=fo.ACT_CLIP, lo=-0.25, hi=0.5)
    np.testing.assert_array_equal(y, np.clip(x, -0.25, 0.5))


def test_scale_linear():
    rng = np.random.default_rng(2)
    x = rng.normal(size=(4, 6, 8)).astype(np.float32)
    # GL_LINEAR down-scaling by 2 is the 2x2 box filter
    box = (x[:, 0::2, 0::2] + x[:, 0::2, 1::2] + x[:, 1::2, 0::2] + x[:, 1::2, 1::2]) / 4
    for deep in (False, True):
        np.testing.assert_allclose(fo.scale(x, down=(2, 2), linear=True, in_pad=1, deep=deep), box, rtol=1e-6, atol=1e-6)
    # up-scaling by 2: interior texels blend 3:1, the border clamps without padding and fades into the zero padding with it
    row = np.arange(8, dtype=np.float32)[None, None, :].repeat(4, 0)
    up = fo.scale(row, up=(2, 1), linear=True, in_pad=0)
    np.testing.assert_allclose(up[0, 0], [0, 0.25, 0.75, 1.25, 1.75, 2.25, 2.75, 3.25, 3.75, 4.25, 4.75, 5.25, 5.75, 6.25, 6.75, 7])
    up = fo.scale(row + 1, up=(2, 1), linear=True, in_pad=1)
    assert up[0, 0, 0] == pytest.approx(0.75) and up[0, 0, -1] == pytest.approx(6.0)
    # a shallow tensor blends with nothing but its own plane; a deep one without padding bleeds into the neighbouring tile
    x8 = rng.normal(size=(8, 4, 4)).astype(np.float32)
    sh, dp = fo.scale(x8, up=(2, 2), linear=True, in_pad=0), fo.scale(x8, up=(2, 2), linear=True, in_pad=0, deep=True)
    np.testing.assert_array_equal(sh[:, 1:-1, 1:-1], dp[:, 1:-1, 1:-1])
    assert not np.array_equal(sh[:4, :, -1], dp[:4, :, -1])
    # deep layers of one texel width / height never interpolate (deepscalelayer.cpp:34)
    one = rng.normal(size=(8, 1, 5)).astype(np.float32)
    np.testing.assert_array_equal(fo.scale(one, up=(2, 2), linear=True, deep=True), fo.scale(one, up=(2, 2), deep=True))


def test_concat_and_rgb2bgr():
    rng = np.random.default_rng(3)
    parts = [rng.normal(size=(c, 5, 6)).astype(np.float32) for c in (3, 8, 6)]
    y = fo.concat(parts, act=fo.ACT_RELU)
    assert y.shape == (17, 5, 6)
    np.testing.assert_array_equal(y, np.maximum(np.concatenate(parts), 0))
    x = rng.normal(size=(7, 3, 4)).astype(np.float32)
    y = fo.rgb2bgr(x)
    np.testing.assert_array_equal(y[[0, 1, 2, 3]], x[[2, 1, 0, 3]])
    np.testing.assert_array_equal(y[4], x[6])
    np.testing.assert_array_equal(y[6], x[4])
    np.testing.assert_array_equal(y[5], x[5])
    x5 = x[:5]
    y5 = fo.rgb2bgr(x5)   # channel 4 sits in lane 0 of the second texel: it trades places with the non-existing lane 2
    assert np.all(y5[4] == 0)
