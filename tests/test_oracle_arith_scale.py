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


@pytest.mark.parametrize("deep", [False, True])
def test_dwconv_equals_block_diagonal_convolution(deep):
    """The depthwise restatement against the regular-convolution oracle (which the reference's convolution known-answer
    tests pin, tests/test_oracle_kat.py): a depthwise 3x3 is a 3x3 convolution whose weight matrix is diagonal."""
    rng = np.random.default_rng(5)
    c, h, w = 10, 9, 12
    x = rng.normal(size=(c, h, w)).astype(np.float32)
    bias = rng.uniform(-0.5, 0.5, c).astype(np.float32)
    wk = rng.normal(size=(c, 3, 3)).astype(np.float32)
    bn = np.concatenate([rng.uniform(0.5, 1.5, c), rng.uniform(-0.2, 0.2, c)]).astype(np.float32)
    full = np.zeros((c, 3, 3, c), np.float32)
    for i in range(c):
        full[i, :, :, i] = wk[i]
    for ds, post_bn, pad, act in [(1, False, 1, fo.ACT_NONE), (2, False, 1, fo.ACT_RELU), (1, True, 1, fo.ACT_NONE), (1, False, 0, fo.ACT_NONE)]:
        if deep and pad == 0:
            continue   # un-padded deep tensors bleed into the neighbouring tile; covered below
        wb_dw = np.concatenate([bias, wk.reshape(-1)] + ([bn] if post_bn else []))
        wb_full = np.concatenate([bias, full.reshape(-1)] + ([bn] if post_bn else []))
        ref = fo.conv2d(x, wb_full, c, 3, downsample=ds, in_pad=pad, flags=fo.POST_BATCHNORM if post_bn else 0, deep=deep, act=act)
        got = fo.dwconv3x3(x, wb_dw, downsample=ds, in_pad=pad, deep=deep, post_bn=post_bn, act=act)
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("deep,mult", [(False, 1), (True, 1), (True, 2), (True, 3)])
def test_dwconv_multiplier_and_residual_equal_sparse_convolution(deep, mult):
    """Channel multiplier (deep only: output channel m * C + c = input channel c through W[c][.][.][m], deepdwconvlayerbase.cpp:234-246,
    288-297) and the residual input (conv_dw_3x3.frag:133-140, shaders/deep/residual.inc) against the pinned regular convolution with
    the corresponding sparse weight matrix."""
    rng = np.random.default_rng(9 + mult)
    c, h, w = 8, 9, 11
    co = c * mult
    x = rng.normal(size=(c, h, w)).astype(np.float32)
    bias = rng.uniform(-0.5, 0.5, co).astype(np.float32)
    wk = rng.normal(size=(c, 3, 3, mult)).astype(np.float32)
    bn = np.concatenate([rng.uniform(0.5, 1.5, co), rng.uniform(-0.2, 0.2, co)]).astype(np.float32)
    res = rng.normal(size=(co, h, w)).astype(np.float32)
    full = np.zeros((co, 3, 3, c), np.float32)
    for m in range(mult):
        for i in range(c):
            full[m * c + i, :, :, i] = wk[i, :, :, m]
    for post_bn, relu_res, bn_res in [(False, False, False), (False, True, False), (True, True, False), (True, False, True)]:
        if bn_res and not deep:
            continue                             # the shallow shader has no batch-norm on its residual
        wb_dw = np.concatenate([bias, wk.reshape(-1)] + ([bn] if post_bn else []))
        wb_full = np.concatenate([bias, full.reshape(-1)] + ([bn] if post_bn else []))
        fl = (fo.POST_BATCHNORM if post_bn else 0) | (fo.RELU_ON_RESIDUAL if relu_res else 0) | (fo.BATCHNORM_ON_RESIDUAL if bn_res else 0)
        ref = fo.conv2d(x, wb_full, co, 3, in_pad=1, flags=fl, deep=deep, residual=res)
        got = fo.dwconv3x3(x, wb_dw, in_pad=1, deep=deep, post_bn=post_bn, multiplier=mult, residual=res, relu_on_residual=relu_res, bn_on_residual=bn_res)
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5)
    with pytest.raises(RuntimeError):
        fo.dwconv3x3(x, np.zeros(co * 10 * 2, np.float32), multiplier=2, deep=False)      # shallow layers have no multiplier


def test_dwconv_quirk_and_precision():
    rng = np.random.default_rng(6)
    c, h, w = 6, 5, 7
    x = rng.normal(size=(c, h, w)).astype(np.float32)
    bias = rng.uniform(0.5, 1.5, c).astype(np.float32)
    wk = rng.normal(size=(c, 3, 3)).astype(np.float32)
    bn = np.concatenate([rng.uniform(0.5, 1.5, c), rng.uniform(-0.2, 0.2, c)]).astype(np.float32)
    wb = np.concatenate([bias, wk.reshape(-1), bn])
    plain = fo.dwconv3x3(x, wb[: c * 10], in_pad=1) - bias[:, None, None]
    # reference read position (convlayer_dw_3x3_vanilla.cpp:66): scale = the bias values, beta = the first C weights
    q = fo.dwconv3x3(x, wb, in_pad=1, post_bn=True, quirks=fo.QUIRK_DW_BN_OFFSET)
    s, beta = bias, wk.reshape(-1)[:c]
    np.testing.assert_allclose(q, plain * s[:, None, None] + (bias * s + beta)[:, None, None], rtol=1e-5, atol=1e-5)
    sane = fo.dwconv3x3(x, wb, in_pad=1, post_bn=True)
    np.testing.assert_allclose(sane, plain * bn[:c, None, None] + (bias * bn[:c] + bn[c:])[:, None, None], rtol=1e-5, atol=1e-5)
    # the deep layer ignores the quirk, dilates, and reduces its parameters to fp16 unless storage is fp32
    d1 = fo.dwconv3x3(x, wb, in_pad=2, deep=True, post_bn=True, dilation=2, quirks=fo.QUIRK_DW_BN_OFFSET)
    d2 = fo.dwconv3x3(x, wb, in_pad=2, deep=True, post_bn=True, dilation=2)
    np.testing.assert_array_equal(d1, d2)
    assert not np.allclose(d2, fo.dwconv3x3(x, wb, in_pad=2, deep=True, post_bn=True, dilation=1))
    h16 = fo.dwconv3x3(fo.half_round(x), wb, in_pad=1, deep=True, prec=fo.FP16_STORE)
    np.testing.assert_allclose(h16, fo.dwconv3x3(fo.half_round(x), wb, in_pad=1, deep=True), rtol=4e-3, atol=4e-3)
    with pytest.raises(RuntimeError):
        fo.dwconv3x3(x, wb, dilation=2)   # conv_dw_3x3.frag has no dilation


def test_transconv3x3_is_convolution_of_zero_stuffed_input():
    """The 3x3 stride-2 transpose convolution of the reference (strata by output parity, convtrans3x3_stride2.frag) is the
    pinned regular-convolution oracle applied to the zero-stuffed input (up[2j][2i] = in[j][i]) with zero padding 1."""
    rng = np.random.default_rng(8)
    ci, co, h, w = 6, 5, 7, 9
    x = rng.normal(size=(ci, h, w)).astype(np.float32)
    wb = np.concatenate([rng.uniform(-0.5, 0.5, co), rng.normal(size=co * 9 * ci) * 0.3, rng.uniform(0.5, 1.5, co), rng.uniform(-0.2, 0.2, co)]).astype(np.float32)
    up = np.zeros((ci, 2 * h, 2 * w), np.float32)
    for act in (fo.ACT_NONE, fo.ACT_RELU):
        up[:, ::2, ::2] = np.maximum(x, 0) if act == fo.ACT_RELU else x
        for post_bn in (False, True):
            ref = fo.conv2d(up, wb, co, 3, in_pad=1, flags=fo.POST_BATCHNORM if post_bn else 0)
            got = fo.transconv(x, wb, co, 3, in_pad=1, post_bn=post_bn, act=act)
            np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5)
    # without input padding the texel one past the edge is the clamped edge texel instead of zero
    got0 = fo.transconv(x, wb, co, 3, in_pad=0)
    np.testing.assert_allclose(got0[:, :-1, :-1], fo.transconv(x, wb, co, 3, in_pad=1)[:, :-1, :-1], rtol=1e-6, atol=1e-6)
    assert not np.allclose(got0[:, :, -1], fo.transconv(x, wb, co, 3, in_pad=1)[:, :, -1])


def test_transconv2x2_strata():
    rng = np.random.default_rng(9)
    ci, co, h, w = 5, 4, 6, 7
    x = rng.normal(size=(ci, h, w)).astype(np.float32)
    wk = rng.normal(size=(co, 2, 2, ci)).astype(np.float32)
    bias = rng.uniform(-0.5, 0.5, co).astype(np.float32)
    wb = np.concatenate([bias, wk.reshape(-1)])
    # every output parity class (b, a) is a 1x1 convolution with W[:, b, a, :] ...
    plain = fo.transconv(x, wb, co, 2, in_pad=1, quirks=0)
    for b in (0, 1):
        for a in (0, 1):
            ref = np.einsum("oc,chw->ohw", wk[:, b, a, :], x) + bias[:, None, None]
            np.testing.assert_allclose(plain[:, b::2, a::2], ref, rtol=1e-5, atol=1e-5)
    # ... of input (i, j); the reference's shader reads row j+1 for odd rows and (i+1, j+1) for odd/odd texels
    # (convtrans2x2_stride2.frag STEP 3 / 4: tc + texStep), with zero padding past the last row / column
    q = fo.transconv(x, wb, co, 2, in_pad=1)
    xp = np.pad(x, ((0, 0), (0, 1), (0, 1)))
    np.testing.assert_allclose(q[:, 0::2, :], plain[:, 0::2, :], rtol=0, atol=0)
    np.testing.assert_allclose(q[:, 1::2, 0::2], np.einsum("oc,chw->ohw", wk[:, 1, 0, :], xp[:, 1:, :-1]) + bias[:, None, None], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(q[:, 1::2, 1::2], np.einsum("oc,chw->ohw", wk[:, 1, 1, :], xp[:, 1:, 1:]) + bias[:, None, None], rtol=1e-5, atol=1e-5)
    with pytest.raises(RuntimeError):
        fo.transconv(x, np.zeros(co * (1 + 16 * ci), np.float32), co, 4)


@pytest.mark.parametrize("kernel", [2, 3])
def test_deep_transconv_is_the_full_convolution_of_the_zero_stuffed_input(kernel):
    """deep::DeepTransConvLayer2x2 / 3x3 (deeptransconv{2x2,3x3}_stride2.{vert,frag}): out[o] = sum_k W[k] u[o - k] with u the
    zero-stuffed input -- checked against the pinned regular-convolution oracle run with the flipped kernel on u shifted by one
    texel (zero padding supplies the texels outside the image, which the deep shader masks to zero)."""
    rng = np.random.default_rng(41 + kernel)
    ci, co, h, w = 6, 5, 5, 7
    x = rng.normal(size=(ci, h, w)).astype(np.float32)
    bias = rng.uniform(-0.5, 0.5, co).astype(np.float32)
    wk = rng.normal(size=(co, kernel, kernel, ci)).astype(np.float32)
    wb = np.concatenate([bias, wk.reshape(-1)])
    got = fo.transconv(x, wb, co, kernel, deep=True)
    # 3x3 kernel holding the (2x2 or 3x3) taps, flipped; zero-stuffed input with one leading zero row / column
    k3 = np.zeros((co, 3, 3, ci), np.float32)
    k3[:, :kernel, :kernel] = wk
    flipped = k3[:, ::-1, ::-1].copy()
    u = np.zeros((ci, 2 * h + 1, 2 * w + 1), np.float32)
    u[:, 1::2, 1::2] = x
    ref = fo.conv2d(u, np.concatenate([bias, flipped.reshape(-1)]), co, 3, in_pad=1)[:, :2 * h, :2 * w]
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5)
    # prefix activation applies to the fetched (and masked) texel; fp16 storage truncates the weights
    relu = fo.transconv(x, wb, co, kernel, deep=True, act=fo.ACT_RELU)
    np.testing.assert_allclose(relu, fo.transconv(np.maximum(x, 0), wb, co, kernel, deep=True), rtol=1e-6, atol=1e-6)
    h16 = fo.transconv(fo.half_round(x), wb, co, kernel, deep=True, prec=fo.FP16_STORE)
    np.testing.assert_allclose(h16, got, rtol=2e-2, atol=2e-2)
    assert not np.array_equal(h16, got)
