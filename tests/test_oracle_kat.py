"""Pins the CPU oracle against the reference's own layer-level known-answer tests.

Each test re-expresses a test of /root/reference/unit_tests (cited per test) against
oracle/fyn_oracle.c.  The reference checks with ASSERT_NEAR(.., 1e-3) on small-integer data; the
oracle must meet the same bound in FP32 mode and in the fp16 modes (small integers are exact in
fp16, which is why the reference's tolerance works with RGBA16F render targets).
"""
import numpy as np
import pytest

import fyn_oracle as fo

PRECS = [fo.FP32, fo.FP16_STORE, fo.FP16_BLEND]


def stack_convolution(bias, kernel2d, cin, cout):
    """unit_tests/layertestbase.cpp:45-60: [bias | O x Ky x Kx x I] with the same 2-D kernel everywhere."""
    k = np.asarray(kernel2d, np.float32)
    w = np.broadcast_to(k[None, :, :, None], (cout, k.shape[0], k.shape[1], cin))
    return np.concatenate([np.full(cout, bias, np.float32), w.reshape(-1).astype(np.float32)])


def antisym_kernel(k):
    """-1 before the centre, 0 at the centre, +1 after (convlayertests.cpp:199-204)."""
    v = np.zeros(k * k, np.float32)
    mid = (k * k - 1) // 2
    v[:mid] = -1
    v[mid + 1:] = 1
    return v.reshape(k, k)


def padded_convolution(x_padded, wb, cout, k, cin, down=1, pre_relu=False):
    """unit_tests/convlayertests.cpp:83-118: valid conv on a pre-padded CHW tensor, own numpy statement."""
    c, h, w = x_padded.shape
    assert c == cin
    pad = (k - 1) // 2
    bias, wt = wb[:cout], wb[cout:cout + cout * k * k * cin].reshape(cout, k, k, cin)
    x = np.maximum(x_padded, 0) if pre_relu else x_padded
    ys = range(pad, h - pad, down)
    xs = range(pad, w - pad, down)
    oh, ow = (h - 2 * pad) // down, (w - 2 * pad) // down
    out = np.zeros((cout, oh, ow), np.float64)
    for ky in range(k):
        for kx in range(k):
            patch = x[:, ky:ky + (h - 2 * pad):down, kx:kx + (w - 2 * pad):down][:, :oh, :ow]
            out += np.einsum("oc,chw->ohw", wt[:, ky, kx, :].astype(np.float64), patch.astype(np.float64))
    assert len(ys) >= oh and len(xs) >= ow
    return (out + bias[:, None, None]).astype(np.float32)


GRID_1x1 = [(64, 64, 4, 4), (64, 80, 4, 4), (128, 80, 4, 8), (56, 56, 64, 64), (128, 80, 16, 8), (256, 128, 12, 4)]
GRID_NXN = [(64, 64, 4, 4), (64, 80, 4, 4), (128, 80, 4, 8), (128, 80, 16, 8), (256, 128, 12, 8)]


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("ds", [1, 2])
@pytest.mark.parametrize("w,h,ci,co", GRID_1x1)
def test_shallow_conv1x1_all_ones(w, h, ci, co, ds, prec):
    """convlayertests.cpp:159-183 (+ grid :426-447): all-ones input and weights -> every output == inchans."""
    x = np.ones((ci, h, w), np.float32)
    wb = stack_convolution(0.0, [[1.0]], ci, co)
    y = fo.conv2d(x, wb, co, 1, downsample=ds, prec=prec)
    assert y.shape == (co, h // ds, w // ds)
    np.testing.assert_allclose(y, ci, atol=1e-3)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("w,h,ci,co", GRID_1x1)
def test_deep_conv1x1_random(w, h, ci, co, prec):
    """convlayertests.cpp:185-225: deep 1x1 on random data in [-10,10] vs paddedConvolution.
    (The reference builds an antisymmetric 1x1 kernel = [0]; we use a non-trivial +1 kernel too.)"""
    rng = np.random.default_rng(w * 131 + h * 7 + ci)
    x = np.round(rng.uniform(-10, 10, (ci, h, w))).astype(np.float32)   # integers: exact in fp16
    for kern in ([[0.0]], [[1.0]]):
        wb = stack_convolution(0.0, kern, ci, co)
        ref = padded_convolution(x, wb, co, 1, ci)
        y = fo.conv2d(x, wb, co, 1, deep=True, prec=prec)
        np.testing.assert_allclose(y, ref, atol=1e-3)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("ds", [1, 2])
@pytest.mark.parametrize("k", [3, 5, 7, 9])
@pytest.mark.parametrize("w,h,ci,co", GRID_NXN[:3])
def test_shallow_convNxN_clamp_to_edge(w, h, ci, co, k, ds, prec):
    """convlayertests.cpp:228-262: constant 1.0 input, antisymmetric kernel, NO padding -> output == 0
    everywhere, which only holds because un-padded shallow convs clamp to the edge (quirk Q3).
    9x9 is not in the reference grid; it follows the same rule and StyleNet depends on it."""
    x = np.ones((ci, h, w), np.float32)
    wb = stack_convolution(0.0, antisym_kernel(k), ci, co)
    y = fo.conv2d(x, wb, co, k, downsample=ds, prec=prec)
    assert y.shape == (co, h // ds, w // ds)
    np.testing.assert_allclose(y, 0.0, atol=1e-3)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("ds", [1, 2])
@pytest.mark.parametrize("k", [3, 5, 7])
@pytest.mark.parametrize("w,h,ci,co", GRID_NXN)
def test_deep_convNxN_zero_padded(w, h, ci, co, k, ds, prec):
    """convlayertests.cpp:265-305 (+ :341-420 fixed 5x5 64x64x4->4 and 3x3 256x128x12->8 s2):
    deep conv, inputPadding=(k-1)/2 real zeros, antisymmetric kernel, vs paddedConvolution."""
    pad = (k - 1) // 2
    x = np.ones((ci, h, w), np.float32)
    wb = stack_convolution(0.0, antisym_kernel(k), ci, co)
    ref = padded_convolution(np.pad(x, ((0, 0), (pad, pad), (pad, pad))), wb, co, k, ci, down=ds)
    y = fo.conv2d(x, wb, co, k, downsample=ds, in_pad=pad, deep=True, prec=prec)
    assert y.shape == ref.shape
    np.testing.assert_allclose(y, ref, atol=1e-3)


@pytest.mark.parametrize("prec", PRECS)
def test_deep_conv_random_integer_prerelu(prec):
    """paddedConvolution's preReLU branch (convlayertests.cpp:103-105) on random integer data."""
    rng = np.random.default_rng(5)
    ci, co, h, w, k = 12, 8, 20, 24, 3
    x = np.round(rng.uniform(-4, 4, (ci, h, w))).astype(np.float32)
    wt = np.round(rng.uniform(-2, 2, (co, k, k, ci))).astype(np.float32)
    wb = np.concatenate([np.round(rng.uniform(-3, 3, co)).astype(np.float32), wt.reshape(-1)])
    ref = padded_convolution(np.pad(x, ((0, 0), (1, 1), (1, 1))), wb, co, k, ci, pre_relu=True)
    y = fo.conv2d(x, wb, co, k, in_pad=1, deep=True, act=fo.ACT_RELU, prec=prec)
    np.testing.assert_allclose(y, ref, atol=1e-3)


def test_network_upload_conv_download_zero():
    """networktests.cpp:60-206: upload -> conv3x3 (4->8, antisymmetric +-1 filter) -> download on an
    all-ones 32x32 input must be exactly 0."""
    img = np.ones((32, 32, 4), np.float32)
    x = fo.upload_hwc(img)
    wb = stack_convolution(0.0, antisym_kernel(3), 4, 8)
    y = fo.conv2d(x, wb, 8, 3)
    host = fo.download_shallow(y)
    assert host.shape == (2, 32, 32, 4)
    assert np.all(host == 0.0)


POOL_GRID = [(8, 8, 4), (200, 200, 4), (80, 40, 12), (50, 50, 23), (40, 40, 80)]
GLOBAL_GRID = [(80, 40, 56), (100, 80, 12), (8, 8, 8), (200, 200, 4), (50, 50, 23), (2, 2, 24), (8, 4, 24), (40, 40, 80)]


@pytest.mark.parametrize("is_max", [True, False])
@pytest.mark.parametrize("w,h,c", POOL_GRID)
def test_pool_2x2(w, h, c, is_max):
    """pooltests.cpp:174-260 with the CPU references :65-117 (pool == stride == 2)."""
    rng = np.random.default_rng(w + h + c)
    x = rng.uniform(-10, 10, (c, h, w)).astype(np.float32)
    y = fo.pool2d(x, pool=2, downsample=2, is_max=is_max)
    blocks = x[:, :h // 2 * 2, :w // 2 * 2].reshape(c, h // 2, 2, w // 2, 2)
    ref = blocks.max(axis=(2, 4)) if is_max else blocks.mean(axis=(2, 4))
    np.testing.assert_allclose(y, ref, atol=1e-4)


@pytest.mark.parametrize("is_max", [True, False])
@pytest.mark.parametrize("w,h,c", GLOBAL_GRID)
def test_pool_global(w, h, c, is_max):
    """pooltests.cpp:262-316 (global avg / max)."""
    rng = np.random.default_rng(3 * w + h + c)
    x = rng.uniform(-10, 10, (c, h, w)).astype(np.float32)
    y = fo.pool2d(x, is_max=is_max, global_=True)
    ref = x.max(axis=(1, 2)) if is_max else x.mean(axis=(1, 2), dtype=np.float64)
    np.testing.assert_allclose(y.reshape(-1), ref, atol=1e-3)


BN_GRID = [(4, 4, 36), (80, 40, 52), (4, 4, 4), (256, 128, 64), (120, 80, 3), (200, 200, 4), (50, 50, 31), (12, 12, 128)]


@pytest.mark.parametrize("deep", [False, True])
@pytest.mark.parametrize("w,h,c", BN_GRID)
def test_batchnorm(w, h, c, deep):
    """misctests.cpp:186-236: x*s+b on random [-10,10] data (reference tolerance 1e-1 with fp16 targets)."""
    rng = np.random.default_rng(w * h + c)
    x = rng.uniform(-10, 10, (c, h, w)).astype(np.float32)
    sb = rng.uniform(-2, 2, 2 * c).astype(np.float32)
    ref = x * sb[:c, None, None] + sb[c:, None, None]
    np.testing.assert_allclose(fo.batchnorm(x, sb, deep=deep), ref, atol=1e-5)
    np.testing.assert_allclose(fo.batchnorm(x, sb, deep=deep, prec=fo.FP16_STORE), ref, atol=1e-1)


@pytest.mark.parametrize("prec", PRECS)
def test_deep_gemm_alternating(prec):
    """misctests.cpp:238-265: 512 -> 256, weights +1/-1 alternating along the input axis, all-ones input -> 0."""
    ci, co = 512, 256
    w = np.tile(np.array([1.0, -1.0], np.float32), ci // 2)
    wb = np.concatenate([np.zeros(co, np.float32), np.tile(w, co)])
    y = fo.conv2d(np.ones((ci, 1, 1), np.float32), wb, co, 1, deep=True, prec=prec)
    np.testing.assert_allclose(y.reshape(-1), 0.0, atol=1e-3)


# ------------------------------------------------------------------------------------------------
# layouts and conversions
# ------------------------------------------------------------------------------------------------

def test_deep_tiling_rule():
    """cpu/cpubuffershape.cpp:430-447 and the examples of SURVEY a11."""
    assert fo.deep_tiling(64) == (4, 4)
    assert fo.deep_tiling(1000) == (18, 14)
    assert fo.deep_tiling(3) == (1, 1)
    assert fo.deep_tiling(2048)[0] * fo.deep_tiling(2048)[1] >= 512
    assert fo.deep_texture_size(64, 112, 112, 1) == (453, 453)
    for c in range(1, 300):
        tx, ty = fo.deep_tiling(c)
        assert tx >= ty >= 1 and tx * ty >= (c + 3) // 4


@pytest.mark.parametrize("c,h,w,pad", [(3, 5, 7, 0), (4, 8, 8, 1), (23, 6, 9, 1), (64, 7, 7, 2), (10, 3, 4, 0)])
def test_layout_roundtrip(c, h, w, pad):
    """pack/unpack of both layouts (unit_tests/layertestbase.cpp:235-317) are inverse, zero elsewhere."""
    rng = np.random.default_rng(c)
    x = rng.normal(size=(c, h, w)).astype(np.float32)
    d = fo.pack_deep(x, pad)
    np.testing.assert_array_equal(fo.unpack_deep(d, c, h, w, pad), x)
    assert np.isclose(np.abs(d).sum(dtype=np.float64), np.abs(x).sum(dtype=np.float64), rtol=1e-6)
    s = fo.pack_shallow(x, pad)
    assert s.shape == ((c + 3) // 4, h + 2 * pad, w + 2 * pad, 4)
    np.testing.assert_array_equal(fo.unpack_shallow(s, c, h, w, pad), x)
    if pad:
        assert np.all(s[:, 0] == 0) and np.all(s[:, :, 0] == 0)
    tx, ty = fo.deep_tiling(c)
    # tile 1 starts one padding gap after tile 0 (single shared gap, deeptiler.cpp:91-94)
    if c > 4 and tx > 1:
        np.testing.assert_array_equal(d[pad:pad + h, pad + w + pad:pad + 2 * w + pad, 0], x[4])


def test_half_truncation_matches_table_scheme():
    """gpu/floatconversion.cpp:44-58: truncation, never rounds up; numpy and C versions agree."""
    rng = np.random.default_rng(0)
    v = np.concatenate([rng.normal(size=4000) * 10.0 ** rng.integers(-9, 5, 4000),
                        [0.0, -0.0, 1.0, 65504.0, 1e6, -1e6, 6e-8, 5.9e-8, 2.0 ** -14, 2.0 ** -24]]).astype(np.float32)
    t = fo.half_trunc(v)
    c = np.array([fo.lib().fyo_half_trunc(float(x)) for x in v], np.float32)
    np.testing.assert_array_equal(t, c)
    fin = np.isfinite(t)
    assert np.all(np.abs(t[fin]) <= np.abs(v[fin]))            # truncation toward zero
    rn = fo.half_round(v)
    assert np.all(np.abs(t[fin] - v[fin]) <= 2 * np.abs(rn[fin] - v[fin]) + 2.0 ** -24 + 1e-30 + np.abs(v[fin]) * 2.0 ** -10)


def test_fractional_conv_sampling_pattern():
    """fraconv3x3.frag:13-19 + convlayerbase_vanilla.cpp:362-371: with Q1/Q2 on, deconv1-style
    (s=0.5, ds=2) 3x3 taps read source columns (x-1, x-1, x) and rows (y-1, y, y);
    with the quirks off the taps are symmetric: columns (x-1, x, x)."""
    h, w = 6, 8
    x = np.arange(h * w, dtype=np.float32).reshape(1, h, w)
    for kx in range(3):
        for ky in range(3):
            wt = np.zeros((1, 3, 3, 1), np.float32)
            wt[0, ky, kx, 0] = 1.0
            wb = np.concatenate([[0.0], wt.reshape(-1)]).astype(np.float32)
            y = fo.conv2d(x, wb, 1, 3, downsample=2, source_step=0.5, fractional=True)
            assert y.shape == (1, h, w)
            dx = [-1, -1, 0][kx]
            dy = [-1, 0, 0][ky]
            ys = np.clip(np.arange(h) + dy, 0, h - 1)
            xs = np.clip(np.arange(w) + dx, 0, w - 1)
            np.testing.assert_array_equal(y[0], x[0][np.ix_(ys, xs)])
            y2 = fo.conv2d(x, wb, 1, 3, downsample=2, source_step=0.5, fractional=True, quirks=0)
            dx2 = [-1, 0, 0][kx]
            xs2 = np.clip(np.arange(w) + dx2, 0, w - 1)
            np.testing.assert_array_equal(y2[0], x[0][np.ix_(ys, xs2)])


def test_fractional_conv_upsamples():
    """fractionalconvlayerNxN_vanilla.cpp:46-49: output = floor(W/(s*ds)); deconv2 (s=.25, ds=2) doubles,
    deconv3 (s=.5) doubles; Q2: only the first horizontal tap sees the activation."""
    x = -np.ones((4, 4, 6), np.float32)
    wb = stack_convolution(0.0, np.ones((3, 3)), 4, 4)
    y = fo.conv2d(x, wb, 4, 3, downsample=2, source_step=0.25, fractional=True, act=fo.ACT_RELU)
    assert y.shape == (4, 8, 12)
    # first tap of each row is ReLU'd to 0, the other two taps pass -1 through: 3 rows * 2 taps * 4 ch * -1
    np.testing.assert_allclose(y, -24.0)
    y = fo.conv2d(x, wb, 4, 3, downsample=2, source_step=0.25, fractional=True, act=fo.ACT_RELU, quirks=0)
    np.testing.assert_allclose(y, 0.0)
    y = fo.conv2d(x, stack_convolution(0.0, np.ones((9, 9)), 4, 3), 3, 9, source_step=0.5, fractional=True)
    assert y.shape == (3, 8, 12)


def test_residual_and_postbn_fold():
    """conv.inc:1-7 + convweightarrayKxKxNxM.cpp:167-181 + residual.inc: out = s*(W*x) + (b*s+beta) + s'*act_r(res)."""
    rng = np.random.default_rng(11)
    ci, co, h, w = 8, 8, 5, 6
    x = rng.normal(size=(ci, h, w)).astype(np.float32)
    res = rng.normal(size=(co, h, w)).astype(np.float32)
    wt = rng.normal(size=(co, 1, 1, ci)).astype(np.float32)
    b = rng.normal(size=co).astype(np.float32)
    s = rng.uniform(0.5, 1.5, co).astype(np.float32)
    beta = rng.normal(size=co).astype(np.float32)
    wb = np.concatenate([b, wt.reshape(-1), s, beta])
    lin = np.einsum("oc,chw->ohw", wt[:, 0, 0, :], x)
    for deep in (False, True):
        y = fo.conv2d(x, wb, co, 1, flags=fo.POST_BATCHNORM, residual=res, deep=deep)
        np.testing.assert_allclose(y, lin * s[:, None, None] + (b * s + beta)[:, None, None] + res, atol=1e-5)
        y = fo.conv2d(x, wb, co, 1, flags=fo.POST_BATCHNORM | fo.BATCHNORM_ON_RESIDUAL | fo.RELU_ON_RESIDUAL,
                      residual=res, deep=deep)
        np.testing.assert_allclose(y, lin * s[:, None, None] + (b * s + beta)[:, None, None]
                                   + np.maximum(res, 0) * s[:, None, None], atol=1e-5)


def test_stem_7x7_underpadded_equals_zero_pad3():
    """the ResNet stem (7x7 s2, inputPadding 1 < 3, single tile) clamps onto the zero ring,
    i.e. behaves like zero padding 3 (== torch conv2d k7 s2 p3)."""
    rng = np.random.default_rng(2)
    x = rng.normal(size=(3, 16, 16)).astype(np.float32)
    wt = rng.normal(size=(4, 7, 7, 3)).astype(np.float32)
    wb = np.concatenate([np.zeros(4, np.float32), wt.reshape(-1)])
    y = fo.conv2d(x, wb, 4, 7, downsample=2, in_pad=1, deep=True)
    xp = np.pad(x, ((0, 0), (3, 3), (3, 3)))
    ref = padded_convolution(xp, wb, 4, 7, 3, down=2)
    np.testing.assert_allclose(y, ref[:, :8, :8], atol=1e-4)


def test_maxpool_3x3_s2_pad1():
    """deepmaxpool.frag:12-61 + resnet50.cpp:215-217: window [2o-1, 2o+1] with zero padding after ReLU."""
    rng = np.random.default_rng(4)
    x = rng.normal(size=(5, 12, 12)).astype(np.float32)
    y = fo.pool2d(x, pool=3, downsample=2, in_pad=1, is_max=True, act=fo.ACT_RELU)
    xr = np.pad(np.maximum(x, 0), ((0, 0), (1, 1), (1, 1)))
    ref = np.stack([[[xr[c, 2 * i:2 * i + 3, 2 * j:2 * j + 3].max() for j in range(6)] for i in range(6)] for c in range(5)])
    np.testing.assert_allclose(y, ref, atol=0)


def test_weight_file_layouts_match_reference_tables():
    """stylenet9x9.cpp:41-56, stylenet3x3.cpp:41-50, resnet50.cpp:539-677 (hard-coded offsets & sizes)."""
    o9 = fo.stylenet_offsets(9)
    assert (o9["conv1"], o9["conv2"], o9["conv3"], o9["deconv1"], o9["deconv2"], o9["deconv3"]) == (0, 2928, 5108, 12348, 19568, 21740)
    assert [o9[f"res{r}_{i}"] for r in range(1, 6) for i in (1, 2)] == [24659 + 14440 * j for j in range(10)]
    assert o9["_total"] == 169059
    o3 = fo.stylenet_offsets(3)
    assert (o3["conv1"], o3["conv2"], o3["conv3"], o3["deconv1"], o3["deconv2"], o3["deconv3"]) == (0, 336, 2516, 9756, 16976, 19148)
    assert (o3["res1_1"], o3["res1_2"], o3["res2_1"], o3["res2_2"], o3["_total"]) == (19475, 33915, 48355, 62795, 77235)
    r = fo.resnet50_offsets()
    ref_bytes = {2: 0, 3: 24, 5: 38424, 6: 38936, 8: 56088, 9: 204312, 7: 270872, 10: 337432, 17: 837144,
                 18: 905752, 20: 1038360, 21: 1629720, 19: 1893912, 33: 5526040, 34: 5794328, 36: 6321688,
                 37: 8684056, 35: 9736728, 57: 33159704, 58: 34220568, 60: 36323864, 61: 45767192,
                 59: 49969688, 62: 58366488, 69: 89889304, 72: 94108184}
    for n, b in ref_bytes.items():
        assert r[n] * 4 == b, n
    assert r["_total"] * 4 == 102304184
    assert 1 not in r and 71 not in r
