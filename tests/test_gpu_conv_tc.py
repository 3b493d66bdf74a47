"""GPU parity tests of the tcgen05 / TMEM convolution family (fyn_conv_tc.cu) through the C ABI.

The tensor-core kernels multiply fp16 activations with fp16-ROUNDED weights and accumulate in fp32, so the exact
model is the fp16-store oracle run with half-rounded weights (and a half-rounded input where the kernel converts
an fp32 upload texture): the kernel must match that within 1 fp16 ulp (+ fp32 summation noise).  Against the
true-fp32-weight oracle the bound is rel-L2 <= 2e-3 (weight rounding 2^-11 per product, random signs), and the
tcgen05 result must agree with the in-library direct (CUDA-core) kernel to the same tolerance.
"""
import numpy as np
import pytest

import fyn_oracle as fo
from fyusenet_b200 import capi
import gpu_util
from gpu_util import assert_close_f16, conv_gpu, ctx, half, random_wb, rel_l2

pytestmark = pytest.mark.gpu


def _round_weights(wb, co, nweights=None):
    """conv weights are fp16 operands of the MMA; bias and post-BN scale / bias stay fp32 in the epilogue"""
    w = np.array(wb, np.float32, copy=True)
    end = len(w) if nweights is None else co + nweights
    w[co:end] = half(w[co:end])
    return w


def _run(x, wb, co, k, ds=1, relu=True, residual=None, relu_res=False, in_pad=0, out_pad=0, res_pad=0, post_bn=False,
         bn_res=False):
    fl = (capi.FLAG_PRE_RELU if relu else 0) | (capi.FLAG_RELU_ON_RESIDUAL if relu_res else 0) | \
         (capi.FLAG_POST_BATCHNORM if post_bn else 0) | (capi.FLAG_BATCHNORM_ON_RESIDUAL if bn_res else 0)
    y, be, _ = conv_gpu(x, wb, out_channels=co, kernel=k, downsample=ds, flags=fl, residual=residual, in_pad=in_pad,
                        out_pad=out_pad, res_pad=res_pad, backend=capi.BACKEND_TC, want_op=True)
    assert be == 2, "tcgen05 family must have been selected"
    yd = conv_gpu(x, wb, out_channels=co, kernel=k, downsample=ds, flags=fl, residual=residual, in_pad=in_pad,
                  out_pad=out_pad, res_pad=res_pad, backend=capi.BACKEND_DIRECT)
    ofl = (fo.RELU_ON_RESIDUAL if relu_res else 0) | (fo.POST_BATCHNORM if post_bn else 0) | (fo.BATCHNORM_ON_RESIDUAL if bn_res else 0)
    okw = dict(downsample=ds, act=fo.ACT_RELU if relu else fo.ACT_NONE, flags=ofl, in_pad=in_pad, out_pad=out_pad)

    def orc(xx, ww, prec):
        xx = np.asarray(xx, np.float32)
        if xx.ndim == 4:
            return np.stack([fo.conv2d(xx[i], ww, co, k, residual=None if residual is None else half(residual[i]), prec=prec, **okw)
                             for i in range(xx.shape[0])])
        return fo.conv2d(xx, ww, co, k, residual=None if residual is None else half(residual), prec=prec, **okw)

    xs = half(x)
    ci = np.asarray(x).shape[-3]
    exact = orc(xs, _round_weights(wb, co, k * k * ci * co), fo.FP16_STORE)
    true32 = orc(xs, wb, fo.FP32)
    assert_close_f16(y, exact, true32, ulps=1.01, extra_abs=4e-5)
    assert rel_l2(y, yd) <= 2e-3
    return y


@pytest.mark.parametrize("w,h", [(381, 29), (128, 16), (130, 7), (64, 48), (257, 11)])
@pytest.mark.parametrize("variant", ["res_1", "res_2_relu", "res_2_plain", "noact"])
def test_tc_res_layers(w, h, variant):
    """StyleNet residual convs: 3x3, 40 -> 40, clamp-to-edge, pre-ReLU, residual (+ReLU) -- stylenet9x9.cpp:145-186.
    Widths that are not multiples of the 128-pixel tile and heights that do not divide into strips."""
    rng = np.random.default_rng(w * 7 + h)
    x = rng.normal(size=(40, h, w)).astype(np.float32)
    wb = random_wb(rng, 40, 40, 3)
    res = rng.normal(size=(40, h, w)).astype(np.float32) if variant.startswith("res_2") else None
    _run(x, wb, 40, 3, relu=(variant != "noact"), residual=res, relu_res=(variant == "res_2_relu"))


@pytest.mark.parametrize("w,h", [(384, 40), (200, 23), (1524, 12)])
def test_tc_conv1_9x9_from_upload_texture(w, h):
    """conv1: 9x9, 3 -> 12 reading the RGB32F upload texture (pixel-pair K-chunks)."""
    rng = np.random.default_rng(w + h)
    img = rng.random((h, w, 3), dtype=np.float32)
    wb = random_wb(rng, 3, 12, 9)
    c = ctx()
    ys = {}
    for be in (capi.BACKEND_TC, capi.BACKEND_DIRECT):
        op = capi.Conv2d(c, wb, width=w, height=h, in_channels=3, out_channels=12, kernel=9, flags=capi.FLAG_PRE_RELU, backend=be)
        tin = c.tensor(w, h, 3, 0, capi.ORDER_SHALLOW, capi.F32, 1, packing=3)
        tout = c.tensor(w, h, 12, 0, capi.ORDER_SHALLOW, capi.F16)
        tin.upload(img)
        op.run(tin, tout)
        ys[be] = tout.read_chw()
        assert op.backend == be
        for o in (tin, tout, op):
            o.destroy()
    x = fo.upload_hwc(img)
    exact = fo.conv2d(half(x), _round_weights(wb, 12), 12, 9, act=fo.ACT_RELU, prec=fo.FP16_STORE)
    true32 = fo.conv2d(x, wb, 12, 9, act=fo.ACT_RELU)
    assert_close_f16(ys[capi.BACKEND_TC], exact, true32, ulps=1.01, extra_abs=4e-5)
    assert rel_l2(ys[capi.BACKEND_TC], ys[capi.BACKEND_DIRECT]) <= 2e-3


@pytest.mark.parametrize("k,ci,co,w,h", [(3, 12, 20, 260, 24), (3, 20, 40, 380, 18), (3, 12, 20, 1524, 8), (5, 8, 16, 140, 20), (3, 16, 64, 96, 12)])
def test_tc_stride2(k, ci, co, w, h):
    """conv2 / conv3: 3x3 stride 2 (stylenet9x9.cpp:139-143) -- even/odd column de-interleave in the ring slots."""
    rng = np.random.default_rng(k + ci + co + w)
    x = rng.normal(size=(ci, h, w)).astype(np.float32)
    _run(x, random_wb(rng, ci, co, k), co, k, ds=2)


def test_tc_general_shapes():
    """Other shapes the family accepts: 3x3 5/7-plane inputs, 5x5 / 7x7 kernels, padding, post-BN, BN on residual, batch."""
    rng = np.random.default_rng(99)
    x = rng.normal(size=(2, 24, 13, 150)).astype(np.float32)
    res = rng.normal(size=(2, 32, 13, 150)).astype(np.float32)
    _run(x, random_wb(rng, 24, 32, 3, post_bn=True), 32, 3, residual=res, post_bn=True, bn_res=True, in_pad=1, out_pad=2, res_pad=1)
    x = rng.normal(size=(12, 20, 70)).astype(np.float32)
    _run(x, random_wb(rng, 12, 8, 5), 8, 5)
    x = rng.normal(size=(8, 20, 70)).astype(np.float32)
    _run(x, random_wb(rng, 8, 12, 7), 12, 7, relu=False)
    x = rng.normal(size=(3, 33, 140)).astype(np.float32)   # fp16 RGBA plane input, 3x3 pixel-pair mode
    _run(x, random_wb(rng, 3, 12, 3), 12, 3)


def test_tc_weight_hot_swap_and_launch_count():
    c = ctx()
    rng = np.random.default_rng(4)
    x = rng.normal(size=(40, 20, 140)).astype(np.float32)
    w1, w2 = random_wb(rng, 40, 40, 3), random_wb(rng, 40, 40, 3)
    op = capi.Conv2d(c, w1, width=140, height=20, in_channels=40, out_channels=40, kernel=3, flags=capi.FLAG_PRE_RELU)
    assert op.backend == 2
    tin, tout = c.tensor(140, 20, 40), c.tensor(140, 20, 40)
    tin.write_chw(x)
    n0 = c.launch_count()
    op.run(tin, tout)
    assert c.launch_count() == n0 + 1
    y1 = tout.read_chw()
    op.load_weights(w2)
    op.run(tin, tout)
    y2 = tout.read_chw()
    for y, w in ((y1, w1), (y2, w2)):
        exact = fo.conv2d(half(x), _round_weights(w, 40), 40, 3, act=fo.ACT_RELU, prec=fo.FP16_STORE)
        assert_close_f16(y, exact, ulps=1.01, extra_abs=4e-5)


@pytest.mark.parametrize("quirks", [capi.QUIRKS_REFERENCE, 0])
@pytest.mark.parametrize("k,ci,co,step,ds,relu,w,h", [
    (3, 40, 20, 0.5, 2, False, 381, 24),    # deconv1: same resolution, taps collapse onto a 2x2 support
    (3, 20, 12, 0.25, 2, True, 381, 20),    # deconv2: 2x upsample, activation on the first tap only (Q2)
    (9, 12, 3, 0.5, 1, True, 200, 18),      # deconv3: 9x9, 2x upsample
    (3, 16, 16, 0.5, 1, True, 130, 9),
    (5, 8, 24, 0.5, 2, False, 64, 33),
])
def test_tc_fractional(k, ci, co, step, ds, relu, w, h, quirks):
    """StyleNet deconv1..3 (stylenet9x9.cpp:192-202) on the tcgen05 family via phase decomposition, with the
    reference quirks on (default) and off.  Merged taps are summed in fp32 and rounded to fp16 once, so the
    comparison is against the fp32-weight oracle: rel-L2 <= 2e-3, element-wise <= 2 fp16 ulp + 2e-3."""
    rng = np.random.default_rng(k * 31 + ci + co + w)
    x = rng.normal(size=(ci, h, w)).astype(np.float32)
    wb = random_wb(rng, ci, co, k)
    fl = capi.FLAG_PRE_RELU if relu else 0
    kw = dict(out_channels=co, kernel=k, downsample=ds, source_step=step, fractional=True, flags=fl, quirks=quirks)
    y, be, _ = conv_gpu(x, wb, backend=capi.BACKEND_TC, want_op=True, **kw)
    assert be == 2
    yd = conv_gpu(x, wb, backend=capi.BACKEND_DIRECT, **kw)
    okw = dict(downsample=ds, source_step=step, fractional=True, act=fo.ACT_RELU if relu else fo.ACT_NONE, quirks=quirks)
    ref_s = fo.conv2d(half(x), wb, co, k, prec=fo.FP16_STORE, **okw)
    ref_f = fo.conv2d(half(x), wb, co, k, prec=fo.FP32, **okw)
    assert y.shape == ref_s.shape
    assert_close_f16(y, ref_s, ref_f, ulps=2.0, extra_abs=2e-3)
    assert rel_l2(y, yd) <= 2e-3


@pytest.mark.parametrize("case", ["stride2", "frac_stacked", "frac9_stacked"])
def test_tc_long_strips_repeated(case):
    """Many jobs per strip, launched back to back (programmatic dependent launch), against the direct kernel.  Regression
    for a pipeline deadlock: with two MMA-issuing warps a warp released ring rows it had not waited for yet whenever
    the window is shorter than two row advances (stride-2 3x3 convs, stacked fractional 3x3), so a slot could be
    refilled -- and its barrier pass a second phase -- before that warp's parity wait (found at 4096x4096)."""
    c = ctx()
    rng = np.random.default_rng(5)
    if case == "stride2":
        ci, co, k, w, h, kw = 12, 20, 3, 640, 1024, dict(downsample=2)
    elif case == "frac_stacked":
        ci, co, k, w, h, kw = 40, 20, 3, 384, 512, dict(downsample=2, source_step=0.5, fractional=True)
    else:
        ci, co, k, w, h, kw = 12, 3, 9, 384, 512, dict(source_step=0.5, fractional=True)
    x = rng.normal(size=(ci, h, w)).astype(np.float32)
    wb = random_wb(rng, ci, co, k)
    outs = {}
    for backend in (capi.BACKEND_TC, capi.BACKEND_DIRECT):
        op = capi.Conv2d(c, wb, width=w, height=h, in_channels=ci, out_channels=co, kernel=k, flags=capi.FLAG_PRE_RELU, backend=backend, **kw)
        tin = c.tensor(w, h, ci)
        tout = c.tensor(op.out_width, op.out_height, co)
        tin.write_chw(x)
        for _ in range(25 if backend == capi.BACKEND_TC else 1):
            op.run(tin, tout)
        c.stream_sync()
        assert op.backend == backend
        outs[backend] = tout.read_chw()
        for o in (tin, tout, op):
            o.destroy()
    assert rel_l2(outs[capi.BACKEND_TC], outs[capi.BACKEND_DIRECT]) <= 2e-3


@pytest.mark.parametrize("k,ds,ci,co,inp,outp,postbn,res,relures,bnres,size,batch", [
    (1, 1, 64, 256, 0, 0, True, False, False, False, 14, 2),     # bottleneck expand, 1x1
    (1, 1, 256, 64, 0, 1, True, False, False, False, 14, 1),     # reduce, output padded for the following 3x3
    (3, 1, 64, 64, 1, 0, True, False, False, False, 28, 2),      # 3x3 on padded tiles
    (3, 2, 128, 128, 1, 0, True, False, False, False, 14, 3),    # 3x3 stride 2
    (1, 2, 256, 512, 0, 0, False, False, False, False, 14, 2),   # projection shortcut, stride 2
    (1, 1, 128, 512, 0, 0, True, True, True, True, 7, 3),        # residual + ReLU + BN on the residual
    (1, 1, 2048, 1000, 0, 0, False, False, False, False, 1, 3),  # GEMM72: 1x1 spatial, N tile with a ragged tail
    (3, 1, 512, 512, 1, 0, True, False, False, False, 7, 1),     # deepest 3x3: 72 K stages
    (7, 2, 3, 64, 1, 1, True, False, False, False, 32, 2),       # the stem: one input plane, 49 taps packed 16 per stage, under-padded
    (3, 1, 4, 24, 1, 0, False, False, False, False, 9, 1),       # tap-packed, four channels
])
def test_deep_conv_tcgen05(k, ds, ci, co, inp, outp, postbn, res, relures, bnres, size, batch):
    """Deep-tiled convolutions on the tensor cores (fyn_conv_deep_tc.cu) against the oracle's deep semantics (fp16-truncated
    weights, fp16 bias / BN parameters, zero border from the tile padding) and against the direct kernel."""
    rng = np.random.default_rng(k * 100 + ci + co + size)
    x = rng.normal(size=(batch, ci, size, size)).astype(np.float32)
    wb = random_wb(rng, ci, co, k, post_bn=postbn)
    so = size // ds
    residual = rng.normal(size=(batch, co, so, so)).astype(np.float32) if res else None
    fl = capi.FLAG_PRE_RELU | (capi.FLAG_POST_BATCHNORM if postbn else 0) | (capi.FLAG_RELU_ON_RESIDUAL if relures else 0) | \
         (capi.FLAG_BATCHNORM_ON_RESIDUAL if bnres else 0)
    kw = dict(out_channels=co, kernel=k, downsample=ds, in_pad=inp, out_pad=outp, flags=fl, residual=residual, deep=True)
    y, be, _ = conv_gpu(x, wb, backend=capi.BACKEND_TC, want_op=True, **kw)
    assert be == 2
    yd = conv_gpu(x, wb, backend=capi.BACKEND_DIRECT, **kw)
    ofl = (fo.POST_BATCHNORM if postbn else 0) | (fo.RELU_ON_RESIDUAL if relures else 0) | (fo.BATCHNORM_ON_RESIDUAL if bnres else 0)
    xs = half(x)
    ref = np.stack([fo.conv2d(xs[i], wb, co, k, downsample=ds, in_pad=inp, out_pad=outp, act=fo.ACT_RELU, flags=ofl, deep=True,
                              residual=None if residual is None else half(residual[i]), prec=fo.FP16_STORE) for i in range(batch)])
    # same operands as the oracle (truncated fp16 weights, fp16 activations), fp32 accumulation in a different order
    assert_close_f16(y, ref, None, ulps=1.01, extra_abs=1e-4 * max(1.0, float(np.abs(ref).max()) / 8))
    assert rel_l2(y, yd) <= 1e-3


DEEP_CASES = [(1, 1, 64, 256, 0, 0, True, False, False, False, 14, 2), (1, 1, 256, 64, 0, 1, True, False, False, False, 14, 1),
              (3, 1, 64, 64, 1, 0, True, False, False, False, 28, 2), (3, 2, 128, 128, 1, 0, True, False, False, False, 14, 3),
              (1, 1, 128, 512, 0, 0, True, True, True, True, 7, 3), (1, 1, 2048, 1000, 0, 0, False, False, False, False, 1, 3),
              (3, 1, 512, 512, 1, 0, True, False, False, False, 7, 1), (7, 2, 3, 64, 1, 1, True, False, False, False, 32, 2),
              (3, 1, 4, 24, 1, 0, False, False, False, False, 9, 1), (1, 1, 64, 16, 0, 0, False, True, False, False, 20, 5),
              (1, 1, 64, 256, 0, 0, True, True, True, False, 56, 12), (3, 1, 64, 64, 1, 0, True, False, False, False, 56, 10),
              (3, 1, 128, 128, 1, 1, True, True, True, False, 28, 7), (3, 1, 256, 272, 1, 0, False, False, False, False, 14, 9),
              (3, 1, 64, 32, 1, 0, True, False, False, False, 61, 3), (3, 1, 64, 64, 2, 0, True, False, False, False, 20, 3)]


@pytest.mark.parametrize("k,ds,ci,co,inp,outp,postbn,res,relures,bnres,size,batch", DEEP_CASES)
def test_deep_conv_persistent_kernel_is_bit_identical(k, ds, ci, co, inp, outp, postbn, res, relures, bnres, size, batch, monkeypatch):
    """Large grids run the persistent kernel (k_conv_deep_tc_p: one CTA per SM, tiles of one CTA pipelined through two TMEM
    halves), small ones the one-tile-per-CTA kernel.  FYN_DEEP_PERSIST=2 forces the persistent kernel on any grid, =0 forbids it:
    same operands, same accumulation order, so every layer shape must give the same bits -- single-tile grids, ragged N tiles,
    a single column group (16 outputs: half of the epilogue warps have nothing to read), the tap-packed stem, 72 K stages, and
    grids of several tiles per CTA with residual."""
    rng = np.random.default_rng(k * 100 + ci + co + size)
    x = half(rng.normal(size=(batch, ci, size, size)).astype(np.float32))
    wb = random_wb(rng, ci, co, k, post_bn=postbn)
    so = size // ds
    residual = half(rng.normal(size=(batch, co, so, so)).astype(np.float32)) if res else None
    fl = capi.FLAG_PRE_RELU | (capi.FLAG_POST_BATCHNORM if postbn else 0) | (capi.FLAG_RELU_ON_RESIDUAL if relures else 0) | \
         (capi.FLAG_BATCHNORM_ON_RESIDUAL if bnres else 0)
    kw = dict(out_channels=co, kernel=k, downsample=ds, in_pad=inp, out_pad=outp, flags=fl, residual=residual, deep=True, backend=capi.BACKEND_TC)
    outs = []
    monkeypatch.setenv("FYN_DEEP_HALO", "0")
    for mode in ("0", "2"):
        monkeypatch.setenv("FYN_DEEP_PERSIST", mode)
        outs.append(conv_gpu(x, wb, **kw))
    np.testing.assert_array_equal(outs[0], outs[1])
    for ring, sets in (("2", "2"), ("3", "1"), ("5", "4")):      # short rings / fewer loader sets: the stage stream wraps inside and across tiles
        monkeypatch.setenv("FYN_DEEP_PRING", ring)
        monkeypatch.setenv("FYN_DEEP_SETS", sets)
        np.testing.assert_array_equal(outs[0], conv_gpu(x, wb, **kw))
    if k == 3 and ds == 1 and ci > 4:
        # 3x3 stride-1 layers on large grids run the halo-tile kernel (k_conv_deep_tc_h3): one gather per tile, nine shifted MMA
        # groups, channel stages outermost -- the same products summed in another order, so the bound is the oracle's, not equality
        monkeypatch.delenv("FYN_DEEP_HALO")
        monkeypatch.delenv("FYN_DEEP_PRING")
        ofl = (fo.POST_BATCHNORM if postbn else 0) | (fo.RELU_ON_RESIDUAL if relures else 0) | (fo.BATCHNORM_ON_RESIDUAL if bnres else 0)
        ref = np.stack([fo.conv2d(x[i], wb, co, k, downsample=ds, in_pad=inp, out_pad=outp, act=fo.ACT_RELU, flags=ofl, deep=True,
                                  residual=None if residual is None else residual[i], prec=fo.FP16_STORE) for i in range(min(batch, 2))])
        for sets in ("1", "2", "3"):
            monkeypatch.setenv("FYN_DEEP_SETS", sets)
            y = conv_gpu(x, wb, **kw).reshape(outs[0].shape)
            assert_close_f16(y.reshape((batch,) + ref.shape[1:])[:ref.shape[0]], ref, None, ulps=1.01, extra_abs=1e-4 * max(1.0, float(np.abs(ref).max()) / 8))
            assert rel_l2(y, outs[0]) <= 2e-4


SPLITK_CASES = [  # k, ds, ci, co, inp, outp, postbn, res, relures, bnres, size, batch -- ResNet-50's layers at batch 1 ... 3
    (1, 1, 2048, 512, 0, 1, True, False, False, False, 7, 1),     # deepest reduction: 32 K stages over a cluster of 8, 32-column tiles
    (1, 1, 512, 2048, 0, 0, True, True, True, False, 7, 1),       # expansion + residual: cluster of 4
    (3, 1, 512, 512, 1, 0, True, False, False, False, 7, 1),      # 72 stages: nine per CTA, the ring wraps
    (1, 1, 1024, 256, 0, 1, True, False, False, False, 14, 1),    # two pixel tiles, the second partly empty
    (3, 1, 256, 256, 1, 0, True, False, False, False, 14, 1),     # 36 stages: uneven K slices (4 / 5 per CTA)
    (1, 2, 1024, 2048, 0, 0, False, False, False, False, 14, 1),  # projection shortcut, stride 2
    (1, 1, 2048, 1000, 0, 0, False, False, False, False, 1, 3),   # GEMM72: ragged last sub-tile (1000 = 15 x 64 + 40)
    (1, 1, 64, 64, 0, 1, True, False, False, False, 56, 1),       # single stage: no cluster, reduction buffer only
    (1, 1, 64, 256, 0, 0, True, True, True, False, 56, 1),        # single stage + residual
    (1, 1, 256, 64, 0, 1, True, False, False, False, 56, 1),      # one stage per CTA of a cluster of 4
    (3, 1, 128, 128, 1, 1, True, True, True, True, 28, 1),        # BN on the residual
    (1, 1, 128, 80, 0, 0, False, False, False, False, 9, 2),      # 80 outputs: the operand image's tile is not a multiple of 64
    (1, 1, 192, 24, 0, 0, False, True, False, False, 5, 3),       # three stages, 24 outputs (32-column TMEM tile, six planes)
]


@pytest.mark.parametrize("k,ds,ci,co,inp,outp,postbn,res,relures,bnres,size,batch", SPLITK_CASES)
def test_deep_conv_splitk_cluster_kernel(k, ds, ci, co, inp, outp, postbn, res, relures, bnres, size, batch, monkeypatch):
    """Small grids run k_conv_deep_tc_sk: the K stages of an output tile split over a thread-block cluster, partial tiles
    reduce-scattered through distributed shared memory, tiles narrowed to <= 64 columns.  Same operands as the one-tile kernel,
    K slices summed separately: checked against the oracle (1 fp16 ulp + summation noise) and against the one-tile kernel
    (FYN_DEEP_SPLITK=0), for the default plan and for forced tile widths / cluster sizes."""
    rng = np.random.default_rng(k * 100 + ci + co + size)
    x = half(rng.normal(size=(batch, ci, size, size)).astype(np.float32))
    wb = random_wb(rng, ci, co, k, post_bn=postbn)
    so = size // ds
    residual = half(rng.normal(size=(batch, co, so, so)).astype(np.float32)) if res else None
    fl = capi.FLAG_PRE_RELU | (capi.FLAG_POST_BATCHNORM if postbn else 0) | (capi.FLAG_RELU_ON_RESIDUAL if relures else 0) | \
         (capi.FLAG_BATCHNORM_ON_RESIDUAL if bnres else 0)
    kw = dict(out_channels=co, kernel=k, downsample=ds, in_pad=inp, out_pad=outp, flags=fl, residual=residual, deep=True, backend=capi.BACKEND_TC)
    ofl = (fo.POST_BATCHNORM if postbn else 0) | (fo.RELU_ON_RESIDUAL if relures else 0) | (fo.BATCHNORM_ON_RESIDUAL if bnres else 0)
    ref = np.stack([fo.conv2d(x[i], wb, co, k, downsample=ds, in_pad=inp, out_pad=outp, act=fo.ACT_RELU, flags=ofl, deep=True,
                              residual=None if residual is None else residual[i], prec=fo.FP16_STORE) for i in range(batch)])
    monkeypatch.setenv("FYN_DEEP_SPLITK", "0")
    monkeypatch.setenv("FYN_DEEP_HALO", "0")
    base = conv_gpu(x, wb, **kw)
    assert gpu_util.LAST_KERNEL in (10, 11)      # one-tile or persistent kernel: the same K order, bit-identical to each other
    monkeypatch.setenv("FYN_DEEP_SPLITK", "2")
    for nt, split in ((None, None), ("64", "2"), ("32", "8"), ("16", "3"), ("64", "1")):
        for name, v in (("FYN_DEEP_SK_NT", nt), ("FYN_DEEP_SK_SPLIT", split)):
            if v is None:
                monkeypatch.delenv(name, raising=False)
            else:
                monkeypatch.setenv(name, v)
        y = conv_gpu(x, wb, **kw)
        assert gpu_util.LAST_KERNEL & 255 == 13, "the split-K cluster kernel must have run"
        if split is not None:
            assert (gpu_util.LAST_KERNEL >> 8) & 255 == min(int(split), k * k * (ci // 64))
        assert_close_f16(y.reshape(ref.shape), ref, None, ulps=1.01, extra_abs=1e-4 * max(1.0, float(np.abs(ref).max()) / 8))
        assert rel_l2(y, base) <= 2e-4


def test_deep_conv_fused_input_batchnorm():
    """fyn_conv2d_set_input_norm: a deep 1x1 convolution that evaluates the batch-norm layer in front of it at the fetch is
    bit-identical to running fyn_batchnorm_run first (same fp32 fma, same fp16 rounding); layers outside the deep-tiled
    tcgen05 family refuse, and so does a run on fp32 tensors."""
    c = ctx()
    rng = np.random.default_rng(31)
    for ci, co, size, relu, res in [(256, 64, 14, True, False), (64, 72, 9, False, True), (512, 128, 7, True, False)]:
        x = rng.normal(size=(3, ci, size, size)).astype(np.float32)
        sb = np.concatenate([rng.uniform(0.5, 1.5, ci), rng.uniform(-0.5, 0.5, ci)]).astype(np.float32)
        wb = np.concatenate([rng.uniform(-0.5, 0.5, co), rng.normal(0, np.sqrt(2.0 / ci), co * ci)]).astype(np.float32)
        flags = capi.FLAG_DEEP | (capi.FLAG_PRE_RELU if relu else 0) | (capi.FLAG_RESIDUAL_INPUT if res else 0)
        conv = capi.Conv2d(c, wb, width=size, height=size, in_channels=ci, out_channels=co, kernel=1, flags=flags)
        assert conv.backend == capi.BACKEND_TC
        bn = capi.BatchNorm(c, sb, width=size, height=size, channels=ci, flags=capi.FLAG_DEEP)
        tin, tbn = (c.tensor(size, size, ci, 0, capi.ORDER_DEEP, capi.F16, 3) for _ in range(2))
        t1, t2 = (c.tensor(size, size, co, 0, capi.ORDER_DEEP, capi.F16, 3) for _ in range(2))
        tres = c.tensor(size, size, co, 0, capi.ORDER_DEEP, capi.F16, 3) if res else None
        tin.write_chw(x)
        if res:
            tres.write_chw(rng.normal(size=(3, co, size, size)).astype(np.float32))
        bn.run(tin, tbn)
        conv.run(tbn, t1, tres)
        conv.set_input_norm(sb)
        conv.run(tin, t2, tres)
        np.testing.assert_array_equal(t1.read_chw(), t2.read_chw())
        conv.set_input_norm(None)
        conv.run(tbn, t2, tres)
        np.testing.assert_array_equal(t1.read_chw(), t2.read_chw())
        for o in (conv, bn, tin, tbn, t1, t2, tres):
            if o is not None:
                o.destroy()
    # 3x3 layers read padding texels (zero in the stand-alone layer's output, bn(0) at a fused fetch): refused
    wb = np.zeros(64 + 64 * 9 * 64, np.float32)
    conv3 = capi.Conv2d(c, wb, width=8, height=8, in_channels=64, out_channels=64, kernel=3, in_padding=1, flags=capi.FLAG_DEEP)
    with pytest.raises(capi.FynError):
        conv3.set_input_norm(np.ones(128, np.float32))
    conv3.destroy()
    shallow = capi.Conv2d(c, np.zeros(8 + 8 * 8, np.float32), width=8, height=8, in_channels=8, out_channels=8, kernel=1)
    with pytest.raises(capi.FynError):
        shallow.set_input_norm(np.ones(16, np.float32))
    shallow.destroy()
    # fp32 tensors take the direct kernel, which does not implement the fusion: the run fails loudly
    wb = np.concatenate([np.zeros(64), rng.normal(size=64 * 64)]).astype(np.float32)
    conv = capi.Conv2d(c, wb, width=6, height=6, in_channels=64, out_channels=64, kernel=1, flags=capi.FLAG_DEEP)
    conv.set_input_norm(np.ones(128, np.float32))
    a, b = (c.tensor(6, 6, 64, 0, capi.ORDER_DEEP, capi.F32) for _ in range(2))
    with pytest.raises(capi.FynError):
        conv.run(a, b)
    for o in (conv, a, b):
        o.destroy()


def test_deep_conv_wide_tiles_on_large_grids(monkeypatch):
    """Short-K layers with >= 256 outputs switch to 256-column CTAs once the grid is large (fyn_conv_deep_tc.cu: second
    operand image, chosen at run time): the result must not depend on the tile width.  Checked against the direct kernel on the
    whole batch and against the oracle on one image (tolerance of the deep tcgen05 tests).  The small-grid run of the bit-exact
    comparison uses the one-tile kernel (FYN_DEEP_SPLITK=0): the split-K cluster kernel that small grids take by default sums
    its K slices separately (multi-stage layers: within the oracle tolerance, checked below, not the same bits)."""
    rng = np.random.default_rng(61)
    for ci, co, size, batch, res, bn in [(64, 256, 56, 8, True, False), (128, 512, 28, 16, False, True), (64, 300, 40, 12, True, True),
                                         (256, 1024, 14, 32, True, True)]:   # multi-stage: one 256-column CTA per SM
        x = half(rng.normal(size=(batch, ci, size, size)).astype(np.float32))
        wb = random_wb(rng, ci, co, 1, post_bn=bn)
        r = half(rng.normal(size=(batch, co, size, size)).astype(np.float32)) if res else None
        flags = capi.FLAG_PRE_RELU | (capi.FLAG_POST_BATCHNORM if bn else 0) | (capi.FLAG_RELU_ON_RESIDUAL if res else 0)
        monkeypatch.setenv("FYN_DEEP_SPLITK", "0")
        y_small, be = conv_gpu(x[:1], wb, out_channels=co, kernel=1, deep=True, flags=flags, residual=None if r is None else r[:1], want_op=True)[:2]
        monkeypatch.delenv("FYN_DEEP_SPLITK")
        y_sk = conv_gpu(x[:1], wb, out_channels=co, kernel=1, deep=True, flags=flags, residual=None if r is None else r[:1])
        assert gpu_util.LAST_KERNEL & 255 == 13 and rel_l2(y_sk, y_small) <= 2e-4
        y, be = conv_gpu(x, wb, out_channels=co, kernel=1, deep=True, flags=flags, residual=r, want_op=True)[:2]
        assert be == capi.BACKEND_TC
        yd = conv_gpu(x, wb, out_channels=co, kernel=1, deep=True, flags=flags, residual=r, backend=capi.BACKEND_DIRECT)
        assert rel_l2(y, yd) <= 1e-4
        np.testing.assert_array_equal(y[0], y_small)            # 128-column tiles (small grid) == 256-column tiles
        ref = fo.conv2d(x[0], wb, co, 1, flags=(fo.POST_BATCHNORM if bn else 0) | (fo.RELU_ON_RESIDUAL if res else 0), deep=True,
                        residual=None if r is None else r[0], act=fo.ACT_RELU, prec=fo.FP16_STORE)
        assert_close_f16(y[0], ref, extra_abs=2e-4)
    # 3x3, 256 -> 256 on 14x14 at batch 128 (the grid size at which ResNet-50's stage-3 layers switch)
    x = half(rng.normal(size=(128, 256, 14, 14)).astype(np.float32))
    wb = random_wb(rng, 256, 256, 3, post_bn=True)
    kw = dict(out_channels=256, kernel=3, in_pad=1, deep=True, flags=capi.FLAG_PRE_RELU | capi.FLAG_POST_BATCHNORM)
    y = conv_gpu(x, wb, **kw)
    assert rel_l2(y, conv_gpu(x, wb, backend=capi.BACKEND_DIRECT, **kw)) <= 1e-4
    assert rel_l2(y[:2], conv_gpu(x[:2], wb, **kw)) <= 2e-4    # (large grid: halo-tile kernel, another summation order)


@pytest.mark.parametrize("k,ds,ci,co,relu", [(1, 1, 8, 5, False), (1, 2, 16, 8, True), (1, 1, 40, 40, True), (1, 1, 3, 12, True), (5, 1, 6, 7, True),
                                              (7, 2, 4, 8, False), (5, 1, 24, 16, True), (7, 1, 12, 12, True), (1, 2, 5, 3, False)])
def test_tc_1x1_5x5_7x7_and_thin_inputs(k, ds, ci, co, relu):
    """vanilla::ConvLayer1x1 (gpu/vanilla/convlayer1x1_vanilla.cpp:81-150) and the 5x5 / 7x7 / few-channel shapes of
    ConvLayerNxN on the tcgen05 family: windows of one row and one tap, inputs that fill only part of an 8-channel chunk."""
    rng = np.random.default_rng(k * 1000 + ci * 10 + co)
    h, w = 26, 150
    x = rng.normal(size=(ci, h, w)).astype(np.float32)
    wb = random_wb(rng, ci, co, k)
    _run(x, wb, co, k, ds=ds, relu=relu)
    res = rng.normal(size=(co, h // ds, w // ds)).astype(np.float32)
    _run(x, wb, co, k, ds=ds, relu=relu, residual=res, relu_res=True)
