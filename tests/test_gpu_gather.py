"""GPU parity of the SURVEY 8f rank-2 layers (scale / padding / relu / clip, add / sub / singleton arithmetic, concat,
rgb2bgr, shallow <-> deep) through the C ABI against the CPU oracle.

Tolerances: these layers copy or combine stored values, so FYN_F32 storage must match the fp32 oracle to 1e-6
(bilinear weights are exact rationals; the only freedom is fma contraction) and FYN_F16 storage to 1 fp16 ulp of the
fp16-store oracle (exactly equal for the pure copies)."""
import numpy as np
import pytest

import fyn_oracle as fo
from fyusenet_b200 import capi
from gpu_util import assert_close_f16, ctx, half

pytestmark = pytest.mark.gpu

ORDER = {False: capi.ORDER_SHALLOW, True: capi.ORDER_DEEP}


def _prep(x, dtype):
    return (half(x), fo.FP16_STORE) if dtype == capi.F16 else (np.asarray(x, np.float32), fo.FP32)


def _compare(y, ref, dtype, exact=False):
    if exact:
        np.testing.assert_array_equal(y, ref)
    elif dtype == capi.F32:
        np.testing.assert_allclose(y, ref, rtol=1e-6, atol=1e-6)
    else:
        assert_close_f16(y, ref)


def _border_is_zero(t, pad):
    if not pad:
        return
    raw = t.download()
    assert np.all(raw[..., :pad, :, :] == 0) and np.all(raw[..., :, :pad, :] == 0)


@pytest.mark.parametrize("dtype", [capi.F16, capi.F32])
@pytest.mark.parametrize("deep", [False, True])
def test_scale_layer(deep, dtype):
    c = ctx()
    rng = np.random.default_rng(11)
    cases = [dict(ch=9, h=6, w=8, up=(2, 2), down=(1, 1), linear=False, ip=0, op=0, act=None),
             dict(ch=9, h=6, w=8, up=(2, 3), down=(1, 1), linear=True, ip=1, op=1, act="relu"),
             dict(ch=12, h=12, w=18, up=(1, 1), down=(2, 2), linear=False, ip=1, op=0, act=None),
             dict(ch=12, h=12, w=18, up=(1, 1), down=(2, 2), linear=True, ip=0, op=1, act=None),
             dict(ch=8, h=9, w=9, up=(1, 1), down=(3, 3), linear=True, ip=0, op=0, act=None),
             dict(ch=8, h=4, w=4, up=(2, 2), down=(1, 1), linear=True, ip=0, op=0, act=None),      # deep: bleeds into the next tile
             dict(ch=23, h=7, w=5, up=(1, 1), down=(1, 1), linear=False, ip=0, op=2, act=None),    # PADDING2D
             dict(ch=23, h=7, w=5, up=(1, 1), down=(1, 1), linear=False, ip=1, op=0, act="relu"),  # RELU
             dict(ch=6, h=7, w=5, up=(1, 1), down=(1, 1), linear=False, ip=0, op=0, act="clip"),   # CLIP
             dict(ch=8, h=1, w=5, up=(2, 2), down=(1, 1), linear=True, ip=0, op=0, act=None),      # deep 1-texel rows: nearest
             dict(ch=40, h=33, w=70, up=(4, 4), down=(1, 1), linear=True, ip=1, op=0, act=None)]
    for cs in cases:
        x = rng.normal(size=(2, cs["ch"], cs["h"], cs["w"])).astype(np.float32)
        flags = (capi.FLAG_DEEP if deep else 0) | {None: 0, "relu": capi.FLAG_PRE_RELU, "clip": capi.FLAG_PRE_CLIP}[cs["act"]]
        op = capi.Scale(c, width=cs["w"], height=cs["h"], channels=cs["ch"], in_padding=cs["ip"], out_padding=cs["op"],
                        up=cs["up"], down=cs["down"], linear=cs["linear"], flags=flags, clip_lo=-0.25, clip_hi=0.5)
        tin = c.tensor(cs["w"], cs["h"], cs["ch"], cs["ip"], ORDER[deep], dtype, 2)
        tout = c.tensor(op.out_width, op.out_height, cs["ch"], cs["op"], ORDER[deep], dtype, 2)
        tin.write_chw(x)
        op.run(tin, tout)
        y = tout.read_chw()
        xs, prec = _prep(x, dtype)
        act = {None: fo.ACT_NONE, "relu": fo.ACT_RELU, "clip": fo.ACT_CLIP}[cs["act"]]
        ref = np.stack([fo.scale(xs[i], up=cs["up"], down=cs["down"], linear=cs["linear"], in_pad=cs["ip"], deep=deep, act=act,
                                 lo=-0.25, hi=0.5, prec=prec) for i in range(2)])
        assert y.shape == ref.shape, cs
        _compare(y, ref, dtype, exact=not cs["linear"])
        if not deep:
            _border_is_zero(tout, cs["op"])
        for o in (tin, tout, op):
            o.destroy()


@pytest.mark.parametrize("dtype", [capi.F16, capi.F32])
@pytest.mark.parametrize("deep", [False, True])
def test_arith_layers(deep, dtype):
    """unit_tests/arithtests.cpp:208-247 (constant operands, |err| <= 0.5 there) plus random operands with activation."""
    c = ctx()
    rng = np.random.default_rng(12)
    flag = capi.FLAG_DEEP if deep else 0
    for a, b, w, h, ch in [(3.0, 30.0, 400, 300, 4), (-2.0, 1.0, 200, 200, 5), (10.0, -10.0, 16, 16, 40), (-100.0, 23.0, 55, 57, 30),
                           (15.0, -16.0, 99, 52, 47)]:
        t1, t2, tout = (c.tensor(w, h, ch, 0, ORDER[deep], dtype) for _ in range(3))
        t1.write_chw(np.full((ch, h, w), a, np.float32))
        t2.write_chw(np.full((ch, h, w), b, np.float32))
        for opc, expect in [(capi.ARITH_ADD, a + b), (capi.ARITH_SUB, a - b), (capi.ARITH_MUL, a * b), (capi.ARITH_DIV, a / b)]:
            op = capi.Arith(c, width=w, height=h, channels=ch, op=opc, operand=b, flags=flag)
            op.run(t1, None, tout)
            y = tout.read_chw()
            assert np.all(np.abs(y - expect) <= 0.5)
            _compare(y, fo.arith(np.full((ch, h, w), a, np.float32), b, opc, prec=_prep(0, dtype)[1]), dtype)
            op.destroy()
            if opc <= capi.ARITH_SUB:
                op = capi.Arith(c, width=w, height=h, channels=ch, op=opc, flags=flag)
                op.run(t1, t2, tout)
                assert np.all(np.abs(tout.read_chw() - expect) <= 0.5)
                op.destroy()
        for o in (t1, t2, tout):
            o.destroy()
    # random operands, ReLU at the fetch, padded tensors, batch 2
    ch, h, w = 14, 9, 21
    x1, x2 = (rng.normal(size=(2, ch, h, w)).astype(np.float32) for _ in range(2))
    t1, t2 = (c.tensor(w, h, ch, 1, ORDER[deep], dtype, 2) for _ in range(2))
    tout = c.tensor(w, h, ch, 2, ORDER[deep], dtype, 2)
    t1.write_chw(x1)
    t2.write_chw(x2)
    (s1, prec), (s2, _) = _prep(x1, dtype), _prep(x2, dtype)
    for opc in (capi.ARITH_ADD, capi.ARITH_SUB):
        op = capi.Arith(c, width=w, height=h, channels=ch, op=opc, in_padding=1, out_padding=2, flags=flag | capi.FLAG_PRE_RELU)
        op.run(t1, t2, tout)
        _compare(tout.read_chw(), fo.arith(s1, s2, opc, act=fo.ACT_RELU, prec=prec), dtype)
        op.destroy()
    with pytest.raises(capi.FynError):
        capi.Arith(c, width=w, height=h, channels=ch, op=capi.ARITH_MUL)
    for o in (t1, t2, tout):
        o.destroy()


@pytest.mark.parametrize("dtype", [capi.F16, capi.F32])
@pytest.mark.parametrize("deep", [False, True])
def test_concat_layer(deep, dtype):
    c = ctx()
    rng = np.random.default_rng(13)
    for chans, relu, ip, op_ in [((8, 4), False, 0, 0), ((3, 8, 6), True, 1, 1), ((5, 5, 5, 5), False, 0, 2), ((40, 24), True, 1, 0),
                                 ((1, 2, 3, 4, 5, 6, 7, 8), False, 0, 0)]:
        h, w = 7, 11
        parts = [rng.normal(size=(2, ch, h, w)).astype(np.float32) for ch in chans]
        tins = [c.tensor(w, h, ch, ip, ORDER[deep], dtype, 2) for ch in chans]
        for t, p in zip(tins, parts):
            t.write_chw(p)
        tout = c.tensor(w, h, sum(chans), op_, ORDER[deep], dtype, 2)
        op = capi.Concat(c, width=w, height=h, channels=chans, in_padding=ip, out_padding=op_,
                         flags=(capi.FLAG_DEEP if deep else 0) | (capi.FLAG_PRE_RELU if relu else 0))
        op.run(tins, tout)
        y = tout.read_chw()
        prec = _prep(0, dtype)[1]
        ref = np.stack([fo.concat([_prep(p[i], dtype)[0] for p in parts], act=fo.ACT_RELU if relu else fo.ACT_NONE, prec=prec) for i in range(2)])
        _compare(y, ref, dtype, exact=True)
        with pytest.raises(capi.FynError):
            op.run(tins[:-1], tout)
        for o in tins + [tout, op]:
            o.destroy()


@pytest.mark.parametrize("dtype", [capi.F16, capi.F32])
def test_rgb2bgr_and_relayout(dtype):
    c = ctx()
    rng = np.random.default_rng(14)
    for ch in (3, 4, 7, 5):
        h, w = 6, 10
        x = rng.normal(size=(ch, h, w)).astype(np.float32)
        tin, tout = c.tensor(w, h, ch, 0, capi.ORDER_SHALLOW, dtype), c.tensor(w, h, ch, 1, capi.ORDER_SHALLOW, dtype)
        tin.write_chw(x)
        op = capi.RGB2BGR(c, width=w, height=h, channels=ch, out_padding=1)
        op.run(tin, tout)
        xs, prec = _prep(x, dtype)
        _compare(tout.read_chw(), fo.rgb2bgr(xs, prec=prec), dtype, exact=True)
        for o in (tin, tout, op):
            o.destroy()
    # shallow -> deep -> shallow, ReLU applied on the way in, identity on the way back
    for ch, h, w, ip, op_ in [(24, 9, 13, 1, 1), (7, 5, 5, 0, 1), (64, 16, 16, 1, 0)]:
        x = rng.normal(size=(2, ch, h, w)).astype(np.float32)
        ts = c.tensor(w, h, ch, ip, capi.ORDER_SHALLOW, dtype, 2)
        td = c.tensor(w, h, ch, op_, capi.ORDER_DEEP, dtype, 2)
        tb = c.tensor(w, h, ch, ip, capi.ORDER_SHALLOW, dtype, 2)
        ts.write_chw(x)
        s2d = capi.Relayout(c, width=w, height=h, channels=ch, in_padding=ip, out_padding=op_, flags=capi.FLAG_PRE_RELU)
        d2s = capi.Relayout(c, width=w, height=h, channels=ch, in_padding=op_, out_padding=ip)
        s2d.run(ts, td)
        d2s.run(td, tb)
        xs, prec = _prep(x, dtype)
        np.testing.assert_array_equal(td.read_chw(), np.maximum(xs, 0))
        np.testing.assert_array_equal(tb.read_chw(), np.maximum(xs, 0))
        # the deep tensor really is the tiled texture of the oracle's packer
        tex = td.download().reshape(2, -1)[0]
        np.testing.assert_array_equal(tex, fo.pack_deep(np.maximum(xs[0], 0), op_).reshape(-1))
        for o in (ts, td, tb, s2d, d2s):
            o.destroy()


@pytest.mark.parametrize("deep", [False, True])
def test_plane_fast_path_matches_generic_kernel(deep, monkeypatch):
    """fp16 add / sub / singleton / concat / relayout run the plane-chunk kernel (k_plane_h4); FYN_GATHER_GENERIC=1 forces the
    one-thread-per-texel gather kernel. Same arithmetic, so the bits must agree -- at every unroll width (H*W of 30 ... 2.5k)."""
    c = ctx()
    rng = np.random.default_rng(99)
    flag = capi.FLAG_DEEP if deep else 0
    for w, h, ch, ip, op_ in [(6, 5, 12, 1, 0), (19, 17, 24, 0, 1), (31, 23, 8, 1, 1), (57, 45, 20, 0, 0)]:
        x1, x2 = (rng.normal(size=(2, ch, h, w)).astype(np.float32) for _ in range(2))
        t1, t2 = (c.tensor(w, h, ch, ip, ORDER[deep], capi.F16, 2) for _ in range(2))
        tout = c.tensor(w, h, ch, op_, ORDER[deep], capi.F16, 2)
        tcat = c.tensor(w, h, 2 * ch, op_, ORDER[deep], capi.F16, 2)
        tre = c.tensor(w, h, ch, op_, ORDER[not deep], capi.F16, 2)
        t1.write_chw(x1)
        t2.write_chw(x2)
        ops = [(capi.Arith(c, width=w, height=h, channels=ch, op=capi.ARITH_ADD, in_padding=ip, out_padding=op_, flags=flag | capi.FLAG_PRE_RELU), (t1, t2, tout)),
               (capi.Arith(c, width=w, height=h, channels=ch, op=capi.ARITH_SUB, in_padding=ip, out_padding=op_, flags=flag), (t1, t2, tout)),
               (capi.Arith(c, width=w, height=h, channels=ch, op=capi.ARITH_DIV, operand=3.0, in_padding=ip, out_padding=op_, flags=flag), (t1, None, tout)),
               (capi.Concat(c, width=w, height=h, channels=(ch, ch), in_padding=ip, out_padding=op_, flags=flag | capi.FLAG_PRE_RELU), ([t1, t2], tcat)),
               (capi.Relayout(c, width=w, height=h, channels=ch, in_padding=ip, out_padding=op_), (t1, tre))]
        for op, args in ops:
            got = []
            for generic in ("0", "1"):
                monkeypatch.setenv("FYN_GATHER_GENERIC", generic)
                op.run(*args)
                got.append(args[-1].download().copy())
            np.testing.assert_array_equal(got[0], got[1])
            op.destroy()
        for o in (t1, t2, tout, tcat, tre):
            o.destroy()


@pytest.mark.parametrize("dtype", [capi.F16, capi.F32])
@pytest.mark.parametrize("deep", [False, True])
def test_depthwise_conv3x3(deep, dtype):
    """vanilla::DepthwiseConvLayer3x3 / deep::DeepDepthwiseConvLayer3x3 against the oracle (itself checked against the pinned
    regular-convolution oracle with a diagonal weight matrix).  Tolerance: FYN_F32 2e-5 (summation order), FYN_F16 1 fp16 ulp
    of the fp16-store oracle + 2e-5."""
    c = ctx()
    rng = np.random.default_rng(21)
    cases = [dict(ch=10, h=9, w=12, ds=1, dil=1, ip=1, op=0, bn=False, relu=False, quirks=None),
             dict(ch=24, h=16, w=20, ds=2, dil=1, ip=1, op=1, bn=False, relu=True, quirks=None),
             dict(ch=7, h=8, w=8, ds=1, dil=1, ip=1, op=0, bn=True, relu=False, quirks=None),     # reference BN read position (shallow)
             dict(ch=7, h=8, w=8, ds=1, dil=1, ip=1, op=0, bn=True, relu=False, quirks=0),
             dict(ch=12, h=11, w=13, ds=1, dil=1, ip=0, op=0, bn=False, relu=False, quirks=None),  # clamp-to-edge / tile bleed
             dict(ch=64, h=28, w=28, ds=1, dil=2, ip=2, op=1, bn=True, relu=True, quirks=None)]    # dilation: deep only
    for cs in cases:
        if cs["dil"] > 1 and not deep:
            with pytest.raises(capi.FynError):
                capi.DwConv3x3(c, np.zeros(cs["ch"] * 12, np.float32), width=cs["w"], height=cs["h"], channels=cs["ch"], dilation=cs["dil"])
            continue
        ch = cs["ch"]
        x = rng.normal(size=(2, ch, cs["h"], cs["w"])).astype(np.float32)
        wb = np.concatenate([rng.uniform(0.5, 1.5, ch), rng.normal(size=ch * 9) * 0.4, rng.uniform(0.5, 1.5, ch), rng.uniform(-0.2, 0.2, ch)]).astype(np.float32)
        flags = (capi.FLAG_DEEP if deep else 0) | (capi.FLAG_POST_BATCHNORM if cs["bn"] else 0) | (capi.FLAG_PRE_RELU if cs["relu"] else 0)
        op = capi.DwConv3x3(c, wb, width=cs["w"], height=cs["h"], channels=ch, downsample=cs["ds"], dilation=cs["dil"], in_padding=cs["ip"],
                            out_padding=cs["op"], flags=flags, quirks=cs["quirks"])
        tin = c.tensor(cs["w"], cs["h"], ch, cs["ip"], ORDER[deep], dtype, 2)
        tout = c.tensor(op.out_width, op.out_height, ch, cs["op"], ORDER[deep], dtype, 2)
        tin.write_chw(x)
        op.run(tin, tout)
        y = tout.read_chw()
        xs, prec = _prep(x, dtype)
        q = capi.QUIRKS_REFERENCE if cs["quirks"] is None else cs["quirks"]
        kw = dict(downsample=cs["ds"], dilation=cs["dil"], in_pad=cs["ip"], deep=deep, post_bn=cs["bn"], quirks=q,
                  act=fo.ACT_RELU if cs["relu"] else fo.ACT_NONE)
        ref = np.stack([fo.dwconv3x3(xs[i], wb, prec=prec, **kw) for i in range(2)])
        assert y.shape == ref.shape
        if dtype == capi.F32:
            np.testing.assert_allclose(y, ref, rtol=2e-5, atol=2e-5)
        else:
            assert_close_f16(y, ref, np.stack([fo.dwconv3x3(xs[i], wb, prec=fo.FP32, **kw) for i in range(2)]), rl2=3e-3)
        # hot-swap the weights
        wb2 = (wb * 0.5).astype(np.float32)
        op.load_weights(wb2)
        op.run(tin, tout)
        ref2 = np.stack([fo.dwconv3x3(xs[i], wb2, prec=prec, **kw) for i in range(2)])
        if dtype == capi.F32:
            np.testing.assert_allclose(tout.read_chw(), ref2, rtol=2e-5, atol=2e-5)
        else:
            assert_close_f16(tout.read_chw(), ref2)
        for o in (tin, tout, op):
            o.destroy()


@pytest.mark.parametrize("dtype", [capi.F16, capi.F32])
@pytest.mark.parametrize("deep,mult", [(False, 1), (True, 1), (True, 2), (True, 3)])
def test_depthwise_conv3x3_multiplier_and_residual(deep, mult, dtype):
    """Channel multiplier > 1 (deep layers, deepdwconvlayerbase.cpp:40-44,234-246,288-297) and the residual input with its
    ReLU / batch-norm options (conv_dw_3x3.frag:133-140, shaders/deep/residual.inc) against the oracle."""
    c = ctx()
    rng = np.random.default_rng(33 + mult)
    ch, h, w = (16, 14, 18) if deep else (10, 9, 12)
    co = ch * mult
    x = rng.normal(size=(2, ch, h, w)).astype(np.float32)
    res = rng.normal(size=(2, co, h, w)).astype(np.float32)
    wb = np.concatenate([rng.uniform(-0.5, 0.5, co), rng.normal(size=ch * 9 * mult) * 0.4, rng.uniform(0.5, 1.5, co), rng.uniform(-0.2, 0.2, co)]).astype(np.float32)
    for bn, relu_res, bn_res, rp in [(False, False, False, 0), (True, True, False, 1), (True, False, True, 0)]:
        if bn_res and not deep:
            with pytest.raises(capi.FynError):
                capi.DwConv3x3(c, wb, width=w, height=h, channels=ch, in_padding=1, flags=capi.FLAG_RESIDUAL_INPUT | capi.FLAG_BATCHNORM_ON_RESIDUAL)
            continue
        flags = (capi.FLAG_DEEP if deep else 0) | (capi.FLAG_POST_BATCHNORM if bn else 0) | capi.FLAG_PRE_RELU | capi.FLAG_RESIDUAL_INPUT | \
                (capi.FLAG_RELU_ON_RESIDUAL if relu_res else 0) | (capi.FLAG_BATCHNORM_ON_RESIDUAL if bn_res else 0)
        op = capi.DwConv3x3(c, wb, width=w, height=h, channels=ch, in_padding=1, out_padding=1, flags=flags, quirks=0, multiplier=mult, res_padding=rp)
        tin = c.tensor(w, h, ch, 1, ORDER[deep], dtype, 2)
        tres = c.tensor(w, h, co, rp, ORDER[deep], dtype, 2)
        tout = c.tensor(w, h, co, 1, ORDER[deep], dtype, 2)
        tin.write_chw(x)
        tres.write_chw(res)
        op.run(tin, tout, residual=tres)
        y = tout.read_chw()
        xs, prec = _prep(x, dtype)
        rs, _ = _prep(res, dtype)
        kw = dict(in_pad=1, deep=deep, post_bn=bn, quirks=0, act=fo.ACT_RELU, multiplier=mult, relu_on_residual=relu_res, bn_on_residual=bn_res)
        ref = np.stack([fo.dwconv3x3(xs[i], wb, prec=prec, residual=rs[i], **kw) for i in range(2)])
        if dtype == capi.F32:
            np.testing.assert_allclose(y, ref, rtol=2e-5, atol=2e-5)
        else:
            assert_close_f16(y, ref, np.stack([fo.dwconv3x3(xs[i], wb, prec=fo.FP32, residual=rs[i], **kw) for i in range(2)]), rl2=3e-3)
        for o in (tin, tres, tout, op):
            o.destroy()
    if not deep:
        with pytest.raises(capi.FynError):          # the shallow layer has no channel multiplier (convlayer_dw_3x3_vanilla.cpp:49-50)
            capi.DwConv3x3(c, np.zeros(ch * 2 * 12, np.float32), width=w, height=h, channels=ch, multiplier=2)


@pytest.mark.parametrize("dtype", [capi.F16, capi.F32])
@pytest.mark.parametrize("kernel", [2, 3])
def test_transpose_conv_stride2(kernel, dtype):
    """vanilla::TransConvLayer2x2 / 3x3 against the oracle (3x3: checked against the pinned convolution oracle on the
    zero-stuffed input; 2x2: against the closed form per output parity class).  The reference holds no test for these layers:
    parity of the strata conventions rests on the code citations in the oracle.  Tolerance: FYN_F32 2e-5, FYN_F16 1 fp16 ulp."""
    c = ctx()
    rng = np.random.default_rng(41 + kernel)
    for ci, co, h, w, ip, op_, bn, relu, quirks in [(6, 5, 7, 9, 1, 0, False, False, None), (12, 20, 16, 10, 1, 1, True, True, None),
                                                     (3, 8, 5, 6, 0, 0, False, True, None), (8, 3, 9, 9, 1, 0, False, False, 0)]:
        x = rng.normal(size=(2, ci, h, w)).astype(np.float32)
        wb = np.concatenate([rng.uniform(-0.5, 0.5, co), rng.normal(size=co * kernel * kernel * ci) * 0.3, rng.uniform(0.5, 1.5, co),
                             rng.uniform(-0.2, 0.2, co)]).astype(np.float32)
        flags = (capi.FLAG_POST_BATCHNORM if bn else 0) | (capi.FLAG_PRE_RELU if relu else 0)
        op = capi.TransConv2d(c, wb, width=w, height=h, in_channels=ci, out_channels=co, kernel=kernel, in_padding=ip, out_padding=op_,
                              flags=flags, quirks=quirks)
        tin = c.tensor(w, h, ci, ip, capi.ORDER_SHALLOW, dtype, 2)
        tout = c.tensor(2 * w, 2 * h, co, op_, capi.ORDER_SHALLOW, dtype, 2)
        tin.write_chw(x)
        op.run(tin, tout)
        y = tout.read_chw()
        xs, prec = _prep(x, dtype)
        kw = dict(in_pad=ip, post_bn=bn, quirks=capi.QUIRKS_REFERENCE if quirks is None else quirks, act=fo.ACT_RELU if relu else fo.ACT_NONE)
        ref = np.stack([fo.transconv(xs[i], wb, co, kernel, prec=prec, **kw) for i in range(2)])
        assert y.shape == ref.shape == (2, co, 2 * h, 2 * w)
        if dtype == capi.F32:
            np.testing.assert_allclose(y, ref, rtol=2e-5, atol=2e-5)
        else:
            assert_close_f16(y, ref)
        _border_is_zero(tout, op_)
        for o in (tin, tout, op):
            o.destroy()
    with pytest.raises(capi.FynError):               # no residual input, like the reference (deeptransconvlayer3x3.cpp:44-46)
        capi.TransConv2d(c, np.zeros(1000, np.float32), width=4, height=4, in_channels=4, out_channels=4, kernel=3, flags=capi.FLAG_RESIDUAL_INPUT)
    with pytest.raises(capi.FynError):
        capi.TransConv2d(c, np.zeros(1000, np.float32), width=4, height=4, in_channels=4, out_channels=4, kernel=5)


@pytest.mark.parametrize("dtype", [capi.F16, capi.F32])
@pytest.mark.parametrize("kernel", [2, 3])
def test_deep_transpose_conv_stride2(kernel, dtype):
    """deep::DeepTransConvLayer2x2 / 3x3 (deeptransconvlayerbase.cpp, deeptransconv{2x2,3x3}_stride2.*) on deep-tiled tensors against
    the oracle, which tests/test_oracle_arith_scale.py checks against the pinned convolution oracle on the zero-stuffed input.
    Several input / output tiles, padded and un-padded tensors (reads outside the image are zero either way).  Tolerance:
    FYN_F32 2e-5, FYN_F16 1 fp16 ulp of the fp16-store oracle (truncated weights, rounded bias)."""
    c = ctx()
    rng = np.random.default_rng(51 + kernel)
    for ci, co, h, w, ip, op_, bn, relu in [(6, 5, 7, 9, 1, 0, False, False), (12, 20, 8, 10, 1, 1, True, True), (64, 32, 14, 14, 0, 1, False, True),
                                            (3, 8, 5, 6, 0, 0, False, True)]:
        x = rng.normal(size=(2, ci, h, w)).astype(np.float32)
        wb = np.concatenate([rng.uniform(-0.5, 0.5, co), rng.normal(size=co * kernel * kernel * ci) * (0.6 / np.sqrt(ci)), rng.uniform(0.5, 1.5, co),
                             rng.uniform(-0.2, 0.2, co)]).astype(np.float32)
        flags = capi.FLAG_DEEP | (capi.FLAG_POST_BATCHNORM if bn else 0) | (capi.FLAG_PRE_RELU if relu else 0)
        op = capi.TransConv2d(c, wb, width=w, height=h, in_channels=ci, out_channels=co, kernel=kernel, in_padding=ip, out_padding=op_, flags=flags)
        tin = c.tensor(w, h, ci, ip, capi.ORDER_DEEP, dtype, 2)
        tout = c.tensor(2 * w, 2 * h, co, op_, capi.ORDER_DEEP, dtype, 2)
        tin.write_chw(x)
        op.run(tin, tout)
        y = tout.read_chw()
        xs, prec = _prep(x, dtype)
        kw = dict(in_pad=ip, post_bn=bn, act=fo.ACT_RELU if relu else fo.ACT_NONE, deep=True)
        ref = np.stack([fo.transconv(xs[i], wb, co, kernel, prec=prec, **kw) for i in range(2)])
        assert y.shape == ref.shape == (2, co, 2 * h, 2 * w)
        if dtype == capi.F32:
            np.testing.assert_allclose(y, ref, rtol=2e-5, atol=2e-5)
        else:
            assert_close_f16(y, ref, np.stack([fo.transconv(xs[i], wb, co, kernel, prec=fo.FP32, **kw) for i in range(2)]), rl2=4e-3)
        for o in (tin, tout, op):
            o.destroy()
