"""GPU parity tests: every layer of the hot path, through the C ABI, against the CPU oracle.

Tolerances (stated per SURVEY 8d):
  * FYN_F32 storage (HIGH_PRECISION): |gpu - oracle_fp32| <= 2e-5 * max(1, |ref|) (fp32 summation order only)
  * FYN_F16 storage (reference default): within 1 fp16 ulp of the fp16-store oracle element-wise for the
    direct kernels, and rel-L2 <= 2e-3 against the fp32 oracle.
"""
import numpy as np
import pytest

import fyn_oracle as fo
from fyusenet_b200 import capi
from gpu_util import assert_close_f16, conv_gpu, conv_oracle, ctx, half, random_wb, rel_l2

pytestmark = pytest.mark.gpu

ACT = {None: (0, fo.ACT_NONE), "relu": (capi.FLAG_PRE_RELU, fo.ACT_RELU)}


def _check(y, x, wb, dtype, **okw):
    if dtype == capi.F32:
        ref = conv_oracle(x, wb, prec=fo.FP32, **okw)
        np.testing.assert_allclose(y, ref, rtol=2e-5, atol=2e-5)
    else:
        xs = half(x)
        okw2 = dict(okw)
        if okw2.get("residual") is not None:
            okw2["residual"] = half(okw2["residual"])
        ref_s = conv_oracle(xs, wb, prec=fo.FP16_STORE, **okw2)
        ref_f = conv_oracle(xs, wb, prec=fo.FP32, **okw2)
        assert_close_f16(y, ref_s, ref_f)


@pytest.mark.parametrize("dtype", [capi.F16, capi.F32])
@pytest.mark.parametrize("k,ds,ci,co,act,res", [
    (3, 1, 40, 40, "relu", "relu"), (3, 1, 40, 40, None, None), (3, 1, 40, 40, "relu", "plain"),
    (9, 1, 3, 12, "relu", None), (3, 2, 12, 20, "relu", None), (3, 2, 20, 40, "relu", None),
    (1, 1, 8, 5, None, None), (5, 1, 6, 7, "relu", None), (7, 2, 4, 8, None, None), (1, 2, 16, 8, "relu", None),
])
def test_shallow_conv(k, ds, ci, co, act, res, dtype):
    """StyleNet's regular convs (stylenet9x9.cpp:137-186): un-padded, clamp-to-edge, pre-ReLU, residual."""
    rng = np.random.default_rng(k * 100 + ci + co)
    h, w = 22, 38
    x = rng.normal(size=(ci, h, w)).astype(np.float32)
    wb = random_wb(rng, ci, co, k)
    residual = rng.normal(size=(co, h // ds, w // ds)).astype(np.float32) if res else None
    fl, oact = ACT[act]
    rfl = capi.FLAG_RELU_ON_RESIDUAL if res == "relu" else 0
    y = conv_gpu(x, wb, out_channels=co, kernel=k, dtype=dtype, downsample=ds, flags=fl | rfl, residual=residual,
                 backend=capi.BACKEND_DIRECT)
    _check(y, x, wb, dtype, out_channels=co, kernel=k, downsample=ds, act=oact,
           flags=(fo.RELU_ON_RESIDUAL if res == "relu" else 0), residual=residual)


@pytest.mark.parametrize("dtype", [capi.F16, capi.F32])
def test_shallow_conv_padding_postbn_dilation(dtype):
    """Padded input / output (zero ring preserved), post-BN fold, BN on residual, horizontal-only dilation."""
    rng = np.random.default_rng(77)
    ci, co, h, w = 8, 12, 14, 18
    x = rng.normal(size=(ci, h, w)).astype(np.float32)
    wb = random_wb(rng, ci, co, 3, post_bn=True)
    residual = rng.normal(size=(co, h, w)).astype(np.float32)
    fl = capi.FLAG_POST_BATCHNORM | capi.FLAG_BATCHNORM_ON_RESIDUAL | capi.FLAG_PRE_RELU
    y, be, raw = conv_gpu(x, wb, out_channels=co, kernel=3, dtype=dtype, in_pad=1, out_pad=2, res_pad=1, flags=fl,
                          residual=residual, dilation=2, backend=capi.BACKEND_DIRECT, want_op=True)
    _check(y, x, wb, dtype, out_channels=co, kernel=3, in_pad=1, out_pad=2, act=fo.ACT_RELU, dilation=2,
           flags=fo.POST_BATCHNORM | fo.BATCHNORM_ON_RESIDUAL, residual=residual)
    # output padding ring stays exactly zero
    assert raw.shape == (1, 3, h + 4, w + 4, 4)
    assert np.all(raw[:, :, :2] == 0) and np.all(raw[:, :, -2:] == 0) and np.all(raw[:, :, :, :2] == 0) and np.all(raw[:, :, :, -2:] == 0)


@pytest.mark.parametrize("quirks", [capi.QUIRKS_REFERENCE, 0])
@pytest.mark.parametrize("dtype", [capi.F16, capi.F32])
@pytest.mark.parametrize("k,ci,co,step,ds,act", [(3, 40, 20, 0.5, 2, None), (3, 20, 12, 0.25, 2, "relu"), (9, 12, 3, 0.5, 1, "relu")])
def test_fractional_conv(k, ci, co, step, ds, act, dtype, quirks):
    """StyleNet deconv1..3 (stylenet9x9.cpp:192-202) incl. quirks Q1/Q2 (fraconv3x3.frag:14-19, fractional.inc)."""
    rng = np.random.default_rng(k + ci)
    h, w = 12, 20
    x = rng.normal(size=(ci, h, w)).astype(np.float32)
    wb = random_wb(rng, ci, co, k)
    fl, oact = ACT[act]
    y = conv_gpu(x, wb, out_channels=co, kernel=k, dtype=dtype, downsample=ds, source_step=step, fractional=True,
                 flags=fl, quirks=quirks, backend=capi.BACKEND_DIRECT)
    assert y.shape == (co, int(h / (step * ds)), int(w / (step * ds)))
    _check(y, x, wb, dtype, out_channels=co, kernel=k, downsample=ds, source_step=step, fractional=True, act=oact,
           quirks=quirks)


def test_conv_from_rgb32f_upload_texture():
    """conv1 reads the upload layer's 3-channel float32 texture directly (uploadlayer.cpp:371-375,
    convlayerbase_vanilla.cpp:205-210); output is fp16."""
    rng = np.random.default_rng(5)
    h, w = 20, 28
    img = rng.random((h, w, 3), dtype=np.float32)
    wb = random_wb(rng, 3, 12, 9)
    c = ctx()
    op = capi.Conv2d(c, wb, width=w, height=h, in_channels=3, out_channels=12, kernel=9, flags=capi.FLAG_PRE_RELU,
                     backend=capi.BACKEND_DIRECT)
    tin = c.tensor(w, h, 3, 0, capi.ORDER_SHALLOW, capi.F32, 1, packing=3)
    tout = c.tensor(w, h, 12, 0, capi.ORDER_SHALLOW, capi.F16)
    tin.upload(img)
    op.run(tin, tout)
    y = tout.read_chw()
    x = fo.upload_hwc(img)
    np.testing.assert_array_equal(tin.read_chw(), x)
    assert_close_f16(y, fo.conv2d(x, wb, 12, 9, act=fo.ACT_RELU, prec=fo.FP16_STORE), fo.conv2d(x, wb, 12, 9, act=fo.ACT_RELU))


@pytest.mark.parametrize("dtype", [capi.F16, capi.F32])
@pytest.mark.parametrize("k,ds,ci,co,inp,outp,postbn,res,bnres,size", [
    (1, 1, 64, 64, 0, 1, True, False, False, 14), (1, 1, 64, 256, 0, 0, False, True, False, 14),
    (3, 1, 64, 64, 1, 0, True, False, False, 14), (3, 2, 32, 32, 1, 0, True, False, False, 14),
    (1, 2, 64, 128, 0, 0, False, False, False, 14), (1, 1, 32, 128, 0, 0, True, True, True, 7),
    (7, 2, 3, 64, 1, 1, True, False, False, 32), (1, 1, 10, 6, 0, 0, False, False, False, 5),
])
def test_deep_conv(k, ds, ci, co, inp, outp, postbn, res, bnres, size, dtype):
    """ResNet-50's deep convs (resnet50.cpp:212-412): fp16-truncated weights + fp16 bias/BN texture in F16
    mode (deepconvlayerbase.cpp:338-394), zero padding from the tile gaps, stride 2, BN on residual; batch 2."""
    rng = np.random.default_rng(k * 10 + ci + co)
    x = rng.normal(size=(2, ci, size, size)).astype(np.float32)
    wb = random_wb(rng, ci, co, k, post_bn=postbn)
    residual = rng.normal(size=(2, co, size // ds, size // ds)).astype(np.float32) if res else None
    fl = capi.FLAG_PRE_RELU | (capi.FLAG_POST_BATCHNORM if postbn else 0) | (capi.FLAG_BATCHNORM_ON_RESIDUAL if bnres else 0)
    y = conv_gpu(x, wb, out_channels=co, kernel=k, dtype=dtype, downsample=ds, in_pad=inp, out_pad=outp, flags=fl,
                 residual=residual, deep=True, backend=capi.BACKEND_DIRECT)
    _check(y, x, wb, dtype, out_channels=co, kernel=k, downsample=ds, in_pad=inp, out_pad=outp, act=fo.ACT_RELU,
           flags=(fo.POST_BATCHNORM if postbn else 0) | (fo.BATCHNORM_ON_RESIDUAL if bnres else 0), residual=residual, deep=True)


def test_deep_gemm_layer():
    """GEMM72 (deepgemmlayer.cpp:66-140): 1x1 conv on 1x1 spatial, 2048 -> 1000, + the misctests KAT."""
    rng = np.random.default_rng(72)
    x = rng.normal(size=(3, 2048, 1, 1)).astype(np.float32)
    wb = random_wb(rng, 2048, 1000, 1)
    y = conv_gpu(x, wb, out_channels=1000, kernel=1, deep=True, backend=capi.BACKEND_DIRECT)
    xs = half(x)
    ref = np.stack([fo.conv2d(xs[i], wb, 1000, 1, deep=True, prec=fo.FP16_STORE) for i in range(3)])
    assert_close_f16(y, ref, ulps=1.5)
    w = np.tile(np.array([1.0, -1.0], np.float32), 256)
    wb = np.concatenate([np.zeros(256, np.float32), np.tile(w, 256)])
    y = conv_gpu(np.ones((512, 1, 1), np.float32), wb, out_channels=256, kernel=1, deep=True)
    np.testing.assert_allclose(y, 0.0, atol=1e-3)


@pytest.mark.parametrize("dtype", [capi.F16, capi.F32])
@pytest.mark.parametrize("deep", [True, False])
def test_pooling(deep, dtype):
    """MaxPool4 (3x3 s2 pad1 pre-ReLU), GlobAvg70 (7x7 mean pre-ReLU) and the pooltests.cpp 2x2 cases."""
    c = ctx()
    rng = np.random.default_rng(9)
    order = capi.ORDER_DEEP if deep else capi.ORDER_SHALLOW
    dflag = capi.FLAG_DEEP if deep else 0
    cases = [dict(ch=64, size=16, pool=3, ds=2, pad=1, is_max=True, glob=False, relu=True),
             dict(ch=36, size=7, pool=7, ds=7, pad=0, is_max=False, glob=True, relu=True),
             dict(ch=23, size=10, pool=2, ds=2, pad=0, is_max=True, glob=False, relu=False),
             dict(ch=12, size=10, pool=2, ds=2, pad=0, is_max=False, glob=False, relu=False),
             dict(ch=8, size=8, pool=8, ds=8, pad=0, is_max=True, glob=True, relu=False)]
    for cs in cases:
        x = rng.normal(size=(2, cs["ch"], cs["size"], cs["size"])).astype(np.float32)
        op = capi.Pool2d(c, width=cs["size"], height=cs["size"], channels=cs["ch"], pool=cs["pool"], downsample=cs["ds"],
                         in_padding=cs["pad"], is_max=cs["is_max"], global_=cs["glob"],
                         flags=dflag | (capi.FLAG_PRE_RELU if cs["relu"] else 0))
        os_ = 1 if cs["glob"] else cs["size"] // cs["ds"]
        tin = c.tensor(cs["size"], cs["size"], cs["ch"], cs["pad"], order, dtype, 2)
        tout = c.tensor(os_, os_, cs["ch"], 0, order, dtype, 2)
        tin.write_chw(x)
        op.run(tin, tout)
        y = tout.read_chw()
        xs = half(x) if dtype == capi.F16 else x
        prec = fo.FP16_STORE if dtype == capi.F16 else fo.FP32
        ref = np.stack([fo.pool2d(xs[i], pool=cs["pool"], downsample=cs["ds"], in_pad=cs["pad"], is_max=cs["is_max"],
                                  global_=cs["glob"], act=fo.ACT_RELU if cs["relu"] else fo.ACT_NONE, prec=prec) for i in range(2)])
        if dtype == capi.F32 or cs["is_max"]:
            np.testing.assert_allclose(y, ref, rtol=1e-6, atol=1e-6)
        else:
            assert_close_f16(y, ref)
        for o in (tin, tout, op):
            o.destroy()


def test_large_grid_bandwidth_kernels_are_bit_identical(monkeypatch):
    """Large grids take other kernels than the small shapes of the oracle tests: the persistent ring pooling kernel (three staged
    windows per block, bulk copies), bulk-copy staging, the multi-row global pooling kernel and the pair-access element-wise kernel.
    Same arithmetic in the same order as the simple kernels (FYN_POOL_SIMPLE-free knobs FYN_POOL_NO_RING / FYN_POOL_NO_BULK), so
    the bits must agree; one image is also checked against the oracle."""
    c = ctx()
    rng = np.random.default_rng(77)
    for size, ch, batch, pool, ds, pad, is_max, relu, odd in [(112, 64, 24, 3, 2, 1, True, True, False), (56, 40, 48, 2, 2, 0, True, False, False),
                                                            (57, 24, 40, 3, 2, 1, True, True, True), (60, 32, 40, 2, 2, 0, False, True, False)]:
        x = rng.normal(size=(batch, ch, size, size)).astype(np.float32)
        so = size // ds
        op = capi.Pool2d(c, width=size, height=size, channels=ch, pool=pool, downsample=ds, in_padding=pad, is_max=is_max,
                         flags=capi.FLAG_DEEP | (capi.FLAG_PRE_RELU if relu else 0))
        tin = c.tensor(size, size, ch, pad, capi.ORDER_DEEP, capi.F16, batch)
        tout = c.tensor(so, so, ch, 0, capi.ORDER_DEEP, capi.F16, batch)
        tin.write_chw(x)
        outs = []
        for knobs in ((), ("FYN_POOL_NO_RING",), ("FYN_POOL_NO_RING", "FYN_POOL_NO_BULK")):
            for k in ("FYN_POOL_NO_RING", "FYN_POOL_NO_BULK"):
                monkeypatch.delenv(k, raising=False)
            for k in knobs:
                monkeypatch.setenv(k, "1")
            tout.write_chw(np.zeros((batch, ch, so, so), np.float32))
            op.run(tin, tout)
            outs.append(tout.read_chw())
        for k in ("FYN_POOL_NO_RING", "FYN_POOL_NO_BULK"):
            monkeypatch.delenv(k, raising=False)
        np.testing.assert_array_equal(outs[0], outs[1])
        np.testing.assert_array_equal(outs[0], outs[2])
        ref = fo.pool2d(half(x[-1]), pool=pool, downsample=ds, in_pad=pad, is_max=is_max, act=fo.ACT_RELU if relu else fo.ACT_NONE, prec=fo.FP16_STORE)
        if is_max:
            np.testing.assert_allclose(outs[0][-1], ref, rtol=1e-6, atol=1e-6)
        else:
            assert_close_f16(outs[0][-1], ref)
        for o in (tin, tout, op):
            o.destroy()
    # global average pooling over several rows of tiles per block, odd and even tile grids
    for size, ch, batch in [(7, 2048, 40), (7, 2048, 256), (7, 100, 9), (5, 512, 64), (5, 512, 700), (8, 260, 17)]:
        x = rng.normal(size=(batch, ch, size, size)).astype(np.float32)
        op = capi.Pool2d(c, width=size, height=size, channels=ch, pool=size, downsample=size, is_max=False, global_=True, flags=capi.FLAG_DEEP | capi.FLAG_PRE_RELU)
        tin = c.tensor(size, size, ch, 0, capi.ORDER_DEEP, capi.F16, batch)
        tout = c.tensor(1, 1, ch, 0, capi.ORDER_DEEP, capi.F16, batch)
        tin.write_chw(x)
        op.run(tin, tout)
        y = tout.read_chw()
        ref = np.maximum(half(x), 0).astype(np.float64).mean(axis=(2, 3)).reshape(y.shape)
        assert_close_f16(y, half(ref), extra_abs=1e-4)
        # the warp-per-tile-group kernel (default for small tiles) adds in the order of the block kernel: same bits
        monkeypatch.setenv("FYN_POOL_GLOBAL_BLOCK", "1")
        tout.write_chw(np.zeros_like(y))
        op.run(tin, tout)
        monkeypatch.delenv("FYN_POOL_GLOBAL_BLOCK")
        np.testing.assert_array_equal(tout.read_chw(), y)
        for o in (tin, tout, op):
            o.destroy()
    # batch-norm on pair-aligned deep tensors (even width, no padding) against the texel-wise kernel's shapes (odd width)
    for size, ch, batch in [(56, 256, 6), (28, 64, 5), (24, 12, 3)]:
        x = rng.normal(size=(batch, ch, size, size)).astype(np.float32)
        sb = np.concatenate([rng.uniform(0.5, 1.5, ch), rng.uniform(-0.5, 0.5, ch)]).astype(np.float32)
        op = capi.BatchNorm(c, sb, width=size, height=size, channels=ch, flags=capi.FLAG_DEEP | capi.FLAG_PRE_RELU)
        tin = c.tensor(size, size, ch, 0, capi.ORDER_DEEP, capi.F16, batch)
        tout = c.tensor(size, size, ch, 0, capi.ORDER_DEEP, capi.F16, batch)
        tin.write_chw(x)
        op.run(tin, tout)
        y = tout.read_chw()
        ref = half(np.maximum(half(x), 0) * sb[:ch, None, None] + sb[ch:, None, None])
        assert_close_f16(y, ref)
        for o in (tin, tout, op):
            o.destroy()


@pytest.mark.parametrize("dtype", [capi.F16, capi.F32])
def test_batchnorm_and_sigmoid(dtype):
    """BN2 (shallow, outputPadding 1, no activation), deep BN (resnet50.cpp BN5..BN66), SigmoidLayer."""
    c = ctx()
    rng = np.random.default_rng(3)
    prec = fo.FP16_STORE if dtype == capi.F16 else fo.FP32
    # the larger cases walk the 2 / 4 / 8 texels-per-thread variants of the fp16 plane-chunk kernel
    for deep, ch, size, outp in [(False, 3, 20, 1), (True, 64, 9, 0), (True, 31, 6, 0), (False, 23, 6, 0),
                                 (True, 8, 20, 0), (False, 7, 30, 1), (True, 20, 40, 0), (True, 12, 57, 1)]:
        order = capi.ORDER_DEEP if deep else capi.ORDER_SHALLOW
        x = rng.uniform(-10, 10, (2, ch, size, size + 3)).astype(np.float32)
        sb = rng.uniform(-2, 2, 2 * ch).astype(np.float32)
        op = capi.BatchNorm(c, sb, width=size + 3, height=size, channels=ch, out_padding=outp,
                            flags=capi.FLAG_DEEP if deep else 0)
        tin = c.tensor(size + 3, size, ch, 0, order, dtype, 2)
        tout = c.tensor(size + 3, size, ch, outp, order, dtype, 2)
        tin.write_chw(x)
        op.run(tin, tout)
        y = tout.read_chw()
        xs = half(x) if dtype == capi.F16 else x
        ref = np.stack([fo.batchnorm(xs[i], sb, deep=deep, prec=prec) for i in range(2)])
        if dtype == capi.F32:
            np.testing.assert_allclose(y, ref, rtol=1e-6, atol=1e-6)
        else:
            assert_close_f16(y, ref)
        raw = tout.download()
        if outp and not deep:
            assert np.all(raw[:, :, 0] == 0) and np.all(raw[:, :, :, 0] == 0)
        for o in (tin, tout, op):
            o.destroy()
    # sigmoid on a 3-channel plane: lane 3 becomes sigmoid(0) = 0.5 like the reference
    h, w = 12, 16
    x = rng.normal(size=(3, h, w)).astype(np.float32) * 3
    op = capi.Sigmoid(c, width=w, height=h, channels=3)
    tin = c.tensor(w, h, 3, 0, capi.ORDER_SHALLOW, dtype)
    tout = c.tensor(w, h, 3, 0, capi.ORDER_SHALLOW, dtype)
    tin.write_chw(x)
    op.run(tin, tout)
    y = tout.read_chw()
    xs = half(x) if dtype == capi.F16 else x
    ref = fo.sigmoid(xs, prec=prec)
    if dtype == capi.F32:
        np.testing.assert_allclose(y, ref, rtol=1e-5, atol=1e-6)
    else:
        assert_close_f16(y, ref, ulps=1.5)
    host = tout.download()
    assert host.shape == (1, 1, h, w, 4)
    np.testing.assert_allclose(host[0, 0, :, :, 3], 0.5)
    np.testing.assert_allclose(host[0, 0, :, :, :3], fo.download_shallow(y, 0.5)[0][..., :3], atol=0)


def test_upload_download_layouts():
    """UploadLayer / DownloadLayer / DeepDownloadLayer host orders (uploadlayer.cpp:360-380,
    downloadlayer.cpp:257-283, deepdownloadlayer.cpp:136-160) and the CHW dump format."""
    c = ctx()
    rng = np.random.default_rng(1)
    # upload into an fp16 RGBA plane with padding (convert kernel path), batch 2
    img = rng.random((2, 9, 13, 3), dtype=np.float32)
    t = c.tensor(13, 9, 3, 1, capi.ORDER_SHALLOW, capi.F16, 2)
    t.upload(img)
    got = t.read_chw()
    np.testing.assert_array_equal(got, half(np.stack([fo.upload_hwc(img[i]) for i in range(2)])))
    raw = t.download()
    assert raw.shape == (2, 1, 11, 15, 4)
    np.testing.assert_array_equal(raw[0, 0], fo.pack_shallow(got[0], 1)[0])
    t.destroy()
    # deep layout download == oracle pack_deep (texel order incl. padding)
    x = half(rng.normal(size=(23, 6, 5)))
    t = c.tensor(5, 6, 23, 1, capi.ORDER_DEEP, capi.F16)
    t.write_chw(x)
    np.testing.assert_array_equal(t.download()[0], fo.pack_deep(x, 1))
    np.testing.assert_array_equal(t.read_chw(), x)
    t.destroy()
    # 1x1x1000 logits: deep download is channel order (cpubuffer.cpp:131-142)
    x = rng.normal(size=(1000, 1, 1)).astype(np.float32)
    t = c.tensor(1, 1, 1000, 0, capi.ORDER_DEEP, capi.F32)
    t.write_chw(x)
    d = t.download()
    assert d.shape == (1, 14, 18, 4)
    np.testing.assert_array_equal(d.reshape(-1)[:1000], x.reshape(-1))
    t.destroy()


def test_byte_upload_and_download():
    """UBYTE upload (gpu/uploadlayer.cpp:51-66: an 8-bit normalised texture, i.e. value / 255) and the RGBA8 download
    ((uint8)(v * 255) like samples/desktop/stylenet.cpp:52-62, after a clamp): exact against numpy, on the vectorised paths
    (RGB32F upload texture, fp16 RGBA download) and on the generic ones."""
    c = ctx()
    rng = np.random.default_rng(3)
    for (w, h, batch) in [(16, 9, 1), (1524, 8, 1), (13, 7, 2)]:
        img = rng.integers(0, 256, size=(batch, h, w, 3), dtype=np.uint8)
        want = img.astype(np.float32) / np.float32(255.0)
        # the RGB32F upload texture of the float path (packing 3, no padding): same texels as uploading img / 255 as floats
        t8 = c.tensor(w, h, 3, 0, capi.ORDER_SHALLOW, capi.F32, batch, packing=3)
        tf = c.tensor(w, h, 3, 0, capi.ORDER_SHALLOW, capi.F32, batch, packing=3)
        t8.upload_u8(img)
        tf.upload(want)
        c.stream_sync()
        np.testing.assert_array_equal(t8.read_chw(), tf.read_chw())
        np.testing.assert_array_equal(np.asarray(t8.read_chw()).reshape(batch, 3, h, w), want.transpose(0, 3, 1, 2))
        t8.destroy()
        tf.destroy()
        # generic path: fp16 RGBA plane with padding
        t = c.tensor(w, h, 3, 1, capi.ORDER_SHALLOW, capi.F16, batch)
        t.upload_u8(img)
        c.stream_sync()
        np.testing.assert_array_equal(np.asarray(t.read_chw()).reshape(batch, 3, h, w), half(want).transpose(0, 3, 1, 2))
        t.destroy()
    # download: fp16 RGBA (vectorised when the texel count allows it), fp32, values outside [0, 1] clamp
    for (w, h, ch, pad, dt) in [(16, 8, 4, 0, capi.F16), (13, 7, 3, 1, capi.F16), (12, 5, 4, 0, capi.F32)]:
        x = rng.uniform(-0.2, 1.2, size=(ch, h, w)).astype(np.float32)
        if dt == capi.F16:
            x = half(x)
        t = c.tensor(w, h, ch, pad, capi.ORDER_SHALLOW, dt)
        t.write_chw(x)
        f = t.download()
        b = t.download_u8()
        assert b.shape == f.shape and b.dtype == np.uint8
        np.testing.assert_array_equal(b, (np.clip(f, 0.0, 1.0) * np.float32(255.0)).astype(np.uint8))
        t.destroy()


def test_reference_kats_on_gpu():
    """The reference's own layer tests replayed on the CUDA path (convlayertests.cpp:159-305, networktests.cpp)."""
    from test_oracle_kat import antisym_kernel, padded_convolution, stack_convolution
    for (w, h, ci, co) in [(64, 64, 4, 4), (128, 80, 4, 8), (56, 56, 64, 64), (256, 128, 12, 4)]:
        for ds in (1, 2):
            y = conv_gpu(np.ones((ci, h, w), np.float32), stack_convolution(0.0, [[1.0]], ci, co), out_channels=co,
                         kernel=1, downsample=ds)
            np.testing.assert_allclose(y, ci, atol=1e-3)
    for k in (3, 5, 7, 9):
        for ds in (1, 2):
            wb = stack_convolution(0.0, antisym_kernel(k), 4, 8)
            y = conv_gpu(np.ones((4, 80, 128), np.float32), wb, out_channels=8, kernel=k, downsample=ds)
            np.testing.assert_allclose(y, 0.0, atol=1e-3)   # clamp-to-edge
            if k < 9:
                pad = (k - 1) // 2
                x = np.ones((12, 64, 48), np.float32)
                wb = stack_convolution(0.0, antisym_kernel(k), 12, 8)
                ref = padded_convolution(np.pad(x, ((0, 0), (pad, pad), (pad, pad))), wb, 8, k, 12, down=ds)
                y = conv_gpu(x, wb, out_channels=8, kernel=k, downsample=ds, in_pad=pad, deep=True)
                np.testing.assert_allclose(y, ref, atol=1e-3)
    # upload -> conv3x3 (4->8, +-1 filter) -> download on all-ones 32x32 == 0 exactly
    c = ctx()
    tin = c.tensor(32, 32, 4, 0, capi.ORDER_SHALLOW, capi.F32)
    tout = c.tensor(32, 32, 8, 0, capi.ORDER_SHALLOW, capi.F16)
    tin.upload(np.ones((32, 32, 4), np.float32))
    op = capi.Conv2d(c, stack_convolution(0.0, antisym_kernel(3), 4, 8), width=32, height=32, in_channels=4,
                     out_channels=8, kernel=3)
    op.run(tin, tout)
    assert np.all(tout.download() == 0.0)


def test_error_behaviour():
    """Shape mismatches are refused with a message (C++ wrapper -> FynException), never silently run."""
    c = ctx()
    op = capi.Conv2d(c, np.zeros(8 + 9 * 4 * 8, np.float32), width=16, height=16, in_channels=4, out_channels=8, kernel=3)
    tin = c.tensor(16, 16, 4)
    bad = c.tensor(15, 16, 8)
    with pytest.raises(capi.FynError, match="output tensor mismatch"):
        op.run(tin, bad)
    with pytest.raises(capi.FynError, match="kernel"):
        capi.Conv2d(c, np.zeros(1000, np.float32), width=16, height=16, in_channels=4, out_channels=8, kernel=4)
    with pytest.raises(capi.FynError, match="fractional"):
        capi.Conv2d(c, np.zeros(1000, np.float32), width=16, height=16, in_channels=4, out_channels=8, kernel=3,
                    fractional=True, source_step=0.5, flags=capi.FLAG_DEEP)
    n0 = c.launch_count()
    good = c.tensor(16, 16, 8)
    op.run(tin, good)
    assert c.launch_count() == n0 + 1


@pytest.mark.parametrize("backend", [capi.BACKEND_DIRECT, capi.BACKEND_TC])
@pytest.mark.parametrize("dtype", [capi.F16, capi.F32])
def test_conv_fused_sigmoid_epilogue(backend, dtype):
    """fyn_conv2d_set_epilogue(SIGMOID): one kernel must give exactly what the conv layer followed by the sigmoid
    layer gives (StyleNet deconv3 -> sigmoid, fractional 9x9 12->3), on both kernel families."""
    if backend == capi.BACKEND_TC and dtype == capi.F32:
        pytest.skip("the tcgen05 family stores fp16")
    c = ctx()
    rng = np.random.default_rng(17)
    h, w, ci, co = 24, 136, 12, 3
    x = rng.normal(size=(ci, h, w)).astype(np.float32)
    wb = random_wb(rng, ci, co, 9)
    kw = dict(width=w, height=h, in_channels=ci, out_channels=co, kernel=9, flags=capi.FLAG_PRE_RELU, source_step=0.5,
              fractional=True, backend=backend)
    conv = capi.Conv2d(c, wb, **kw)
    ow, oh = conv.out_width, conv.out_height
    tin = c.tensor(w, h, ci, 0, capi.ORDER_SHALLOW, dtype)
    tmid = c.tensor(ow, oh, co, 0, capi.ORDER_SHALLOW, dtype)
    t2 = c.tensor(ow, oh, co, 0, capi.ORDER_SHALLOW, dtype)
    t1 = c.tensor(ow, oh, co, 0, capi.ORDER_SHALLOW, dtype)
    tin.write_chw(x)
    sig = capi.Sigmoid(c, width=ow, height=oh, channels=co)
    conv.run(tin, tmid)
    sig.run(tmid, t2)
    want = t2.download().copy()            # raw texels including the padding lane (sigmoid(0) = 0.5)
    conv.set_epilogue(capi.EPILOGUE_SIGMOID)
    conv.run(tin, t1)
    got = t1.download()
    assert conv.backend == backend
    np.testing.assert_array_equal(got, want)
    conv.set_epilogue(capi.EPILOGUE_NONE)
    conv.run(tin, t1)
    np.testing.assert_array_equal(t1.download(), tmid.download())
    for o in (tin, tmid, t1, t2, conv, sig):
        o.destroy()
