"""ctypes/numpy front end of the CPU oracle (oracle/fyn_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (fyusenet_b200/) never imports it.

Besides thin wrappers of the C layer functions this module restates, independently of the
product's C++ host engine, the two sample networks of the reference:

  * StyleNet 3x3 / 9x9  -- /root/reference/samples/samplenetworks/stylenet3x3.cpp:114-234,
                           stylenet9x9.cpp:120-273 (topology), :41-56 (weight offsets)
  * ResNet-50           -- /root/reference/samples/samplenetworks/resnet50.cpp:200-516 (topology),
                           :539-677 (weight offsets / file order)

and the synthetic weight / image generators of SURVEY.md section 8(d) (the reference's data/*.dat
are git-LFS stubs).  Parity status: layer level pinned by the reference's unit-test KATs
(tests/test_oracle_kat.py); whole-network outputs: parity unpinned.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libfyn_oracle.so"

# layer flags (fyusenet/base/layerflags.h:33-53)
RESIDUAL_INPUT, RELU_ON_RESIDUAL, BATCHNORM_ON_RESIDUAL, POST_BATCHNORM, DEEP = 1, 2, 4, 8, 16
ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_CLIP = 0, 1, 2, 3
FP32, FP16_STORE, FP16_BLEND = 0, 1, 2
Q1_FRAC3_ASYM, Q2_FRAC_ACT_FIRST, Q7_MAXPOOL3_COL, QUIRKS_REFERENCE = 1, 2, 4, 31


class _Act(C.Structure):
    _fields_ = [("type", C.c_int), ("leak", C.c_float), ("lo", C.c_float), ("hi", C.c_float)]


class _Conv(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("inChannels", C.c_int), ("outChannels", C.c_int),
                ("kernel", C.c_int), ("downsample", C.c_int), ("dilation", C.c_int),
                ("inPadding", C.c_int), ("outPadding", C.c_int), ("flags", C.c_uint), ("act", _Act),
                ("sourceStep", C.c_float), ("fractional", C.c_int), ("deep", C.c_int),
                ("quirks", C.c_int), ("prec", C.c_int)]


class _Pool(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("channels", C.c_int), ("poolX", C.c_int),
                ("poolY", C.c_int), ("downsample", C.c_int), ("inPadding", C.c_int), ("isMax", C.c_int),
                ("global_", C.c_int), ("act", _Act), ("quirks", C.c_int), ("prec", C.c_int)]


def build(force: bool = False) -> Path:
    """Compile the C oracle (gcc, a few seconds)."""
    src = _HERE / "fyn_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < max(
            src.stat().st_mtime, (_HERE / "fyn_oracle.h").stat().st_mtime):
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.fyo_half_round.restype = C.c_float
        _lib.fyo_half_round.argtypes = [C.c_float]
        _lib.fyo_half_trunc.restype = C.c_float
        _lib.fyo_half_trunc.argtypes = [C.c_float]
        _lib.fyo_half_trunc_bits.restype = C.c_uint16
        _lib.fyo_half_trunc_bits.argtypes = [C.c_float]
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _act(act=ACT_NONE, leak=0.0, lo=0.0, hi=0.0):
    return _Act(int(act), float(leak), float(lo), float(hi))


# ------------------------------------------------------------------------------------------------
# layouts
# ------------------------------------------------------------------------------------------------

def deep_tiling(channels: int):
    tx, ty = C.c_int(), C.c_int()
    lib().fyo_deep_tiling(int(channels), C.byref(tx), C.byref(ty))
    return tx.value, ty.value


def deep_texture_size(channels, w, h, pad):
    tw, th = C.c_int(), C.c_int()
    lib().fyo_deep_texture_size(int(channels), int(w), int(h), int(pad), C.byref(tw), C.byref(th))
    return tw.value, th.value


def pack_deep(chw, pad):
    chw = _f32(chw)
    c, h, w = chw.shape
    tw, th = deep_texture_size(c, w, h, pad)
    out = np.zeros((th, tw, 4), np.float32)
    lib().fyo_pack_deep(_fp(chw), c, h, w, int(pad), _fp(out))
    return out


def unpack_deep(tex, c, h, w, pad):
    tex = _f32(tex)
    out = np.zeros((c, h, w), np.float32)
    lib().fyo_unpack_deep(_fp(tex), int(c), int(h), int(w), int(pad), _fp(out))
    return out


def pack_shallow(chw, pad):
    chw = _f32(chw)
    c, h, w = chw.shape
    out = np.zeros(((c + 3) // 4, h + 2 * pad, w + 2 * pad, 4), np.float32)
    lib().fyo_pack_shallow(_fp(chw), c, h, w, int(pad), _fp(out))
    return out


def unpack_shallow(planes, c, h, w, pad):
    planes = _f32(planes)
    out = np.zeros((c, h, w), np.float32)
    lib().fyo_unpack_shallow(_fp(planes), int(c), int(h), int(w), int(pad), _fp(out))
    return out


def half_round(a):
    return np.asarray(a, np.float32).astype(np.float16).astype(np.float32)


def half_trunc(a):
    """fp16 truncation of gpu/floatconversion.cpp:44-58, vectorised, checked against the C version in tests."""
    a = np.ascontiguousarray(a, np.float32)
    f = a.view(np.uint32)
    sign = ((f >> 16) & 0x8000).astype(np.uint32)
    e = ((f >> 23) & 0xFF).astype(np.int32) - 127
    man = (f & 0x007FFFFF).astype(np.uint32)
    out = np.zeros(a.shape, np.uint32)
    den = (e >= -24) & (e < -14)
    sh = np.clip(-e - 14, 0, 31).astype(np.uint32)
    sh2 = np.clip(-e - 1, 0, 31).astype(np.uint32)
    out = np.where(den, (0x0400 >> sh) + (man >> sh2), out)
    nrm = (e >= -14) & (e <= 15)
    out = np.where(nrm, ((e + 15).astype(np.uint32) << 10) + (man >> 13), out)
    out = np.where((e > 15) & (e < 128), 0x7C00, out)
    out = np.where(e >= 128, 0x7C00 + (man >> 13), out)
    bits = (out | sign).astype(np.uint16)
    return bits.view(np.float16).astype(np.float32)


# ------------------------------------------------------------------------------------------------
# layers (CHW float32 in / out, the reference's dump format base/layerbase.h:160-172)
# ------------------------------------------------------------------------------------------------

def conv2d(x, wb, out_channels, kernel, *, downsample=1, dilation=1, in_pad=0, out_pad=0, flags=0,
           act=ACT_NONE, leak=0.0, clip=(0.0, 0.0), source_step=1.0, fractional=False, deep=False,
           residual=None, quirks=QUIRKS_REFERENCE, prec=FP32):
    x = _f32(x)
    wb = _f32(wb)
    ci, h, w = x.shape
    p = _Conv(w, h, ci, int(out_channels), int(kernel), int(downsample), int(dilation), int(in_pad),
              int(out_pad), int(flags), _act(act, leak, clip[0], clip[1]), float(source_step),
              int(bool(fractional)), int(bool(deep)), int(quirks), int(prec))
    need = out_channels + kernel * kernel * ci * out_channels + (2 * out_channels if flags & POST_BATCHNORM else 0)
    if wb.size < need:
        raise ValueError(f"weight blob too small: {wb.size} < {need}")
    wo, ho = C.c_int(), C.c_int()
    lib().fyo_conv2d_outdims(C.byref(p), C.byref(wo), C.byref(ho))
    out = np.zeros((out_channels, ho.value, wo.value), np.float32)
    res = None
    if residual is not None:
        res = _f32(residual)
        if res.shape != out.shape:
            raise ValueError(f"residual shape {res.shape} != output shape {out.shape}")
        p.flags |= RESIDUAL_INPUT
    rc = lib().fyo_conv2d(C.byref(p), _fp(x), _fp(wb), _fp(res) if res is not None else None, _fp(out))
    if rc != 0:
        raise RuntimeError(f"fyo_conv2d failed rc={rc}")
    return out


def pool2d(x, *, pool=2, downsample=2, in_pad=0, is_max=True, global_=False, act=ACT_NONE, leak=0.0,
           quirks=QUIRKS_REFERENCE, prec=FP32):
    x = _f32(x)
    c, h, w = x.shape
    px = py = int(pool)
    p = _Pool(w, h, c, px, py, int(downsample), int(in_pad), int(bool(is_max)), int(bool(global_)),
              _act(act, leak), int(quirks), int(prec))
    if global_:
        out = np.zeros((c, 1, 1), np.float32)
    else:
        out = np.zeros((c, h // downsample, w // downsample), np.float32)
    rc = lib().fyo_pool2d(C.byref(p), _fp(x), _fp(out))
    if rc != 0:
        raise RuntimeError(f"fyo_pool2d failed rc={rc}")
    return out


def batchnorm(x, scale_bias, *, deep=False, act=ACT_NONE, prec=FP32):
    x = _f32(x)
    sb = _f32(scale_bias)
    c, h, w = x.shape
    out = np.zeros_like(x)
    a = _act(act)
    lib().fyo_batchnorm(_fp(x), c, h, w, _fp(sb), int(bool(deep)), C.byref(a), int(prec), _fp(out))
    return out


def sigmoid(x, *, act=ACT_NONE, prec=FP32):
    x = _f32(x)
    out = np.zeros_like(x)
    a = _act(act)
    lib().fyo_sigmoid(_fp(x), C.c_size_t(x.size), C.byref(a), int(prec), _fp(out))
    return out


def scale(x, *, up=(1, 1), down=(1, 1), linear=False, in_pad=0, deep=False, act=ACT_NONE, lo=0.0, hi=0.0, prec=FP32):
    """ScaleLayer / DeepScaleLayer (scalelayer.cpp:40-60, scaling.frag); all factors 1 = PADDING2D / RELU / CLIP."""
    x = _f32(x)
    c, h, w = x.shape
    wo, ho = C.c_int(), C.c_int()
    lib().fyo_scale_outdims(w, h, up[0], up[1], down[0], down[1], C.byref(wo), C.byref(ho))
    out = np.zeros((c, ho.value, wo.value), np.float32)
    a = _act(act, 0.0, lo, hi)
    rc = lib().fyo_scale(_fp(x), c, h, w, int(in_pad), int(bool(deep)), up[0], up[1], down[0], down[1], int(bool(linear)),
                         C.byref(a), int(prec), _fp(out))
    if rc != 0:
        raise RuntimeError(f"fyo_scale failed rc={rc}")
    return out


ARITH_ADD, ARITH_SUB, ARITH_MUL, ARITH_DIV = 0, 1, 2, 3
QUIRK_DW_BN_OFFSET = 8
QUIRK_TRANS2X2_NEXT = 16


def transconv(x, wb, out_channels, kernel, *, in_pad=0, post_bn=False, quirks=QUIRKS_REFERENCE, act=ACT_NONE, leak=0.0, prec=FP32, deep=False):
    """Stride-2 transpose convolution; shallow layout (transconvlayerbase_vanilla.cpp, convtrans{2x2,3x3}_stride2.frag) or, with
    deep=True, the deep-tiled variant (deeptransconvlayerbase.cpp, deeptransconv{2x2,3x3}_stride2.*: other tap alignment, zero
    outside the image, fp16-truncated weights with fp16 storage)."""
    x = _f32(x)
    wb = _f32(wb)
    ci, h, w = x.shape
    assert wb.size >= out_channels * (1 + kernel * kernel * ci) + (2 * out_channels if post_bn else 0)
    out = np.zeros((out_channels, 2 * h, 2 * w), np.float32)
    a = _act(act, leak)
    if deep:
        rc = lib().fyo_transconv_deep(_fp(x), ci, h, w, int(in_pad), int(out_channels), int(kernel), int(bool(post_bn)), _fp(wb), C.byref(a), int(prec), _fp(out))
        if rc != 0:
            raise RuntimeError(f"fyo_transconv_deep failed rc={rc}")
        return out
    rc = lib().fyo_transconv(_fp(x), ci, h, w, int(in_pad), int(out_channels), int(kernel), int(bool(post_bn)), int(quirks), _fp(wb),
                             C.byref(a), int(prec), _fp(out))
    if rc != 0:
        raise RuntimeError(f"fyo_transconv failed rc={rc}")
    return out


def dwconv3x3(x, wb, *, downsample=1, dilation=1, in_pad=0, deep=False, post_bn=False, quirks=0, act=ACT_NONE, leak=0.0, prec=FP32,
              multiplier=1, residual=None, relu_on_residual=False, bn_on_residual=False):
    """Depthwise 3x3 convolution (convlayer_dw_3x3_vanilla.cpp / deepdwconvlayer3x3.cpp); wb = bias[Co], W[C][3][3][mult], (bn) with
    Co = C * mult; output channel m * C + c = input channel c through multiplier m (deep layers only for mult > 1)."""
    x = _f32(x)
    wb = _f32(wb)
    c, h, w = x.shape
    co = c * multiplier
    assert wb.size >= co + c * 9 * multiplier + (2 * co if post_bn and not (quirks & QUIRK_DW_BN_OFFSET and not deep) else 0)
    out = np.zeros((co, h // downsample, w // downsample), np.float32)
    a = _act(act, leak)
    res = None
    if residual is not None:
        res = _f32(residual)
        assert res.shape == out.shape
    rc = lib().fyo_dwconv3x3_ex(_fp(x), c, h, w, int(in_pad), int(bool(deep)), int(downsample), int(dilation), int(bool(post_bn)),
                                int(quirks), _fp(wb), C.byref(a), int(prec), int(multiplier), _fp(res) if res is not None else None,
                                (1 if relu_on_residual else 0) | (2 if bn_on_residual else 0), _fp(out))
    if rc != 0:
        raise RuntimeError(f"fyo_dwconv3x3 failed rc={rc}")
    return out


def arith(a, b, op, *, act=ACT_NONE, prec=FP32):
    """AddSubLayer (b = array, ADD / SUB) or SingletonArithmeticLayer (b = scalar)."""
    a = _f32(a)
    out = np.zeros_like(a)
    ac = _act(act)
    if np.isscalar(b):
        rc = lib().fyo_arith(_fp(a), None, C.c_size_t(a.size), int(op), C.c_float(float(b)), C.byref(ac), int(prec), _fp(out))
    else:
        b = _f32(b)
        assert b.shape == a.shape
        rc = lib().fyo_arith(_fp(a), _fp(b), C.c_size_t(a.size), int(op), C.c_float(0.0), C.byref(ac), int(prec), _fp(out))
    if rc != 0:
        raise RuntimeError(f"fyo_arith failed rc={rc}")
    return out


def concat(inputs, *, act=ACT_NONE, prec=FP32):
    """ConcatLayer / DeepConcatLayer: channels back to back (concatlayer.cpp:60-75), activation on every input."""
    return np.concatenate([scale(x, act=act, prec=prec) for x in inputs], axis=0)


def rgb2bgr(x, *, prec=FP32):
    x = _f32(x)
    c, h, w = x.shape
    out = np.zeros_like(x)
    lib().fyo_rgb2bgr(_fp(x), c, h, w, int(prec), _fp(out))
    return out


def upload_hwc(hwc):
    """gpu/uploadlayer.cpp:360-380: host [H][W][C] float32 -> C-channel float32 texture (CHW here)."""
    hwc = _f32(hwc)
    h, w, c = hwc.shape
    out = np.zeros((c, h, w), np.float32)
    lib().fyo_upload_hwc_to_chw(_fp(hwc), c, h, w, _fp(out))
    return out


def download_shallow(chw, fill=0.0):
    """gpu/downloadlayer.cpp:257-283: -> host [planes][H][W][4] float32."""
    chw = _f32(chw)
    c, h, w = chw.shape
    out = np.zeros(((c + 3) // 4, h, w, 4), np.float32)
    lib().fyo_download_shallow(_fp(chw), c, h, w, C.c_float(fill), _fp(out))
    return out


# ------------------------------------------------------------------------------------------------
# StyleNet (3x3 / 9x9)
# ------------------------------------------------------------------------------------------------

def stylenet_layers(ksize: int):
    """Layer table in weight-FILE order: (name, kernel, cin, cout).
    stylenet9x9.cpp:41-56 / stylenet3x3.cpp:41-50: conv1..3, deconv1..3, then the res blocks."""
    nres = 5 if ksize == 9 else 2
    layers = [("conv1", ksize, 3, 12), ("conv2", 3, 12, 20), ("conv3", 3, 20, 40),
              ("deconv1", 3, 40, 20), ("deconv2", 3, 20, 12), ("deconv3", ksize, 12, 3)]
    for r in range(1, nres + 1):
        layers += [(f"res{r}_1", 3, 40, 40), (f"res{r}_2", 3, 40, 40)]
    return layers


def stylenet_offsets(ksize: int):
    offs, o = {}, 0
    for name, k, ci, co in stylenet_layers(ksize):
        offs[name] = o
        o += co + co * k * k * ci
    offs["_total"] = o
    return offs


def stylenet_synthetic_weights(ksize: int, seed: int | None = None) -> np.ndarray:
    """He-normal conv weights, U(-0.05,0.05) biases; numpy PCG64 generator
    (seed 112 for 3x3, 9112 for 9x9).  The second conv of each residual block is scaled by 0.5
    to keep 5 stacked blocks inside the fp16 range."""
    if seed is None:
        seed = 9112 if ksize == 9 else 112
    rng = np.random.default_rng(seed)
    offs = stylenet_offsets(ksize)
    blob = np.zeros(offs["_total"], np.float32)
    for name, k, ci, co in stylenet_layers(ksize):
        o = offs[name]
        blob[o:o + co] = rng.uniform(-0.05, 0.05, co)
        w = rng.normal(0.0, np.sqrt(2.0 / (k * k * ci)), (co, k, k, ci))
        if name.startswith("res") and name.endswith("_2"):
            w *= 0.5
        blob[o + co:o + co + w.size] = w.reshape(-1)
    return blob


def synthetic_image(h: int, w: int, index: int = 0) -> np.ndarray:
    """float32 [H][W][3] in [0,1) (equivalent of the sample's uint8/255, samples/desktop/stylenet.cpp:45-47)."""
    rng = np.random.default_rng(1000 + index)
    return rng.random((h, w, 3), dtype=np.float32)


def stylenet_forward(weights, img_hwc, ksize=9, *, prec=FP32, quirks=QUIRKS_REFERENCE, dump=None):
    """upload -> conv1 -> conv2 -> conv3 -> res blocks -> deconv1..3 -> sigmoid -> download.
    Returns host RGBA float32 [H][W][4] (alpha = sigmoid(0) = 0.5, see SURVEY A.5: compare RGB only).
    `dump` (dict) receives every layer's CHW output keyed by layer name."""
    offs = stylenet_offsets(ksize)
    nres = 5 if ksize == 9 else 2
    wts = _f32(weights)

    def wb(name):
        return wts[offs[name]:]

    def rec(name, t):
        if dump is not None:
            dump[name] = t
        return t

    common = dict(prec=prec, quirks=quirks)
    x = rec("upload", upload_hwc(img_hwc))
    x = rec("conv1", conv2d(x, wb("conv1"), 12, ksize, act=ACT_RELU, **common))
    x = rec("conv2", conv2d(x, wb("conv2"), 20, 3, downsample=2, act=ACT_RELU, **common))
    x = rec("conv3", conv2d(x, wb("conv3"), 40, 3, downsample=2, act=ACT_RELU, **common))
    for r in range(1, nres + 1):
        # res2_1 has no prefix activation; res1_2 applies ReLU to the residual (stylenet9x9.cpp:149-163)
        a1 = ACT_NONE if r == 2 else ACT_RELU
        y = rec(f"res{r}_1", conv2d(x, wb(f"res{r}_1"), 40, 3, act=a1, **common))
        fl = RESIDUAL_INPUT | (RELU_ON_RESIDUAL if r == 1 else 0)
        x = rec(f"res{r}_2", conv2d(y, wb(f"res{r}_2"), 40, 3, act=ACT_RELU, flags=fl, residual=x, **common))
    x = rec("deconv1", conv2d(x, wb("deconv1"), 20, 3, downsample=2, source_step=0.5, fractional=True, **common))
    x = rec("deconv2", conv2d(x, wb("deconv2"), 12, 3, downsample=2, source_step=0.25, fractional=True,
                              act=ACT_RELU, **common))
    x = rec("deconv3", conv2d(x, wb("deconv3"), 3, ksize, source_step=0.5, fractional=True, act=ACT_RELU, **common))
    x = rec("sigmoid", sigmoid(x, prec=prec))
    return download_shallow(x, fill=0.5)[0]


# ------------------------------------------------------------------------------------------------
# ResNet-50
# ------------------------------------------------------------------------------------------------

_STAGES = [(64, 256, 3, 56), (128, 512, 4, 28), (256, 1024, 6, 14), (512, 2048, 3, 7)]


def resnet50_layers():
    """Layer table in reference layer-NUMBER order (resnet50.cpp:200-423).  Each entry is a dict:
    no, kind in {bn, conv, maxpool, gap, gemm}, and for convs: k, cin, cout, size (input HxW),
    ds, in_pad, out_pad, act, post_bn, residual (layer number or None), bn_on_res, input (layer no)."""
    L = []
    L.append(dict(no=2, kind="bn", c=3, size=224, deep=False, out_pad=1, input=0))
    L.append(dict(no=3, kind="conv", k=7, cin=3, cout=64, size=224, ds=2, in_pad=1, out_pad=1, act=ACT_NONE,
                  post_bn=True, residual=None, bn_on_res=False, input=2))
    L.append(dict(no=4, kind="maxpool", c=64, size=112, input=3))
    L.append(dict(no=5, kind="bn", c=64, size=56, deep=True, out_pad=0, input=4))
    no = 6
    prev = 5        # layer feeding the first 1x1 of the block
    shortcut = None  # residual source
    for si, (mid, out, nblocks, size) in enumerate(_STAGES):
        cin = 64 if si == 0 else _STAGES[si - 1][1]
        for b in range(nblocks):
            last = (b == nblocks - 1)  # Conv17/33/57/69: postBN + BN on residual
            if b == 0:
                # first block of a stage: 1x1a (at the input resolution), projection shortcut, 3x3 (stride 2
                # for stages > 0), 1x1c + residual(shortcut).  Numbering: a, shortcut, 3x3, c.
                in_size = size if si == 0 else size * 2
                ds = 1 if si == 0 else 2
                if si == 0:
                    a_in = prev
                    L.append(dict(no=no, kind="conv", k=1, cin=cin, cout=mid, size=in_size, ds=1, in_pad=0, out_pad=1,
                                  act=ACT_RELU, post_bn=True, residual=None, bn_on_res=False, input=a_in))
                    a_no = no
                    no += 1
                else:
                    a_no = pending_a  # noqa: F821  (created as the tail of the previous stage)
                L.append(dict(no=no, kind="conv", k=1, cin=cin, cout=out, size=in_size, ds=ds, in_pad=0, out_pad=0,
                              act=ACT_RELU, post_bn=False, residual=None, bn_on_res=False, input=prev))
                sc_no = no
                no += 1
                L.append(dict(no=no, kind="conv", k=3, cin=mid, cout=mid, size=in_size, ds=ds, in_pad=1, out_pad=0,
                              act=ACT_RELU, post_bn=True, residual=None, bn_on_res=False, input=a_no))
                b_no = no
                no += 1
                L.append(dict(no=no, kind="conv", k=1, cin=mid, cout=out, size=size, ds=1, in_pad=0, out_pad=0,
                              act=ACT_RELU, post_bn=False, residual=sc_no, bn_on_res=False, input=b_no))
                shortcut = no
                no += 1
            else:
                L.append(dict(no=no, kind="bn", c=out, size=size, deep=True, out_pad=0, input=shortcut))
                bn_no = no
                no += 1
                L.append(dict(no=no, kind="conv", k=1, cin=out, cout=mid, size=size, ds=1, in_pad=0, out_pad=1,
                              act=ACT_RELU, post_bn=True, residual=None, bn_on_res=False, input=bn_no))
                a_no = no
                no += 1
                L.append(dict(no=no, kind="conv", k=3, cin=mid, cout=mid, size=size, ds=1, in_pad=1, out_pad=0,
                              act=ACT_RELU, post_bn=True, residual=None, bn_on_res=False, input=a_no))
                b_no = no
                no += 1
                L.append(dict(no=no, kind="conv", k=1, cin=mid, cout=out, size=size, ds=1, in_pad=0, out_pad=0,
                              act=ACT_RELU, post_bn=last, residual=shortcut, bn_on_res=last, input=b_no))
                shortcut = no
                no += 1
        prev = shortcut
        if si < 3:
            # first 1x1 of the NEXT stage runs at this stage's resolution (Conv18/34/58)
            nmid = _STAGES[si + 1][0]
            L.append(dict(no=no, kind="conv", k=1, cin=out, cout=nmid, size=size, ds=1, in_pad=0, out_pad=1,
                          act=ACT_RELU, post_bn=True, residual=None, bn_on_res=False, input=prev))
            pending_a = no
            no += 1
    L.append(dict(no=70, kind="gap", c=2048, size=7, input=69))
    L.append(dict(no=72, kind="gemm", cin=2048, cout=1000, input=70))
    return L


def _resnet_blob_size(l):
    if l["kind"] == "bn":
        return 2 * l["c"]
    if l["kind"] == "conv":
        return l["cout"] + l["k"] ** 2 * l["cin"] * l["cout"] + (2 * l["cout"] if l["post_bn"] else 0)
    if l["kind"] == "gemm":
        return l["cout"] + l["cin"] * l["cout"]
    return 0


def resnet50_offsets():
    """Float offsets per layer number.  File order (resnet50.cpp:539-677): layer-number order except
    that in the first block of each stage the projection shortcut is stored AFTER the block's last
    1x1 (e.g. 6, 8, 9, 7 and 18, 20, 21, 19)."""
    layers = {l["no"]: l for l in resnet50_layers()}
    order = []
    nos = sorted(layers)
    first_shortcuts = {7: 9, 19: 21, 35: 37, 59: 61}  # shortcut -> stored after this layer
    for n in nos:
        if n in first_shortcuts:
            continue
        order.append(n)
        for sc, after in first_shortcuts.items():
            if after == n:
                order.append(sc)
    offs, o = {}, 0
    for n in order:
        sz = _resnet_blob_size(layers[n])
        if sz:
            offs[n] = o
            o += sz
    offs["_total"] = o
    return offs


def resnet50_synthetic_weights(seed: int = 50) -> np.ndarray:
    """He-normal convs, BN scale U(0.8,1.2), BN bias U(-0.05,0.05); BN2 = ImageNet
    normalisation; the last 1x1 of every bottleneck is scaled by 0.5 to keep 16 blocks in fp16 range."""
    rng = np.random.default_rng(seed)
    offs = resnet50_offsets()
    blob = np.zeros(offs["_total"], np.float32)
    for l in resnet50_layers():
        n = l["no"]
        if n not in offs:
            continue
        o = offs[n]
        if l["kind"] == "bn":
            c = l["c"]
            if n == 2:
                mean = np.array([0.485, 0.456, 0.406], np.float32)
                std = np.array([0.229, 0.224, 0.225], np.float32)
                blob[o:o + 3] = 1.0 / std
                blob[o + 3:o + 6] = -mean / std
            else:
                blob[o:o + c] = rng.uniform(0.8, 1.2, c)
                blob[o + c:o + 2 * c] = rng.uniform(-0.05, 0.05, c)
        else:
            k = l.get("k", 1)
            ci, co = l["cin"], l["cout"]
            blob[o:o + co] = rng.uniform(-0.05, 0.05, co)
            w = rng.normal(0.0, np.sqrt(2.0 / (k * k * ci)), (co, k, k, ci)).astype(np.float32)
            if l["kind"] == "conv" and l.get("residual") is not None:
                w *= 0.5
            blob[o + co:o + co + w.size] = w.reshape(-1)
            if l["kind"] == "conv" and l["post_bn"]:
                b = o + co + w.size
                blob[b:b + co] = rng.uniform(0.8, 1.2, co)
                blob[b + co:b + 2 * co] = rng.uniform(-0.05, 0.05, co)
    return blob


def resnet50_forward(weights, img_hwc, *, prec=FP32, quirks=QUIRKS_REFERENCE, dump=None):
    """Returns logits[1000] (no softmax: samples/desktop/resnet.cpp:158-172 takes argmax of raw logits)."""
    wts = _f32(weights)
    offs = resnet50_offsets()
    outs = {0: upload_hwc(img_hwc)}
    for l in resnet50_layers():
        n = l["no"]
        x = outs[l["input"]]
        if l["kind"] == "bn":
            y = batchnorm(x, wts[offs[n]:offs[n] + 2 * l["c"]], deep=l["deep"], prec=prec)
        elif l["kind"] == "conv":
            flags = (POST_BATCHNORM if l["post_bn"] else 0) | (BATCHNORM_ON_RESIDUAL if l["bn_on_res"] else 0)
            res = outs[l["residual"]] if l["residual"] is not None else None
            y = conv2d(x, wts[offs[n]:], l["cout"], l["k"], downsample=l["ds"], in_pad=l["in_pad"],
                       out_pad=l["out_pad"], flags=flags, act=l["act"], deep=True, residual=res,
                       quirks=quirks, prec=prec)
        elif l["kind"] == "maxpool":
            y = pool2d(x, pool=3, downsample=2, in_pad=1, is_max=True, act=ACT_RELU, quirks=quirks, prec=prec)
        elif l["kind"] == "gap":
            y = pool2d(x, is_max=False, global_=True, act=ACT_RELU, prec=prec)
        elif l["kind"] == "gemm":
            y = conv2d(x, wts[offs[n]:], l["cout"], 1, deep=True, prec=prec)
        else:
            raise AssertionError(l["kind"])
        outs[n] = y
        if dump is not None:
            dump[n] = y
    return outs[72].reshape(-1)


def num_threads() -> int:
    """OpenMP threads the oracle library currently uses."""
    return int(lib().fyo_set_threads(0))


def set_num_threads(n: int) -> int:
    """Overrides OMP_NUM_THREADS (torchrun exports 1); returns the thread count in effect."""
    return int(lib().fyo_set_threads(int(n)))
