"""Golden vectors (tests/golden/layers_v1.npz, written by tests/golden/make_golden.py from the CPU oracle; the GL reference
cannot run here).  The CPU test freezes the oracle against them bit for bit; the GPU tests compare the CUDA
layers with the stored outputs directly -- no oracle call on that path.

GPU tolerances (fp16 storage): tensor-core convolutions use fp16-rounded weights, so <= 2e-3 rel-L2 and 2e-2 max-abs; the
bandwidth-bound layers <= 1 fp16 ulp (+2e-5); the copies exact; the StyleNet frame <= 6e-3 (1.5 / 255)."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
import make_golden  # noqa: E402

GOLDEN = np.load(Path(__file__).resolve().parent / "golden" / "layers_v1.npz")


def _g(name, key):
    return GOLDEN[f"{name}/{key}"]


def test_oracle_reproduces_golden_vectors():
    cases = make_golden.cases()
    assert sorted({k.split("/")[0] for k in GOLDEN.files}) == sorted(cases)
    for name, (fn, stored) in cases.items():
        for k, v in stored.items():
            np.testing.assert_array_equal(v, _g(name, k), err_msg=f"{name}/{k}: generator input changed")
        np.testing.assert_array_equal(fn(), _g(name, "out"), err_msg=f"{name}: the oracle moved")


@pytest.mark.gpu
def test_cuda_layers_match_golden_vectors():
    from fyusenet_b200 import capi, hostapi
    from gpu_util import assert_close_f16, conv_gpu, ctx, rel_l2
    c = ctx()

    def conv_close(y, ref):
        assert rel_l2(y, ref) <= 2e-3 and float(np.abs(y - ref).max()) <= 2e-2 * max(1.0, float(np.abs(ref).max()))

    n = "conv3x3_shallow_relu"
    conv_close(conv_gpu(_g(n, "x"), _g(n, "wb"), out_channels=8, kernel=3, in_pad=1, flags=capi.FLAG_PRE_RELU), _g(n, "out"))
    n = "conv9x9_shallow"
    conv_close(conv_gpu(_g(n, "x"), _g(n, "wb"), out_channels=12, kernel=9, flags=capi.FLAG_PRE_RELU), _g(n, "out"))
    n = "fraconv3x3_step025_ds2"
    conv_close(conv_gpu(_g(n, "x"), _g(n, "wb"), out_channels=12, kernel=3, downsample=2, source_step=0.25, fractional=True,
                        flags=capi.FLAG_PRE_RELU), _g(n, "out"))
    n = "conv3x3_deep_bn_residual"
    conv_close(conv_gpu(_g(n, "x"), _g(n, "wb"), out_channels=72, kernel=3, in_pad=1, deep=True, residual=_g(n, "res"),
                        flags=capi.FLAG_PRE_RELU | capi.FLAG_POST_BATCHNORM | capi.FLAG_RELU_ON_RESIDUAL), _g(n, "out"))

    def run1(op, x, out_shape, in_pad=0, out_pad=0, deep=False, extra=()):
        order = capi.ORDER_DEEP if deep else capi.ORDER_SHALLOW
        ch, h, w = x.shape
        tin = c.tensor(w, h, ch, in_pad, order, capi.F16)
        tout = c.tensor(out_shape[2], out_shape[1], out_shape[0], out_pad, order, capi.F16)
        tin.write_chw(x)
        op.run(tin, *extra, tout)
        y = tout.read_chw()
        for o in (tin, tout, op):
            o.destroy()
        return y

    n = "maxpool3x3_s2_deep"
    x = _g(n, "x")
    y = run1(capi.Pool2d(c, width=12, height=12, channels=64, pool=3, downsample=2, in_padding=1, is_max=True,
                         flags=capi.FLAG_DEEP | capi.FLAG_PRE_RELU), x, _g(n, "out").shape, in_pad=1, deep=True)
    np.testing.assert_array_equal(y, _g(n, "out"))
    n = "globavg7x7_deep"
    y = run1(capi.Pool2d(c, width=7, height=7, channels=36, pool=7, downsample=7, is_max=False, global_=True,
                         flags=capi.FLAG_DEEP | capi.FLAG_PRE_RELU), _g(n, "x"), _g(n, "out").shape, deep=True)
    assert_close_f16(y, _g(n, "out"))
    n = "batchnorm_deep"
    y = run1(capi.BatchNorm(c, _g(n, "sb"), width=9, height=6, channels=31, flags=capi.FLAG_DEEP), _g(n, "x"), _g(n, "out").shape, deep=True)
    assert_close_f16(y, _g(n, "out"))
    n = "sigmoid"
    y = run1(capi.Sigmoid(c, width=8, height=6, channels=3), _g(n, "x"), _g(n, "out").shape)
    assert_close_f16(y, _g(n, "out"), ulps=1.5)
    n = "scale_linear_x2_pad1"
    y = run1(capi.Scale(c, width=8, height=6, channels=9, in_padding=1, up=(2, 2), linear=True), _g(n, "x"), _g(n, "out").shape, in_pad=1)
    assert_close_f16(y, _g(n, "out"))
    n = "scale_nearest_div2_deep"
    y = run1(capi.Scale(c, width=8, height=6, channels=9, in_padding=1, down=(2, 2), flags=capi.FLAG_DEEP), _g(n, "x"), _g(n, "out").shape,
             in_pad=1, deep=True)
    np.testing.assert_array_equal(y, _g(n, "out"))
    n = "sub_relu"
    t2 = c.tensor(8, 6, 9, 0, capi.ORDER_SHALLOW, capi.F16)
    t2.write_chw(_g(n, "y"))
    y = run1(capi.Arith(c, width=8, height=6, channels=9, op=capi.ARITH_SUB, flags=capi.FLAG_PRE_RELU), _g(n, "x"), _g(n, "out").shape, extra=(t2,))
    t2.destroy()
    assert_close_f16(y, _g(n, "out"))
    n = "dwconv3x3_shallow_bn_refquirk"
    y = run1(capi.DwConv3x3(c, _g(n, "wb"), width=12, height=9, channels=10, in_padding=1, flags=capi.FLAG_POST_BATCHNORM), _g(n, "x"),
             _g(n, "out").shape, in_pad=1)
    assert_close_f16(y, _g(n, "out"))
    n = "stylenet3x3_24x16"
    from fyusenet_b200 import synthetic
    net = hostapi.StyleNet(3, 24, 16)
    net.load_weights(synthetic.stylenet_weights(3))
    net.setup()
    net.set_input(_g(n, "img"))
    net.forward()
    got = net.output_rgba()[0].copy()
    net.destroy()
    assert got.shape == _g(n, "out").shape
    assert float(np.abs(got[..., :3] - _g(n, "out")[..., :3]).max()) <= 6e-3
