"""Helpers shared by the -m gpu parity tests: run one layer through the C ABI and return CHW output."""
import numpy as np

import fyn_oracle as fo
from fyusenet_b200 import capi

_ctx = None
LAST_KERNEL = 0   # fyn_conv2d_last_kernel of the most recent conv_gpu call


def ctx():
    global _ctx
    if _ctx is None:
        _ctx = capi.Context(0)
    return _ctx


def half(a):
    return np.asarray(a, np.float32).astype(np.float16).astype(np.float32)


def ulp16(ref):
    """One fp16 unit in the last place at |ref| (2^-10 relative for normals, 2^-24 absolute floor)."""
    a = np.maximum(np.abs(ref).astype(np.float64), 2.0 ** -14)
    return (2.0 ** (np.floor(np.log2(a)) - 10)).astype(np.float64)


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def assert_close_f16(got, ref_store, ref_fp32=None, ulps=1.01, rl2=2e-3, extra_abs=2e-5):
    """fp16-storage parity: within `ulps` fp16 ulp of the fp16-store oracle element-wise, plus `extra_abs`
    (fp32 summation-order noise, which exceeds one fp16 ulp for results that cancel to ~0; larger for kernels
    whose operands are fp16-rounded, e.g. tensor-core weights), and rel-L2 vs the fp32 oracle."""
    got = np.asarray(got, np.float64)
    tol = ulps * ulp16(ref_store) + extra_abs
    err = np.abs(got - ref_store)
    worst = float((err / tol).max())
    assert worst <= 1.0, f"max err {err.max():.3e} = {worst:.2f} x tolerance (at {np.unravel_index((err / tol).argmax(), err.shape)})"
    if ref_fp32 is not None:
        assert rel_l2(got, ref_fp32) <= rl2, f"rel-L2 vs fp32 oracle {rel_l2(got, ref_fp32):.3e} > {rl2}"


def conv_gpu(x, wb, *, out_channels, kernel, dtype=capi.F16, downsample=1, dilation=1, in_pad=0, out_pad=0,
             res_pad=0, flags=0, leaky=0.0, source_step=1.0, fractional=False, residual=None, deep=False,
             quirks=capi.QUIRKS_REFERENCE, backend=capi.BACKEND_AUTO, in_dtype=None, in_packing=0, want_op=False):
    """x: [C][H][W] or [N][C][H][W].  Returns CHW output (and the op backend if want_op)."""
    c = ctx()
    x = np.asarray(x, np.float32)
    batch = x.shape[0] if x.ndim == 4 else 1
    ci, h, w = x.shape[-3:]
    order = capi.ORDER_DEEP if deep else capi.ORDER_SHALLOW
    fl = flags | (capi.FLAG_DEEP if deep else 0) | (capi.FLAG_RESIDUAL_INPUT if residual is not None else 0)
    op = capi.Conv2d(c, wb, width=w, height=h, in_channels=ci, out_channels=out_channels, kernel=kernel,
                     downsample=downsample, dilation=dilation, in_padding=in_pad, out_padding=out_pad,
                     res_padding=res_pad, flags=fl, leaky=leaky, source_step=source_step, fractional=fractional,
                     quirks=quirks, backend=backend)
    tin = c.tensor(w, h, ci, in_pad, order, dtype if in_dtype is None else in_dtype, batch, in_packing)
    tout = c.tensor(op.out_width, op.out_height, out_channels, out_pad, order, dtype, batch)
    tres = None
    tin.write_chw(x)
    if residual is not None:
        tres = c.tensor(op.out_width, op.out_height, out_channels, res_pad, order, dtype, batch)
        tres.write_chw(residual)
    op.run(tin, tout, tres)
    y = tout.read_chw()
    raw = tout.download()
    be = op.backend
    global LAST_KERNEL
    LAST_KERNEL = op.last_kernel
    for t in (tin, tout, tres):
        if t is not None:
            t.destroy()
    op.destroy()
    if want_op:
        return y, be, raw
    return y


def conv_oracle(x, wb, *, out_channels, kernel, prec, flags=0, act=fo.ACT_NONE, leaky=0.0, residual=None, **kw):
    x = np.asarray(x, np.float32)
    if x.ndim == 4:
        res = residual if residual is not None else [None] * x.shape[0]
        return np.stack([conv_oracle(x[i], wb, out_channels=out_channels, kernel=kernel, prec=prec, flags=flags,
                                     act=act, leaky=leaky, residual=res[i], **kw) for i in range(x.shape[0])])
    return fo.conv2d(x, wb, out_channels, kernel, flags=flags, act=act, leak=leaky, residual=residual, prec=prec, **kw)


def random_wb(rng, ci, co, k, post_bn=False, scale=None):
    std = np.sqrt(2.0 / (k * k * ci)) if scale is None else scale
    parts = [rng.uniform(-0.5, 0.5, co), rng.normal(0, std, co * k * k * ci)]
    if post_bn:
        parts += [rng.uniform(0.5, 1.5, co), rng.uniform(-0.5, 0.5, co)]
    return np.concatenate(parts).astype(np.float32)
