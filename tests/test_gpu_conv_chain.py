"""GPU tests of the persistent convolution chain (fyn_conv_chain.cu) through the C ABI.

The chain runs the SAME arithmetic as the single-layer tcgen05 kernels (same step tables, weight images and epilogue), so the
bar is bit-identity with the ops run one by one -- whose parity with the oracle is the subject of test_gpu_conv_tc.py -- plus
one direct comparison with the oracle on a StyleNet-like trunk.
"""
import os

import numpy as np
import pytest

import fyn_oracle as fo
from fyusenet_b200 import capi
from gpu_util import ctx, half, random_wb, rel_l2

pytestmark = pytest.mark.gpu


def _trunk_flags(nlayers, style="stylenet"):
    """(flags, residual_from) per layer.  stylenet: blocks of two layers, the second adds the block input; res1_2 ReLUs its
    residual and res2_1 has no prefix activation (stylenet9x9.cpp:145-186)."""
    flags, rfrom = [], []
    for i in range(nlayers):
        f = capi.FLAG_PRE_RELU
        r = -2
        if style == "stylenet":
            if i == 2:
                f = 0
            if i % 2 == 1:
                f |= capi.FLAG_RESIDUAL_INPUT
                r = i - 2
                if i == 1:
                    f |= capi.FLAG_RELU_ON_RESIDUAL
        elif style == "plain":
            pass
        elif style == "leaky":
            if i % 2 == 1:
                f |= capi.FLAG_RESIDUAL_INPUT
                r = i - 2
        flags.append(f)
        rfrom.append(r)
    return flags, rfrom


def _build(c, w, h, ch, k, nlayers, style, seed, post_bn=False, leaky=0.0, pad=None):
    pad = k // 2 if pad is None else pad
    rng = np.random.default_rng(seed)
    flags, rfrom = _trunk_flags(nlayers, style)
    ops, wbs = [], []
    for i in range(nlayers):
        wb = random_wb(rng, ch, ch, k, post_bn=post_bn)
        fl = flags[i] | (capi.FLAG_POST_BATCHNORM if post_bn else 0)
        ops.append(capi.Conv2d(c, wb, width=w, height=h, in_channels=ch, out_channels=ch, kernel=k, in_padding=pad,
                               out_padding=pad, res_padding=pad, flags=fl, leaky=leaky, backend=capi.BACKEND_TC))
        wbs.append(wb)
    return ops, wbs, flags, rfrom


def _run_one_by_one(c, ops, rfrom, x, batch, w, h, ch, pad):
    """Reference execution: every layer through fyn_conv2d_run on its own tensor."""
    tens = [c.tensor(w, h, ch, pad, capi.ORDER_SHALLOW, capi.F16, batch) for _ in range(len(ops) + 1)]
    tens[0].write_chw(x)
    for i, op in enumerate(ops):
        res = tens[rfrom[i] + 1] if rfrom[i] >= -1 else None
        op.run(tens[i], tens[i + 1], res)
    y = tens[-1].read_chw()
    for t in tens:
        t.destroy()
    return y


def _run_chain(c, ops, rfrom, x, batch, w, h, ch, pad, repeats=1):
    chain = capi.ConvChain(c, ops, rfrom)
    tin = c.tensor(w, h, ch, pad, capi.ORDER_SHALLOW, capi.F16, batch)
    tout = c.tensor(w, h, ch, pad, capi.ORDER_SHALLOW, capi.F16, batch)
    tin.write_chw(x)
    ys = []
    for _ in range(repeats):
        assert chain.run(tin, tout), "chain declined tensors it must cover"
        ys.append(tout.read_chw())
    chain.destroy()
    tin.destroy()
    tout.destroy()
    return ys


@pytest.mark.parametrize("w,h,ch,nlayers,style,pad", [
    (40, 24, 40, 4, "stylenet", 1),       # one column block, two short strips
    (381, 116, 40, 10, "stylenet", 0),    # StyleNet trunk: clamp-to-edge tensors without padding, three column blocks (last one ragged)
    (381, 116, 40, 10, "stylenet", 1),    # the same on zero-padded tensors
    (130, 37, 40, 5, "stylenet", 0),      # odd layer count, a 2-pixel column block
    (257, 9, 24, 3, "plain", 1),          # channels that do not fill the last chunk pair; fewer rows than strips
    (96, 64, 16, 6, "leaky", 0),          # layers that run row-stacked on their own: the chain uses their single-row plans
    (128, 30, 16, 4, "leaky", 1),
])
def test_chain_is_bit_identical_to_single_layers(w, h, ch, nlayers, style, pad):
    c = ctx()
    ops, wbs, flags, rfrom = _build(c, w, h, ch, 3, nlayers, style, seed=w + h + ch, leaky=0.1 if style == "leaky" else 0.0, pad=pad)
    rng = np.random.default_rng(5)
    x = half(rng.normal(size=(ch, h, w)))
    want = _run_one_by_one(c, ops, rfrom, x, 1, w, h, ch, pad)
    got = _run_chain(c, ops, rfrom, x, 1, w, h, ch, pad, repeats=3)
    for y in got:                                    # repeated launches: the epoch-tagged progress counters need no reset
        np.testing.assert_array_equal(y, want)
    for op in ops:
        op.destroy()


def test_chain_batch_and_strip_variants(monkeypatch):
    """Batch 2 (strips of both images share the grid) and forced strip geometries: tall strips, a single strip per CTA."""
    c = ctx()
    w, h, ch, n = 200, 50, 40, 6
    ops, wbs, flags, rfrom = _build(c, w, h, ch, 3, n, "stylenet", seed=77, pad=0)
    rng = np.random.default_rng(6)
    x = half(rng.normal(size=(2, ch, h, w)))
    want = _run_one_by_one(c, ops, rfrom, x, 2, w, h, ch, 0)
    np.testing.assert_array_equal(_run_chain(c, ops, rfrom, x, 2, w, h, ch, 0)[0], want)
    for env in ({"FYN_CHAIN_SH": "13"}, {"FYN_CHAIN_NSUB": "1"}, {"FYN_CHAIN_SH": "1"}, {"FYN_CHAIN_SLOTS": "6"}, {"FYN_CHAIN_SLOTS": "8", "FYN_CHAIN_SH": "2"},
                {"FYN_CHAIN_EPI": "8"}, {"FYN_CHAIN_TMA": "1"}, {"FYN_CHAIN_FENCE": "3"}, {"FYN_CHAIN_NSUB": "3", "FYN_CHAIN_SH": "2"}, {"FYN_CHAIN_NSUB": "4", "FYN_CHAIN_SH": "1"}, {"FYN_CHAIN_NSUB": "4", "FYN_CHAIN_SH": "3"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        np.testing.assert_array_equal(_run_chain(c, ops, rfrom, x, 2, w, h, ch, 0)[0], want, err_msg=str(env))
        for k in env:
            monkeypatch.delenv(k)
    for op in ops:
        op.destroy()


def test_chain_matches_oracle():
    """A four-layer StyleNet-like trunk against the CPU oracle (fp16 storage, half-rounded weights = the exact model of the
    tensor-core kernels): same bound as the single layers, accumulated over four layers."""
    c = ctx()
    w, h, ch, n = 150, 40, 40, 4
    ops, wbs, flags, rfrom = _build(c, w, h, ch, 3, n, "stylenet", seed=3, pad=0)
    rng = np.random.default_rng(8)
    x = half(rng.normal(size=(ch, h, w)))
    got = _run_chain(c, ops, rfrom, x, 1, w, h, ch, 0)[0]
    acts = [x]
    for i in range(n):
        wb = np.array(wbs[i], np.float32, copy=True)
        wb[ch:] = half(wb[ch:])
        res = acts[rfrom[i] + 1] if rfrom[i] >= -1 else None
        ofl = fo.RELU_ON_RESIDUAL if flags[i] & capi.FLAG_RELU_ON_RESIDUAL else 0
        acts.append(fo.conv2d(acts[-1], wb, ch, 3, act=fo.ACT_RELU if flags[i] & capi.FLAG_PRE_RELU else fo.ACT_NONE, flags=ofl,
                              in_pad=0, out_pad=0, residual=res, prec=fo.FP16_STORE))
    assert rel_l2(got, acts[-1]) <= 2e-3
    assert np.abs(got - acts[-1]).max() <= 2e-2 * np.abs(acts[-1]).max()
    for op in ops:
        op.destroy()


def test_chain_rejects_what_it_cannot_run():
    c = ctx()
    rng = np.random.default_rng(1)
    mk = lambda **kw: capi.Conv2d(c, random_wb(rng, kw.get("ci", 40), kw.get("co", 40), 3), width=64, height=32, in_channels=kw.get("ci", 40),
                                  out_channels=kw.get("co", 40), kernel=3, in_padding=1, out_padding=1, res_padding=1, flags=kw.get("flags", 0),
                                  downsample=kw.get("ds", 1), backend=capi.BACKEND_TC)
    a, b = mk(), mk()
    other = mk(ci=40, co=24)
    res0 = mk(flags=capi.FLAG_RESIDUAL_INPUT)
    with pytest.raises(capi.FynError):
        capi.ConvChain(c, [a, other], None)                  # different geometry
    with pytest.raises(capi.FynError):
        capi.ConvChain(c, [a, res0], [-2, 0])                # residual that is not the previous layer's input
    with pytest.raises(capi.FynError):
        capi.ConvChain(c, [a], None)                         # a chain has at least two layers
    chain = capi.ConvChain(c, [a, b], None)
    # tensors in a format the chain does not cover: declined, nothing enqueued
    t32a = c.tensor(64, 32, 40, 1, capi.ORDER_SHALLOW, capi.F32, 1)
    t32b = c.tensor(64, 32, 40, 1, capi.ORDER_SHALLOW, capi.F32, 1)
    assert chain.run(t32a, t32b) is False
    chain.destroy()
    for t in (t32a, t32b):
        t.destroy()
    for op in (a, b, other, res0):
        op.destroy()
