#!/usr/bin/env python
"""Writes tests/golden/layers_v1.npz: seeded inputs, parameters and ORACLE outputs of one small case per layer family.

The reference itself cannot run in the build container (OpenGL), so these vectors are produced by the CPU
oracle (oracle/fyn_oracle.c, pinned against the reference's own known-answer tests in tests/test_oracle_kat.py and
tests/test_oracle_arith_scale.py).  They freeze the oracle: tests/test_golden.py fails if a later change to the oracle
moves any value, and the GPU suite compares the CUDA layers with the stored outputs without calling the oracle.

    python tests/golden/make_golden.py        (rewrites the file; commit the result together with the oracle change)
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "oracle"))
import fyn_oracle as fo  # noqa: E402


def cases():
    """name -> (callable producing the oracle output, dict of stored arrays)"""
    rng = np.random.default_rng(20261017)
    out = {}

    def wb_conv(co, k, ci, bn=False):
        parts = [rng.uniform(-0.5, 0.5, co), rng.normal(0, np.sqrt(2.0 / (k * k * ci)), co * k * k * ci)]
        if bn:
            parts += [rng.uniform(0.5, 1.5, co), rng.uniform(-0.3, 0.3, co)]
        return np.concatenate(parts).astype(np.float32)

    h16 = fo.half_round
    x = h16(rng.normal(size=(12, 10, 14)).astype(np.float32))
    wb = wb_conv(8, 3, 12)
    out["conv3x3_shallow_relu"] = (lambda x=x, wb=wb: fo.conv2d(x, wb, 8, 3, in_pad=1, act=fo.ACT_RELU, prec=fo.FP16_STORE), dict(x=x, wb=wb))
    x = h16(rng.normal(size=(3, 12, 16)).astype(np.float32))
    wb = wb_conv(12, 9, 3)
    out["conv9x9_shallow"] = (lambda x=x, wb=wb: fo.conv2d(x, wb, 12, 9, act=fo.ACT_RELU, prec=fo.FP16_STORE), dict(x=x, wb=wb))
    x = h16(rng.normal(size=(20, 8, 10)).astype(np.float32))
    wb = wb_conv(12, 3, 20)
    out["fraconv3x3_step025_ds2"] = (lambda x=x, wb=wb: fo.conv2d(x, wb, 12, 3, downsample=2, source_step=0.25, fractional=True, act=fo.ACT_RELU, prec=fo.FP16_STORE),
                                     dict(x=x, wb=wb))
    x = h16(rng.normal(size=(64, 7, 7)).astype(np.float32))
    wb = wb_conv(72, 3, 64, bn=True)
    res = h16(rng.normal(size=(72, 7, 7)).astype(np.float32))
    out["conv3x3_deep_bn_residual"] = (lambda x=x, wb=wb, res=res: fo.conv2d(x, wb, 72, 3, in_pad=1, flags=fo.POST_BATCHNORM | fo.RELU_ON_RESIDUAL, deep=True,
                                                                           residual=res, act=fo.ACT_RELU, prec=fo.FP16_STORE), dict(x=x, wb=wb, res=res))
    x = h16(rng.normal(size=(64, 12, 12)).astype(np.float32))
    out["maxpool3x3_s2_deep"] = (lambda x=x: fo.pool2d(x, pool=3, downsample=2, in_pad=1, is_max=True, act=fo.ACT_RELU, prec=fo.FP16_STORE), dict(x=x))
    x = h16(rng.normal(size=(36, 7, 7)).astype(np.float32))
    out["globavg7x7_deep"] = (lambda x=x: fo.pool2d(x, pool=7, downsample=7, is_max=False, global_=True, act=fo.ACT_RELU, prec=fo.FP16_STORE), dict(x=x))
    x = h16(rng.uniform(-4, 4, size=(31, 6, 9)).astype(np.float32))
    sb = rng.uniform(-2, 2, 62).astype(np.float32)
    out["batchnorm_deep"] = (lambda x=x, sb=sb: fo.batchnorm(x, sb, deep=True, prec=fo.FP16_STORE), dict(x=x, sb=sb))
    x = h16(rng.normal(size=(3, 6, 8)).astype(np.float32) * 3)
    out["sigmoid"] = (lambda x=x: fo.sigmoid(x, prec=fo.FP16_STORE), dict(x=x))
    x = h16(rng.normal(size=(9, 6, 8)).astype(np.float32))
    out["scale_linear_x2_pad1"] = (lambda x=x: fo.scale(x, up=(2, 2), linear=True, in_pad=1, prec=fo.FP16_STORE), dict(x=x))
    out["scale_nearest_div2_deep"] = (lambda x=x: fo.scale(x, down=(2, 2), in_pad=1, deep=True, prec=fo.FP16_STORE), dict(x=x))
    y = h16(rng.normal(size=(9, 6, 8)).astype(np.float32))
    out["sub_relu"] = (lambda x=x, y=y: fo.arith(x, y, fo.ARITH_SUB, act=fo.ACT_RELU, prec=fo.FP16_STORE), dict(x=x, y=y))
    x = h16(rng.normal(size=(10, 9, 12)).astype(np.float32))
    wb = np.concatenate([rng.uniform(0.5, 1.5, 10), rng.normal(size=90) * 0.4, rng.uniform(0.5, 1.5, 10), rng.uniform(-0.2, 0.2, 10)]).astype(np.float32)
    out["dwconv3x3_shallow_bn_refquirk"] = (lambda x=x, wb=wb: fo.dwconv3x3(x, wb, in_pad=1, post_bn=True, quirks=fo.QUIRK_DW_BN_OFFSET, prec=fo.FP16_STORE),
                                            dict(x=x, wb=wb))
    w3 = fo.stylenet_synthetic_weights(3)
    img = fo.synthetic_image(16, 24, 5)
    out["stylenet3x3_24x16"] = (lambda w3=w3, img=img: fo.stylenet_forward(w3, img, 3, prec=fo.FP16_STORE), dict(img=img))
    return out


def main():
    arrays = {}
    for name, (fn, stored) in cases().items():
        for k, v in stored.items():
            arrays[f"{name}/{k}"] = v
        arrays[f"{name}/out"] = fn().astype(np.float32)
    path = Path(__file__).resolve().parent / "layers_v1.npz"
    np.savez_compressed(path, **arrays)
    print(f"{path}: {len(arrays)} arrays, {path.stat().st_size} bytes")


if __name__ == "__main__":
    main()
