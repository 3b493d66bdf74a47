#!/usr/bin/env python
"""Randomised cross-check of the tcgen05 convolution family against the direct (exact-fp32) kernel through the C ABI:
random kernel sizes, channel counts, strides, fractional steps, paddings, residual / activation flags, image sizes and
batch.  python tools/fuzz_conv.py [cases] [seed] [deep]     (a stuck case reports itself after 60 s instead of hanging;
"deep" fuzzes the deep-tiled family of fyn_conv_deep_tc.cu: 1x1 / 3x3 on multiples of 64 channels, tap-packed <= 4 channels)"""
import faulthandler
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fyusenet_b200 import capi  # noqa: E402


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b.astype(np.float64)), 1e-30))


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    ctx = capi.Context(0)
    ran = skipped = 0
    worst = 0.0
    for case in range(cases):
        kw = {}
        kind = rng.choice(["regular", "stride2", "frac1", "frac2", "rgb", "small"])
        k = int(rng.choice([3, 3, 3, 5, 7, 9]))
        ci = int(rng.choice([8, 12, 16, 20, 24, 32, 40, 48]))
        if kind == "small":                  # 1x1 kernels and thin inputs on the plane-chunk path
            k = int(rng.choice([1, 1, 1, 3, 5]))
            ci = int(rng.choice([1, 2, 3, 5, 6, 7, 8, 16, 40])) if k == 1 else int(rng.choice([5, 6, 7]))
            if rng.random() < 0.4:
                kw["downsample"] = 2
        co = int(rng.integers(1, 41))
        if kind == "stride2":
            kw["downsample"] = 2
        elif kind == "frac1":
            kw.update(downsample=2, source_step=0.5, fractional=True)
        elif kind == "frac2":
            kw.update(source_step=0.5, fractional=True)
            if rng.random() < 0.5:
                kw.update(downsample=2, source_step=0.25)
        elif kind == "rgb":
            ci = int(rng.choice([1, 3, 4]))
        w = int(rng.integers(9, 700))
        h = int(rng.integers(9, 200))
        batch = int(rng.choice([1, 1, 1, 2]))
        relu = bool(rng.random() < 0.7)
        in_pad = int(rng.choice([0, 0, 1])) if kind in ("regular", "stride2", "small") else 0
        out_pad = int(rng.choice([0, 0, 1]))
        res = bool(rng.random() < 0.3) and kind != "rgb"
        flags = (capi.FLAG_PRE_RELU if relu else 0) | (capi.FLAG_RESIDUAL_INPUT if res else 0)
        if res and rng.random() < 0.5:
            flags |= capi.FLAG_RELU_ON_RESIDUAL
        post_bn = bool(rng.random() < 0.2)
        if post_bn:
            flags |= capi.FLAG_POST_BATCHNORM
        wb = [rng.uniform(-0.5, 0.5, co), rng.normal(0, np.sqrt(2.0 / (k * k * ci)), co * k * k * ci)]
        if post_bn:
            wb += [rng.uniform(0.5, 1.5, co), rng.uniform(-0.5, 0.5, co)]
        wb = np.concatenate(wb).astype(np.float32)
        desc = dict(width=w, height=h, in_channels=ci, out_channels=co, kernel=k, in_padding=in_pad, out_padding=out_pad, flags=flags, **kw)
        faulthandler.dump_traceback_later(60, exit=True)
        outs = {}
        try:
            for backend in (capi.BACKEND_TC, capi.BACKEND_DIRECT):
                try:
                    op = capi.Conv2d(ctx, wb, backend=backend, **desc)
                except capi.FynError:
                    outs = None              # shape outside the tcgen05 family (or illegal): nothing to compare
                    break
                if op.out_width <= 0 or op.out_height <= 0:
                    outs = None
                    op.destroy()
                    break
                tin = ctx.tensor(w, h, ci, in_pad, capi.ORDER_SHALLOW, capi.F16, batch)
                tout = ctx.tensor(op.out_width, op.out_height, co, out_pad, capi.ORDER_SHALLOW, capi.F16, batch)
                tres = ctx.tensor(op.out_width, op.out_height, co, 0, capi.ORDER_SHALLOW, capi.F16, batch) if res else None
                r2 = np.random.default_rng(case)
                tin.write_chw(r2.normal(size=(batch, ci, h, w)).astype(np.float32))
                if res:
                    tres.write_chw(r2.normal(size=(batch, co, op.out_height, op.out_width)).astype(np.float32))
                for _ in range(3):
                    op.run(tin, tout, tres)
                ctx.stream_sync()
                outs[backend] = (tout.read_chw(), op.backend)
                for o in (tin, tout, tres, op):
                    if o is not None:
                        o.destroy()
        finally:
            faulthandler.cancel_dump_traceback_later()
        if outs is None:
            skipped += 1
            continue
        (y, be), (yd, _) = outs[capi.BACKEND_TC], outs[capi.BACKEND_DIRECT]
        err = rel_l2(y, yd)
        worst = max(worst, err)
        ran += 1
        status = "ok" if err <= 2.5e-3 else "MISMATCH"
        if status != "ok" or case % 25 == 0:
            print(f"case {case}: {kind} k{k} {ci}->{co} {w}x{h} b{batch} relu={relu} res={res} bn={post_bn} pads={in_pad}/{out_pad} {kw} backend={be} rel-L2 {err:.2e} {status}", flush=True)
        if status != "ok":
            raise SystemExit(1)
    print(f"fuzz: {ran} cases compared, {skipped} outside the family, worst rel-L2 {worst:.2e}")


def main_deep():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    ctx = capi.Context(0)
    ran = skipped = 0
    worst = 0.0
    for case in range(cases):
        kind = rng.choice(["1x1", "3x3", "3x3", "stem"])
        if kind == "stem":
            k, ci, in_pad = int(rng.choice([3, 5, 7])), int(rng.choice([1, 3, 4])), int(rng.choice([0, 1, 3]))
        else:
            k = 1 if kind == "1x1" else 3
            ci = 64 * int(rng.choice([1, 1, 2, 3, 4, 8]))
            in_pad = 1 if k == 3 else int(rng.choice([0, 1]))
        co = int(rng.choice([int(rng.integers(1, 300)), 64, 128, 256]))
        ds = int(rng.choice([1, 1, 2]))
        w, h = int(rng.integers(1, 60)), int(rng.integers(1, 60))
        if w // ds < 1 or h // ds < 1:
            ds = 1
        batch = int(rng.choice([1, 1, 2, 5]))
        relu = bool(rng.random() < 0.6)
        out_pad = int(rng.choice([0, 1]))
        res = bool(rng.random() < 0.4)
        post_bn = bool(rng.random() < 0.5)
        flags = capi.FLAG_DEEP | (capi.FLAG_PRE_RELU if relu else 0) | (capi.FLAG_RESIDUAL_INPUT if res else 0) | (capi.FLAG_POST_BATCHNORM if post_bn else 0)
        if res and rng.random() < 0.5:
            flags |= capi.FLAG_RELU_ON_RESIDUAL
        if res and post_bn and rng.random() < 0.5:
            flags |= capi.FLAG_BATCHNORM_ON_RESIDUAL
        res_pad = int(rng.choice([0, 1])) if res else 0
        wb = [rng.uniform(-0.5, 0.5, co), rng.normal(0, np.sqrt(2.0 / (k * k * ci)), co * k * k * ci)]
        if post_bn:
            wb += [rng.uniform(0.5, 1.5, co), rng.uniform(-0.5, 0.5, co)]
        wb = np.concatenate(wb).astype(np.float32)
        desc = dict(width=w, height=h, in_channels=ci, out_channels=co, kernel=k, downsample=ds, in_padding=in_pad, out_padding=out_pad,
                    res_padding=res_pad, flags=flags)
        faulthandler.dump_traceback_later(60, exit=True)
        outs = {}
        try:
            for backend in (capi.BACKEND_TC, capi.BACKEND_DIRECT):
                try:
                    op = capi.Conv2d(ctx, wb, backend=backend, **desc)
                except capi.FynError:
                    outs = None
                    break
                tin = ctx.tensor(w, h, ci, in_pad, capi.ORDER_DEEP, capi.F16, batch)
                tout = ctx.tensor(op.out_width, op.out_height, co, out_pad, capi.ORDER_DEEP, capi.F16, batch)
                tres = ctx.tensor(op.out_width, op.out_height, co, res_pad, capi.ORDER_DEEP, capi.F16, batch) if res else None
                r2 = np.random.default_rng(case)
                tin.write_chw(r2.normal(size=(batch, ci, h, w)).astype(np.float32))
                if res:
                    tres.write_chw(r2.normal(size=(batch, co, op.out_height, op.out_width)).astype(np.float32))
                for _ in range(2):
                    op.run(tin, tout, tres)
                ctx.stream_sync()
                outs[backend] = (tout.read_chw(), op.backend)
                for o in (tin, tout, tres, op):
                    if o is not None:
                        o.destroy()
        finally:
            faulthandler.cancel_dump_traceback_later()
        if outs is None:
            skipped += 1
            continue
        (y, be), (yd, _) = outs[capi.BACKEND_TC], outs[capi.BACKEND_DIRECT]
        err = rel_l2(y, yd)
        worst = max(worst, err)
        ran += 1
        status = "ok" if err <= 1.5e-3 and be == 2 else "MISMATCH"
        if status != "ok" or case % 25 == 0:
            print(f"case {case}: deep {kind} k{k} {ci}->{co} {w}x{h} ds{ds} b{batch} relu={relu} res={res} bn={post_bn} flags={flags} pads={in_pad}/{out_pad}/{res_pad} backend={be} rel-L2 {err:.2e} {status}", flush=True)
        if status != "ok":
            raise SystemExit(1)
    print(f"fuzz deep: {ran} cases compared, {skipped} outside the family, worst rel-L2 {worst:.2e}")


if __name__ == "__main__":
    main_deep() if len(sys.argv) > 3 and sys.argv[3] == "deep" else main()
