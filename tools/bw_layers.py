#!/usr/bin/env python
"""Achieved HBM bandwidth of the bandwidth-bound layers (bytes = input once + output once at the storage type),
timed with CUDA events through the C ABI on ResNet-50 / StyleNet shapes whose working set exceeds the 126 MB L2:

    python tools/bw_layers.py [reps]          -> one JSON line per layer: us, GB/s, fraction of the measured copy bandwidth

Peak = MEASURED_PEAKS.json hbm_gbs (driver-measured copy), else the 6650 GB/s fallback of the profiling guide."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fyusenet_b200 import capi  # noqa: E402


def peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    ctx = capi.Context(0)
    hbm, which = peak()
    e0, e1 = ctx.event_create(), ctx.event_create()
    rng = np.random.default_rng(0)
    D, S = capi.ORDER_DEEP, capi.ORDER_SHALLOW
    B = 128

    def timed(name, fn, nbytes, note=""):
        for _ in range(3):
            fn()
        ctx.stream_sync()
        ctx.event_record(e0)
        for _ in range(reps):
            fn()
        ctx.event_record(e1)
        ctx.event_sync(e1)
        us = ctx.elapsed_ms(e0, e1) / reps * 1e3
        gbs = nbytes / (us * 1e-6) / 1e9
        print(json.dumps({"layer": name, "us": round(us, 1), "MB": round(nbytes / 1e6, 1), "GB/s": round(gbs, 1), "frac_of_hbm": round(gbs / hbm, 3),
                          "peak": hbm, "peak_source": which, "note": note}), flush=True)

    def tbytes(t):
        return t.geom.bytes

    # ---- ResNet-50 shapes, batch 128, fp16 deep tensors
    tin = ctx.tensor(112, 112, 64, 1, D, capi.F16, B)
    tout = ctx.tensor(56, 56, 64, 1, D, capi.F16, B)
    op = capi.Pool2d(ctx, width=112, height=112, channels=64, pool=3, downsample=2, in_padding=1, out_padding=1, is_max=True, flags=capi.FLAG_DEEP | capi.FLAG_PRE_RELU)
    timed("maxpool 3x3 s2 112x112x64 -> 56x56 (ResNet MaxPool4), batch 128", lambda: op.run(tin, tout), tbytes(tin) + tbytes(tout))
    for o in (tin, tout, op):
        o.destroy()
    tin = ctx.tensor(7, 7, 2048, 0, D, capi.F16, B * 4)
    tout = ctx.tensor(1, 1, 2048, 0, D, capi.F16, B * 4)
    op = capi.Pool2d(ctx, width=7, height=7, channels=2048, is_max=False, global_=True, flags=capi.FLAG_DEEP | capi.FLAG_PRE_RELU)
    timed("global avgpool 7x7x2048 (ResNet AvgPool70), batch 512", lambda: op.run(tin, tout), tbytes(tin) + tbytes(tout))
    for o in (tin, tout, op):
        o.destroy()
    tin = ctx.tensor(56, 56, 256, 0, D, capi.F16, B)
    tout = ctx.tensor(56, 56, 256, 0, D, capi.F16, B)
    sb = np.concatenate([rng.uniform(0.5, 1.5, 256), rng.uniform(-0.5, 0.5, 256)]).astype(np.float32)
    op = capi.BatchNorm(ctx, sb, width=56, height=56, channels=256, flags=capi.FLAG_DEEP | capi.FLAG_PRE_RELU)
    timed("deep batch-norm 56x56x256, batch 128", lambda: op.run(tin, tout), tbytes(tin) + tbytes(tout))
    ta = ctx.tensor(56, 56, 256, 0, D, capi.F16, B)
    add = capi.Arith(ctx, width=56, height=56, channels=256, op=capi.ARITH_ADD, flags=capi.FLAG_DEEP)
    timed("deep add 56x56x256 (two inputs), batch 128", lambda: add.run(tin, ta, tout), 3 * tbytes(tin))
    tcat = ctx.tensor(56, 56, 512, 0, D, capi.F16, B)
    cat = capi.Concat(ctx, width=56, height=56, channels=(256, 256), flags=capi.FLAG_DEEP)
    timed("deep concat 2 x 56x56x256, batch 128", lambda: cat.run([tin, ta], tcat), 4 * tbytes(tin))
    for o in (tin, tout, ta, tcat, op, add, cat):
        o.destroy()
    # ---- StyleNet shapes (1524x1856), batch 8 so that the working set exceeds L2
    W, H, NB = 1524, 1856, 8
    tin = ctx.tensor(W, H, 3, 0, S, capi.F16, NB)
    tout = ctx.tensor(W, H, 3, 0, S, capi.F16, NB)
    op = capi.Sigmoid(ctx, width=W, height=H, channels=3)
    timed("sigmoid 1524x1856x3 (shallow fp16 RGBA planes), batch 8", lambda: op.run(tin, tout), tbytes(tin) + tbytes(tout))
    op.destroy()
    stage = ctx.device_alloc(tbytes(tout) * 2)
    import ctypes as C
    L = capi.lib()
    timed("download widen fp16 RGBA -> fp32 RGBA (DownloadLayer), batch 8",
          lambda: capi.check(L.fyn_download_convert(tout._h, C.c_void_p(stage), None)), tbytes(tout) * 3)
    timed("download fp16 RGBA -> RGBA8 (byte download), batch 8",
          lambda: capi.check(L.fyn_download_u8_convert(tout._h, C.c_void_p(stage), None)), tbytes(tout) + tbytes(tout) // 2)
    ctx.device_free(stage)
    for o in (tin, tout):
        o.destroy()
    print(json.dumps({"note": "upload conversions run behind a host->device copy in the API (fyn_upload_*_async); they are timed with the rest of the "
                              "frame in bench.py's end-to-end figures"}))


if __name__ == "__main__":
    main()
