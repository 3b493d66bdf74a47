#!/usr/bin/env python
"""Runs single StyleNet-9x9 layers at the headline size (1524x1856) through the C ABI; used under ncu and for
quick per-layer timing:  python tools/prof_layers.py <layer> [reps] [backend]   (layer: conv1 conv2 conv3 res res2
deconv1 deconv2 deconv3 sigmoid all)."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fyusenet_b200 import capi  # noqa: E402

W, H = 1524, 1856
LAYERS = {
    #            k  ci  co  div ds frac step  relu  res
    "conv1":    (9, 3, 12, 1, 1, False, 1.0, True, False),
    "conv2":    (3, 12, 20, 1, 2, False, 1.0, True, False),
    "conv3":    (3, 20, 40, 2, 2, False, 1.0, True, False),
    "res":      (3, 40, 40, 4, 1, False, 1.0, True, False),
    "res2":     (3, 40, 40, 4, 1, False, 1.0, True, True),
    "deconv1":  (3, 40, 20, 4, 2, True, 0.5, False, False),
    "deconv2":  (3, 20, 12, 4, 2, True, 0.25, True, False),
    "deconv3":  (9, 12, 3, 2, 1, True, 0.5, True, False),
}


def run(name, reps, backend):
    ctx = capi.Context(0)
    rng = np.random.default_rng(0)
    k, ci, co, div, ds, frac, step, relu, res = LAYERS[name]
    w, h = W // div, H // div
    wb = np.concatenate([rng.uniform(-.1, .1, co), rng.normal(0, np.sqrt(2.0 / (k * k * ci)), co * k * k * ci)]).astype(np.float32)
    fl = (capi.FLAG_PRE_RELU if relu else 0) | (capi.FLAG_RESIDUAL_INPUT if res else 0)
    op = capi.Conv2d(ctx, wb, width=w, height=h, in_channels=ci, out_channels=co, kernel=k, downsample=ds, flags=fl,
                     source_step=step, fractional=frac, backend=backend)
    if name == "conv1":
        tin = ctx.tensor(w, h, 3, 0, capi.ORDER_SHALLOW, capi.F32, 1, packing=3)
        tin.upload(rng.random((h, w, 3), dtype=np.float32))
    else:
        tin = ctx.tensor(w, h, ci)
        tin.write_chw(rng.normal(size=(ci, h, w)).astype(np.float32))
    tout = ctx.tensor(op.out_width, op.out_height, co)
    tres = None
    if res:
        tres = ctx.tensor(op.out_width, op.out_height, co)
        tres.write_chw(rng.normal(size=(co, op.out_height, op.out_width)).astype(np.float32))
    ctx.stream_sync()
    e0, e1 = ctx.event_create(), ctx.event_create()
    for _ in range(3):
        op.run(tin, tout, tres)
    ctx.stream_sync()
    ctx.event_record(e0)
    for _ in range(reps):
        op.run(tin, tout, tres)
    ctx.event_record(e1)
    ctx.event_sync(e1)
    ms = ctx.elapsed_ms(e0, e1) / reps
    print(f"{name:8s} backend={op.backend} {ms * 1e3:9.1f} us/launch")


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "all"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    backend = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    for n in (LAYERS if name == "all" else [name]):
        run(n, reps, backend)
