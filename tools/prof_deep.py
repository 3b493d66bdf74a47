#!/usr/bin/env python
"""Single deep-tiled convolution layers of ResNet-50 at batch B through the C ABI (timing, and the target of ncu captures):
    python tools/prof_deep.py <case|all> [batch] [reps]
cases: expand (1x1 64->256 + residual @56), reduce (1x1 256->64 @56), c3 (3x3 64->64 @56), expand28 (1x1 128->512 + residual @28),
       c3_14 (3x3 256->256 @14), reduce7 (1x1 2048->512 @7), stem (7x7 s2 3->64 @224)
Prints microseconds per launch and the fraction of the HBM roofline (input + output + residual + weights once)."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fyusenet_b200 import capi  # noqa: E402

CASES = {  # k, ds, ci, co, size, residual
    "expand": (1, 1, 64, 256, 56, True), "reduce": (1, 1, 256, 64, 56, False), "c3": (3, 1, 64, 64, 56, False),
    "expand28": (1, 1, 128, 512, 28, True), "c3_14": (3, 1, 256, 256, 14, False), "reduce7": (1, 1, 2048, 512, 7, False),
    "stem": (7, 2, 3, 64, 224, False), "reduce_bn": (1, 1, 256, 64, 56, False),
}


def run(name, batch, reps):
    ctx = capi.Context(0)
    rng = np.random.default_rng(0)
    k, ds, ci, co, size, res = CASES[name]
    pad = (k - 1) // 2 if name != "stem" else 1
    wb = np.concatenate([rng.uniform(-.1, .1, co), rng.normal(0, np.sqrt(2.0 / (k * k * ci)), co * k * k * ci), rng.uniform(0.5, 1.5, co), rng.uniform(-.1, .1, co)]).astype(np.float32)
    fl = capi.FLAG_DEEP | capi.FLAG_PRE_RELU | capi.FLAG_POST_BATCHNORM | (capi.FLAG_RESIDUAL_INPUT | capi.FLAG_RELU_ON_RESIDUAL if res else 0)
    op = capi.Conv2d(ctx, wb, width=size, height=size, in_channels=ci, out_channels=co, kernel=k, downsample=ds, in_padding=pad, flags=fl)
    if name.endswith("_bn"):    # the batch-norm layer in front of the convolution, evaluated at the fetch (fyn_conv2d_set_input_norm)
        op.set_input_norm(np.concatenate([rng.uniform(0.5, 1.5, ci), rng.uniform(-0.5, 0.5, ci)]).astype(np.float32))
    order = capi.ORDER_DEEP
    tin = ctx.tensor(size, size, ci, pad, order, capi.F16, batch)
    so = size // ds
    tout = ctx.tensor(so, so, co, 0, order, capi.F16, batch)
    tres = ctx.tensor(so, so, co, 0, order, capi.F16, batch) if res else None
    nbytes = tin.geom.bytes + tout.geom.bytes + (tres.geom.bytes if res else 0) + co * k * k * ci * 2
    flops = 2.0 * batch * so * so * co * k * k * ci
    ctx.stream_sync()
    e0, e1 = ctx.event_create(), ctx.event_create()
    for _ in range(2):
        op.run(tin, tout, tres)
    ctx.stream_sync()
    ctx.event_record(e0)
    for _ in range(reps):
        op.run(tin, tout, tres)
    ctx.event_record(e1)
    ctx.event_sync(e1)
    us = ctx.elapsed_ms(e0, e1) / reps * 1e3
    bound_us = max(nbytes / 6551e9, flops / 1378.8e12) * 1e6
    print(json.dumps({"case": name, "batch": batch, "us": round(us, 1), "MB": round(nbytes / 1e6, 1), "GB/s": round(nbytes / us / 1e3, 1),
                      "roofline_us": round(bound_us, 1), "frac": round(bound_us / us, 3)}), flush=True)
    for o in (op, tin, tout, tres):
        if o is not None:
            o.destroy()


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "all"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    for n in (CASES if name == "all" else name.split(",")):
        run(n, batch, reps)
