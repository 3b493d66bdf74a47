#!/usr/bin/env python
"""StyleNet's residual trunk (ten 3x3 40->40 layers at 381x464, clamp-to-edge tensors) through the C ABI: the persistent chain
kernel (fyn_conv_chain) against the ten single-layer launches.  python tools/prof_chain.py [reps] [layers] [w] [h]
With FYN_B200_LIB=fyusenet_b200/lib_prof/libfyusenet_b200.so (make -C fyusenet_b200/csrc PROF=1) the kernel prints its role profiles."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fyusenet_b200 import capi  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    nl = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    w = int(sys.argv[3]) if len(sys.argv) > 3 else 381
    h = int(sys.argv[4]) if len(sys.argv) > 4 else 464
    ch = 40
    ctx = capi.Context(0)
    rng = np.random.default_rng(0)
    ops, rfrom = [], []
    for i in range(nl):
        fl = capi.FLAG_PRE_RELU if i != 2 else 0
        r = -2
        if i % 2 == 1:
            fl |= capi.FLAG_RESIDUAL_INPUT | (capi.FLAG_RELU_ON_RESIDUAL if i == 1 else 0)
            r = i - 2
        wb = np.concatenate([rng.uniform(-.1, .1, ch), rng.normal(0, np.sqrt(1.0 / (9 * ch)), ch * 9 * ch)]).astype(np.float32)
        ops.append(capi.Conv2d(ctx, wb, width=w, height=h, in_channels=ch, out_channels=ch, kernel=3, flags=fl, backend=capi.BACKEND_TC))
        rfrom.append(r)
    tens = [ctx.tensor(w, h, ch, 0, capi.ORDER_SHALLOW, capi.F16, 1) for _ in range(nl + 1)]
    tens[0].write_chw(rng.normal(size=(ch, h, w)).astype(np.float32))
    tout = ctx.tensor(w, h, ch, 0, capi.ORDER_SHALLOW, capi.F16, 1)
    chain = capi.ConvChain(ctx, ops, rfrom)
    e0, e1 = ctx.event_create(), ctx.event_create()

    def single():
        for i, op in enumerate(ops):
            op.run(tens[i], tens[i + 1], tens[rfrom[i] + 1] if rfrom[i] >= -1 else None)

    def chained():
        assert chain.run(tens[0], tout)

    for name, fn in (("single-layer launches", single), ("chain kernel", chained)):
        for _ in range(3):
            fn()
        ctx.stream_sync()
        ctx.event_record(e0)
        for _ in range(reps):
            fn()
        ctx.event_record(e1)
        ctx.event_sync(e1)
        ms = ctx.elapsed_ms(e0, e1) / reps
        print(f"{name}: {ms * 1e3:.1f} us per trunk of {nl} layers ({ms * 1e3 / nl:.2f} us per layer)")
    same = np.array_equal(tens[-1].read_chw(), tout.read_chw())
    print("bit-identical:", same)


if __name__ == "__main__":
    main()
