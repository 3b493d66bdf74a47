#!/usr/bin/env python
"""Which stage bounds the asynchronous end-to-end pipeline?  StyleNet 9x9 1524x1856 with upload and/or download layers."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from fyusenet_b200 import capi, hostapi, synthetic  # noqa: E402

W, H = 1524, 1856
weights = synthetic.stylenet_weights(9)
img = synthetic.image(H, W, 0)
for up, down in [(True, True), (True, False), (False, True)]:
    net = hostapi.StyleNet(9, W, H, upload=up, download=down)
    net.asynchronous()
    net.load_weights(weights)
    if not up:
        ctx = capi.Context(0)
        tin = ctx.tensor(W, H, 3, 0, capi.ORDER_SHALLOW, capi.F32, 1, packing=3)
        tin.upload(img)
        ctx.stream_sync()
        net.set_input_tensor(tin)
    net.setup()
    if up:
        for k in range(hostapi.async_slots()):
            net.input_buffer_slot(k)[:] = img.reshape(-1)
    for _ in range(5):
        net.forward()
    net.finish()
    t0 = time.perf_counter()
    n = 40
    for _ in range(n):
        net.forward()
    net.finish()
    dt = (time.perf_counter() - t0) / n
    print(f"upload={up} download={down}: {dt * 1e3:.3f} ms/frame ({1 / dt:.0f} frames/s)")
    net.destroy()
