#!/usr/bin/env python
"""ResNet-50 device-resident timing through the host engine: total per forward (CUDA events around `reps` forwards with the
upload / download layers skipped) and, with `layers`, the per-layer event times of one more pass.

    python tools/prof_resnet.py [batch] [reps] [layers]        env: FYN_DEEP_PERSIST=0 -> one-tile-per-CTA kernel"""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fyusenet_b200 import capi, hostapi  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    ctx = capi.Context(0)
    net = hostapi.ResNet50(device=0, batch=batch)
    net.load_weights((np.random.default_rng(50).standard_normal(net.weight_floats) * 0.02).astype(np.float32))
    net.setup()
    net.input_buffer()[:] = np.random.default_rng(1).random(batch * 224 * 224 * 3, dtype=np.float32)
    net.forward()
    net.skip_io(True)
    if os.environ.get("FYN_PROF_GRAPH"):
        net.enable_graph(True)
    for _ in range(4):
        net.forward()
    net.finish()
    e0, e1 = ctx.event_create(), ctx.event_create()
    ctx.event_record(e0, net.stream)
    for _ in range(reps):
        net.forward()
    ctx.event_record(e1, net.stream)
    ctx.event_sync(e1)
    ms = ctx.elapsed_ms(e0, e1) / reps
    out = {"batch": batch, "ms": round(ms, 3), "img_per_s": round(batch / ms * 1e3, 1), "persist": os.environ.get("FYN_DEEP_PERSIST", "default"), "graph": net.graph_active}
    if len(sys.argv) > 3:
        net.enable_timings(True)
        for _ in range(2):
            net.forward()
        net.finish()
        rows = sorted(((net.layer_timing(l["number"])[0] / 2 * 1e3, l["name"]) for l in net.layers()), reverse=True)
        out["top_layers_us"] = [(n, round(t, 1)) for t, n in rows[:60]]
    print(json.dumps(out))
    net.destroy()


if __name__ == "__main__":
    main()
