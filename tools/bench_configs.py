#!/usr/bin/env python
"""Device-resident throughput of the other BASELINE.json configs on one GPU (bench.py covers configs[1]):
StyleNet3x3 @512x624 (configs[0]), StyleNet9x9 @4096x4096 (C5, single GPU), ResNet-50 224x224 batch 1 / 32 (C3 / C4 slice).
python tools/bench_configs.py"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fyusenet_b200 import capi, hostapi, synthetic  # noqa: E402


def stylenet(ksize, w, h, steps):
    ctx = capi.Context(0)
    net = hostapi.StyleNet(ksize, w, h, upload=False, download=False)
    net.load_weights(synthetic.stylenet_weights(ksize))
    tin = ctx.tensor(w, h, 3, 0, capi.ORDER_SHALLOW, capi.F32, 1, packing=3)
    tin.upload(synthetic.image(h, w, 0))
    ctx.stream_sync()
    net.set_input_tensor(tin)
    net.setup()
    for _ in range(5):
        net.forward()
    net.finish()
    e0, e1 = ctx.event_create(), ctx.event_create()
    ctx.event_record(e0, net.stream)
    for _ in range(steps):
        net.forward()
    ctx.event_record(e1, net.stream)
    net.finish()
    ms = ctx.elapsed_ms(e0, e1) / steps
    fams = sorted({l["family"] for l in net.layers() if l["family"]})
    net.destroy()
    return {"workload": f"StyleNet{ksize}x{ksize} {w}x{h}", "frames_per_s": 1e3 / ms, "ms_per_frame": ms, "kernel_families": fams}


def resnet(batch, steps):
    # weights: random values of the right file size are enough for a timing run (parity is covered by the tests)
    rng = np.random.default_rng(0)
    net = hostapi.ResNet50(batch=batch)
    net.load_weights((rng.standard_normal(25576046) * 0.01).astype(np.float32))
    net.setup()
    net.set_input(rng.random((batch, 224, 224, 3), dtype=np.float32))
    net.forward()
    t0 = time.perf_counter()
    for _ in range(steps):
        net.forward()            # synchronous API: upload + layers + download
    dt = (time.perf_counter() - t0) / steps
    # device-resident figure: the same step without the two PCIe layers, from per-layer CUDA events of a second pass
    net.enable_timings(True)
    for _ in range(2):
        net.forward()
    net.finish()
    io_ms = sum(net.layer_timing(l["number"])[0] / 2 for l in net.layers() if l["name"] in ("upload", "download"))
    all_ms = sum(net.layer_timing(l["number"])[0] / 2 for l in net.layers())
    net.destroy()
    return {"workload": f"ResNet-50 224x224 batch {batch} (end to end, synchronous API)", "img_per_s": batch / dt, "ms_per_step": dt * 1e3,
            "upload_download_ms": io_ms, "img_per_s_device_resident": batch / ((all_ms - io_ms) * 1e-3),
            "note": "device-resident = sum of the per-layer CUDA-event times without the upload / download layers"}


if __name__ == "__main__":
    import faulthandler
    faulthandler.dump_traceback_later(40, exit=True)        # a stuck configuration reports where instead of hanging the box
    which = sys.argv[1:] if len(sys.argv) > 1 else ["all"]
    jobs = {"s3": lambda: stylenet(3, 512, 624, 200), "s9": lambda: stylenet(9, 1524, 1856, 50), "s9_4096": lambda: stylenet(9, 4096, 4096, 10),
            "r1": lambda: resnet(1, 20), "r32": lambda: resnet(32, 5), "r128": lambda: resnet(128, 3), "r512": lambda: resnet(512, 2)}
    which = [k for w in which for k in (["s3", "s9", "s9_4096", "r1", "r32"] if w == "all" else [w])]
    for k in which:
        print(json.dumps(jobs[k]()), flush=True)
