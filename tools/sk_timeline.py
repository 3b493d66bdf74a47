#!/usr/bin/env python
"""In-kernel timeline of the split-K cluster kernels (k_conv_deep_tc_sk) of one ResNet-50 batch-1 forward.

Needs the instrumented build:  make -C fyusenet_b200/csrc EXTRA=-DFYN_SK_TIMELINE   (touch fyn_conv_deep_tc.cu first; rebuild without EXTRA afterwards)
Every launch writes %globaltimer stamps into a device buffer (min start / max end over its CTAs, the phases of CTA (0,0,0)'s thread 0);
this script runs three eager forwards and prints one line per launch of the last one (ns).  Evidence: profiles/r02_resnet_b1_sk_timeline.txt."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent.parent))
from fyusenet_b200 import capi, hostapi
ctx = capi.Context(0)
net = hostapi.ResNet50(device=0, batch=1)
net.load_weights((np.random.default_rng(50).standard_normal(net.weight_floats) * 0.02).astype(np.float32))
net.setup()
net.input_buffer()[:] = np.random.default_rng(1).random(224 * 224 * 3, dtype=np.float32)
net.forward(); net.skip_io(True)
for _ in range(5): net.forward()
net.finish()
L = capi.lib()
L.fyn_debug_sk_timeline(None, 1)
for _ in range(3): net.forward()
net.finish()
buf = np.zeros((2048, 16), np.uint64)
L.fyn_debug_sk_timeline(buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), 0)
rows = buf[buf[:, 1] > 0]
rows = rows[-53:]
t0 = rows[0, 0]
prev_end = None
print("start  end | kernel_len  gap_from_prev_end | firstwait-start | cta0: setup wait gather mma reduce epi")
for r in rows:
    st, en = int(r[0] - t0), int(r[1] - t0)
    ts = r[2:9].astype(np.int64)
    d = np.diff(ts)
    print(f"{st:7d} {en:7d} | {en-st:6d} {'' if prev_end is None else st-prev_end:>6} | {int(r[9]-r[0]):6d} | " + " ".join(f"{int(x):5d}" for x in d) + " || pre %d issue %d data %d cvt+sts %d fence %d arrive %d" % (int(r[10]-r[4]), int(r[11]-r[10]), int(r[12]-r[11]), int(r[13]-r[12]), int(r[14]-r[13]), int(r[15]-r[14])))
    prev_end = en
print("total", int(rows[-1,1]-rows[0,0]))
