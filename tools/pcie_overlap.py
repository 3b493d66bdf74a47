#!/usr/bin/env python
"""PCIe check for the end-to-end arm: H2D of one input frame (33.9 MB) and D2H of one output frame (45.3 MB), alone and
concurrently on two streams (pinned host memory).  python tools/pcie_overlap.py"""
import time

import torch

dev = torch.device("cuda", 0)
n_in, n_out = 1524 * 1856 * 3, 1524 * 1856 * 4
h_in = torch.empty(n_in, dtype=torch.float32).pin_memory()
h_out = torch.empty(n_out, dtype=torch.float32).pin_memory()
d_in = torch.empty(n_in, dtype=torch.float32, device=dev)
d_out = torch.empty(n_out, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, reps=30):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


for _ in range(2):
    a, b, c = run(True, False), run(False, True), run(True, True)
print(f"H2D 33.9 MB alone {a:.3f} ms ({33.94 / a:.1f} GB/s)   D2H 45.3 MB alone {b:.3f} ms ({45.26 / b:.1f} GB/s)   both concurrently {c:.3f} ms per pair")
