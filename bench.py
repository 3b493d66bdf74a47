#!/usr/bin/env python
"""bench.py -- headline benchmark of the FyuseNet-on-B200 hot path.

    python bench.py --gpus N --steps K --warmup W            # this backend
    python bench.py --impl reference --gpus N --steps K ...   # CPU baseline arm (oracle port, host cores)

Workload (BASELINE.json configs[1]): StyleNet 9x9, one 1524x1856 RGB frame per step, synthetic weights/images
(the reference's data/*.dat are git-LFS stubs).  A "step" = one full forward pass of the network on one frame.

  value  : frames/s with the input frame resident in HBM (network built without upload / download layers),
           timed with CUDA events on the network's stream over exactly K steps after W warm-ups.
  e2e    : frames/s through the reference-facing API (StyleNet9x9::forward with upload + download layers):
           every step copies the frame from pinned host memory to the device and reads the RGBA result back.
  N > 1  : one process per GPU (torchrun), frame-level replicas -- the path does not shard below a frame at this
           size ("replicas only"); no data-path collective; weak scaling; time = max over ranks.

One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
for p in (ROOT, ROOT / "oracle"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

WIDTH, HEIGHT, KSIZE = 1524, 1856, 9
METRIC = "stylenet9x9_1524x1856_frames_per_s"

# algorithmic work per frame (SURVEY 8d / BASELINE.md section 2; fp16 storage, channels padded to 4)
FRAME_GFLOP = 95.11
FRAME_MB = 713.1


def layer_algorithmic(ksize=KSIZE, w=WIDTH, h=HEIGHT):
    """Per-layer algorithmic FLOPs and bytes (input read once + output written once + residual read once +
    weights once; fp16 activations, fp32 RGB upload texture for conv1, channels padded to multiples of 4)."""
    def pad4(c):
        return 4 * ((c + 3) // 4)
    L = {}
    def conv(name, k, ci, co, wi, hi, wo, ho, res=False, in_bytes_per_px=None):
        flops = 2.0 * k * k * ci * co * wo * ho
        inb = wi * hi * (in_bytes_per_px if in_bytes_per_px else pad4(ci) * 2)
        outb = wo * ho * pad4(co) * 2
        wb = (co + k * k * ci * co) * 4
        L[name] = dict(flops=flops, bytes=inb + outb + wb + (outb if res else 0))
    conv("conv1", ksize, 3, 12, w, h, w, h, in_bytes_per_px=12)
    conv("conv2", 3, 12, 20, w, h, w // 2, h // 2)
    conv("conv3", 3, 20, 40, w // 2, h // 2, w // 4, h // 4)
    for r in range(1, 6):
        conv(f"res{r}_1", 3, 40, 40, w // 4, h // 4, w // 4, h // 4)
        conv(f"res{r}_2", 3, 40, 40, w // 4, h // 4, w // 4, h // 4, res=True)
    conv("deconv1", 3, 40, 20, w // 4, h // 4, w // 4, h // 4)
    conv("deconv2", 3, 20, 12, w // 4, h // 4, w // 2, h // 2)
    conv("deconv3", ksize, 12, 3, w // 2, h // 2, w, h)
    L["sigmoid"] = dict(flops=0.0, bytes=2 * w * h * 4 * 2)
    return L


class ClockSampler:
    """Samples SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md's clocks line).  The timed region of the
    headline is a few milliseconds, shorter than one period of `nvidia-smi -lms`, so the samples come from NVML in this
    process (the same counters nvidia-smi prints: clocks.sm, clocks.max.sm, power.draw, clocks_event_reasons.*), every 2 ms
    from a thread; `nvidia-smi -lms 100` is the fallback when NVML cannot be loaded.  rows: [index, sm, max sm, power W, ...]."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, device_index: int):
        self.idx = device_index
        self.rows = []
        self.proc = None
        self.nvml = None
        self.t = None
        self._stop = threading.Event()

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        # CUDA_VISIBLE_DEVICES renumbers the devices NVML sees in board order: go through the PCI bus id when torch knows it
        try:
            import torch
            bus = torch.cuda.get_device_properties(self.idx).pci_bus_id
            dom = torch.cuda.get_device_properties(self.idx).pci_domain_id
            dev = torch.cuda.get_device_properties(self.idx).pci_device_id
            return pynvml, pynvml.nvmlDeviceGetHandleByPciBusId(f"{dom:08x}:{bus:02x}:{dev:02x}.0")
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.idx)

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.source = "nvml"
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.source = "nvidia-smi"
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        mx = float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM))
        bits = [n.nvmlClocksEventReasonHwSlowdown, n.nvmlClocksEventReasonHwThermalSlowdown, n.nvmlClocksEventReasonSwThermalSlowdown,
                n.nvmlClocksEventReasonSwPowerCap]
        it, watts = 0, float("nan")
        while True:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                if it % 8 == 1:     # (the power query is the slow one: every 8th sample, and not the first)
                    try:
                        watts = n.nvmlDeviceGetPowerUsage(self.handle) / 1e3
                    except Exception:
                        watts = float("nan")
                it += 1
                mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                self.rows.append([str(self.idx), sm, mx, watts, hex(mask)] + ["Active" if mask & b else "Not Active" for b in bits])
            except Exception:
                pass
            if self._stop.wait(0.002):
                return

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.t.join(timeout=2)
        elif self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"], "samples": 0}
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nme, v in zip(self.NAMES, r[5:9]):
                    if str(v).lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def cpu_port_frames_per_s(weights, rows: int, repeats: int = 1):
    """Times the oracle port (oracle/fyn_oracle.c, OpenMP over all host cores, fp32) on a `rows`-high band of the
    1524-wide frame and scales to full frames.  Returns (frames/s, seconds per sample, sample description)."""
    import fyn_oracle as fo
    img = fo.synthetic_image(rows, WIDTH, 0)
    t0 = time.perf_counter()
    for _ in range(repeats):
        fo.stylenet_forward(weights, img, KSIZE, prec=fo.FP32)
    dt = (time.perf_counter() - t0) / repeats
    frac = rows / HEIGHT
    return frac / dt, dt, f"{WIDTH}x{rows} band ({frac:.3f} frame), fp32, {repeats} run(s)"


def run_reference(args, rank, world):
    """--impl reference: the CPU arm.  The reference's GL shader path cannot run on this box (no EGL/Mesa),
    so this times the oracle port of the same path on all host cores.  A step is a full frame when K+W full frames fit
    in ~3 minutes, otherwise a horizontal band of the frame (scaled to frames)."""
    if rank != 0:
        return
    import fyn_oracle as fo
    fo.lib()
    weights = fo.stylenet_synthetic_weights(KSIZE)
    cores = fo.set_num_threads(os.cpu_count() or 1)      # all host threads (torchrun exports OMP_NUM_THREADS=1)
    _, dt_q, _ = cpu_port_frames_per_s(weights, 464)            # quarter frame: estimates the frame time
    t_frame = 4.0 * dt_q
    rows = HEIGHT
    budget = 180.0
    if (args.steps + args.warmup) * t_frame > budget:
        rows = max(64, int(HEIGHT * budget / ((args.steps + args.warmup) * t_frame)) // 4 * 4)
    for _ in range(args.warmup):
        cpu_port_frames_per_s(weights, rows)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_port_frames_per_s(weights, rows)
    dt = time.perf_counter() - t0
    fps = args.steps * (rows / HEIGHT) / dt
    sample = f"{args.steps} x {WIDTH}x{rows} ({rows / HEIGHT:.3f} frame each), OpenMP oracle port, fp32 (reference GL path not runnable: no EGL/Mesa)"
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "StyleNet 9x9 (stylenet9x9 layout, synthetic He weights) 1524x1856 RGB frame, BASELINE configs[1]",
                       "step": f"{WIDTH}x{rows} rows per step, scaled to frames"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# secondary workloads: the other half of BASELINE.json's metric (ResNet-50 img/s) and the sharded configs C4 / C5.
# They ride in the `secondary` block of the same JSON line; the headline metric stays StyleNet 9x9 @1524x1856.
# ----------------------------------------------------------------------------------------------------------------------
RESNET_B512_ROOFLINE_IMG_S = 77700.0     # 6.59 ms per 512 images on one GPU
RESNET_B1_ROOFLINE_IMG_S = 51100.0
S9_4096_ROOFLINE_FPS = 1477.0


def _timed_wall(fn, steps, barrier):
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    return time.perf_counter() - t0


def secondary_resnet_sharded(ctx, comm, rank, world, local_rank, barrier, max_over_ranks, total=512, steps=5, warmup=3):
    """BASELINE configs[3]: ResNet-50 224x224, 512 images sharded contiguously over the ranks (strong scaling), weights
    replicated, one NCCL all-gather of the [512/world, 1000] logits per step INSIDE the timed region (fyn_allgather_logits).
    Device-resident: images resident in HBM (upload / download layers skipped), logits gathered into device memory.
    e2e: every step uploads this rank's images from pinned host memory, runs the layers, gathers and copies the [512, 1000]
    logits to pinned host memory."""
    from fyusenet_b200 import hostapi, multigpu
    b, e = multigpu.shard_range(total, rank, world)
    n_local = e - b
    nmax = (total + world - 1) // world
    rng = np.random.default_rng(50 + rank)
    net = hostapi.ResNet50(device=local_rank, batch=n_local)
    net.load_weights((np.random.default_rng(50).standard_normal(net.weight_floats) * 0.02).astype(np.float32))
    net.setup()
    net.input_buffer()[:] = rng.random(n_local * 224 * 224 * 3, dtype=np.float32)
    gathered = ctx.device_alloc(world * nmax * 1000 * 4)
    host_logits = ctx.host_alloc(world * nmax * 1000)
    logits_t = net.layer_tensor(72)
    stream = net.stream

    def step_e2e():
        net.forward()                                                   # H2D of the images + 70 layers + D2H of the local logits + sync
        comm.allgather_logits(logits_t, nmax, gathered, stream)
        ctx.memcpy_d2h(host_logits, gathered, stream)
        ctx.stream_sync(stream)

    def step_dev():
        net.forward()                                                   # layers only (skip_io)
        comm.allgather_logits(logits_t, nmax, gathered, stream)

    for _ in range(warmup):
        step_e2e()
    e2e_s = max_over_ranks(_timed_wall(step_e2e, steps, barrier))
    finite = bool(np.isfinite(host_logits).all())
    net.skip_io(True)
    for _ in range(2):
        step_dev()
    ctx.stream_sync(stream)
    ev0, ev1 = ctx.event_create(), ctx.event_create()
    barrier()
    ctx.event_record(ev0, stream)
    for _ in range(steps):
        step_dev()
    ctx.event_record(ev1, stream)
    ctx.event_sync(ev1)
    dev_ms = max_over_ranks(ctx.elapsed_ms(ev0, ev1))
    barrier()
    net.destroy()
    # the same end-to-end step with 8-bit images (ResNet50::setByteInput: value / 255 on the device, what the reference's sample does
    # on the host): 3 instead of 12 bytes per pixel over PCIe
    byte_e2e = None
    try:
        bnet = hostapi.ResNet50(device=local_rank, batch=n_local)
        bnet.set_byte_input(True)
        bnet.load_weights((np.random.default_rng(50).standard_normal(bnet.weight_floats) * 0.02).astype(np.float32))
        bnet.setup()
        bnet.input_buffer()[:] = rng.integers(0, 256, n_local * 224 * 224 * 3, dtype=np.uint8)
        blogits = bnet.layer_tensor(72)
        bstream = bnet.stream

        def step_bytes():
            bnet.forward()
            comm.allgather_logits(blogits, nmax, gathered, bstream)
            ctx.memcpy_d2h(host_logits, gathered, bstream)
            ctx.stream_sync(bstream)

        for _ in range(warmup):
            step_bytes()
        b_s = max_over_ranks(_timed_wall(step_bytes, steps, barrier))
        byte_e2e = {"value": total * steps / b_s, "unit": "img/s", "ms_per_step": 1e3 * b_s / steps, "h2d_bytes_per_step": int(n_local * 224 * 224 * 3),
                    "d2h_bytes_per_step": int(n_local * 1000 * 4 + world * nmax * 1000 * 4), "finite": bool(np.isfinite(host_logits).all()),
                    "api": "ResNet50::setByteInput(): uint8 RGB images in (value / 255 on the device)"}
        bnet.destroy()
    except Exception as exc:  # noqa: BLE001  (the float path above is the contract; this is an extra)
        byte_e2e = {"error": str(exc)[:200]}
    ctx.device_free(gathered)
    value = total * steps / (dev_ms / 1e3)
    return {"workload": f"ResNet-50 224x224, {total} images batch-sharded over {world} GPU(s), NCCL all-gather of the logits inside the timed region (BASELINE configs[3])",
            "scaling": "strong", "value": value, "unit": "img/s", "ms_per_step": dev_ms / steps, "steps": steps, "images_per_rank": n_local,
            "collective": f"fyn_allgather_logits: NCCL all-gather of float32 [{nmax}, 1000] per rank" if world > 1 else "none (one rank): logits converted to float32 on the device",
            "roofline_frac": value / (RESNET_B512_ROOFLINE_IMG_S * world),
            "e2e": {"value": total * steps / e2e_s, "unit": "img/s", "ms_per_step": 1e3 * e2e_s / steps,
                    "h2d_bytes_per_step": int(n_local * 224 * 224 * 3 * 4), "d2h_bytes_per_step": int(n_local * 1000 * 4 + world * nmax * 1000 * 4), "finite": finite,
                    "byte_io": byte_e2e}}


def secondary_resnet_b1(ctx, rank, world, local_rank, barrier, max_over_ranks, steps=200, warmup=20):
    """BASELINE configs[2]: ResNet-50 batch 1 (replicas for N > 1).  59 launches per image: the device layers are replayed from
    a CUDA graph (Engine::enableGraph); the eager figure is reported beside it."""
    from fyusenet_b200 import hostapi
    net = hostapi.ResNet50(device=local_rank, batch=1)
    net.load_weights((np.random.default_rng(50).standard_normal(net.weight_floats) * 0.02).astype(np.float32))
    net.setup()
    net.input_buffer()[:] = np.random.default_rng(7 + rank).random(224 * 224 * 3, dtype=np.float32)
    out = {}
    for mode in ("eager", "graph"):
        net.skip_io(False)
        net.enable_graph(mode == "graph")
        for _ in range(warmup):
            net.forward()
        e2e_s = max_over_ranks(_timed_wall(net.forward, steps, barrier))
        net.skip_io(True)
        for _ in range(3):
            net.forward()
        net.finish()
        ev0, ev1 = ctx.event_create(), ctx.event_create()
        barrier()
        ctx.event_record(ev0, net.stream)
        for _ in range(steps):
            net.forward()
        ctx.event_record(ev1, net.stream)
        ctx.event_sync(ev1)
        dev_ms = max_over_ranks(ctx.elapsed_ms(ev0, ev1))
        out[mode] = {"value": world * steps / (dev_ms / 1e3), "ms_per_step": dev_ms / steps, "e2e_value": world * steps / e2e_s,
                     "e2e_ms_per_step": 1e3 * e2e_s / steps, "graph_active": net.graph_active}
    finite = bool(np.isfinite(net.logits()).all())
    net.destroy()
    best = "graph" if out["graph"]["value"] >= out["eager"]["value"] else "eager"
    return {"workload": f"ResNet-50 224x224 batch 1 (BASELINE configs[2]), {'replicas x' + str(world) if world > 1 else 'single GPU'}",
            "scaling": "weak", "value": out[best]["value"], "unit": "img/s", "ms_per_step": out[best]["ms_per_step"], "mode": best, "steps": steps,
            "roofline_frac": out[best]["value"] / (RESNET_B1_ROOFLINE_IMG_S * world), "eager": out["eager"], "graph": out["graph"],
            "e2e": {"value": out[best]["e2e_value"], "unit": "img/s", "ms_per_step": out[best]["e2e_ms_per_step"],
                    "h2d_bytes_per_step": 224 * 224 * 3 * 4, "d2h_bytes_per_step": 1008 * 4, "finite": finite}}


def secondary_stylenet_bands(ctx, comm, rank, world, local_rank, barrier, max_over_ranks, size=4096, steps=10, warmup=3):
    """BASELINE configs[4]: StyleNet 9x9 on one 4096x4096 frame, row-banded over the ranks (strong scaling) with the per-layer
    halo exchange over NVLink (fyn_halo_exchange: peer stores, one kernel per layer); one rank runs the whole frame."""
    from fyusenet_b200 import hostapi, multigpu, synthetic
    margin = int(os.environ.get("FYN_HALO_MARGIN", multigpu.HALO_MARGIN_SPARSE))
    ib, ie, skip, keep = multigpu.stylenet_halo_band_plan(size, world, margin)[rank]
    h = ie - ib
    net = hostapi.StyleNet(KSIZE, size, h, upload=True, download=True, device=local_rank)
    net.load_weights(synthetic.stylenet_weights(KSIZE))
    net.setup()
    exchanges = 0
    if world > 1:
        net.set_halo_exchange(comm, margin, h)
        exchanges = net.halo_exchanges
    chained = net.chained_layers
    # this rank's rows of the synthetic frame (generated per rank: only the shape matters for the timing)
    net.input_buffer()[:] = np.random.default_rng(4096 + rank).random(h * size * 3, dtype=np.float32)
    for _ in range(warmup):
        net.forward()
    e2e_s = max_over_ranks(_timed_wall(net.forward, steps, barrier))       # band upload + layers + exchanges + band download per step
    out_bytes = int(net.output_rgba()[0].nbytes)
    finite = bool(np.isfinite(net.output_rgba()[0]).all())
    net.skip_io(True)
    for _ in range(2):
        net.forward()
    net.finish()
    ev0, ev1 = ctx.event_create(), ctx.event_create()
    barrier()
    ctx.event_record(ev0, net.stream)
    for _ in range(steps):
        net.forward()
    ctx.event_record(ev1, net.stream)
    ctx.event_sync(ev1)
    dev_ms = max_over_ranks(ctx.elapsed_ms(ev0, ev1))
    barrier()
    pushed = comm.info()["halo_bytes_pushed"] if world > 1 else 0
    frames = warmup + steps + 2 + steps
    net.destroy()
    value = steps / (dev_ms / 1e3)
    return {"workload": f"StyleNet 9x9 {size}x{size} frame, {world} row band(s) of {keep} rows + {margin}-row margins, per-layer halo exchange (BASELINE configs[4])",
            "scaling": "strong", "value": value, "unit": "frames/s", "ms_per_step": dev_ms / steps, "steps": steps,
            "exchange": f"fyn_halo_exchange: peer stores over NVLink (CUDA IPC), {exchanges} exchange(s) per frame where a layer would reach spoilt margin rows (Engine::planHalo)" if world > 1 else "none (one rank)",
            "exchanges_per_frame": exchanges, "chained_layers": chained,
            "nvlink_bytes_pushed_per_frame_rank0": int(pushed // max(frames, 1)),
            "roofline_frac": value / (S9_4096_ROOFLINE_FPS * world),
            "e2e": {"value": steps / e2e_s, "unit": "frames/s", "ms_per_step": 1e3 * e2e_s / steps,
                    "h2d_bytes_per_step": int(h * size * 3 * 4), "d2h_bytes_per_step": out_bytes, "finite": finite}}


def sustained_run(net, ctx, local_rank, seconds=3.0):
    """>= `seconds` of back-to-back device-resident forwards with the clock sampler running: the sustained figure next to the
    burst `value` (the kernels are issue / latency bound, i.e. clock sensitive; this box settles well below its boost clock)."""
    sampler = ClockSampler(local_rank)
    # calibrate the chunk so that the host stays ahead of the device without queueing minutes of work
    ev0, ev1 = ctx.event_create(), ctx.event_create()
    sampler.start()
    t_start = time.perf_counter()
    frames, dev_ms = 0, 0.0
    while time.perf_counter() - t_start < seconds:
        ctx.event_record(ev0, net.stream)
        for _ in range(500):
            net.forward()
        ctx.event_record(ev1, net.stream)
        ctx.event_sync(ev1)
        dev_ms += ctx.elapsed_ms(ev0, ev1)
        frames += 500
    clocks = sampler.stop()
    power = []
    for r in sampler.rows:
        try:
            if np.isfinite(float(r[3])):
                power.append(float(r[3]))
        except Exception:
            pass
    return {"value": frames / (dev_ms / 1e3), "unit": "frames/s", "frames": frames, "seconds": dev_ms / 1e3, "ms_per_step": dev_ms / frames,
            "clocks": clocks, "power_w_median": float(np.median(power)) if power else None, "power_w_max": max(power) if power else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the ResNet-50 / banded StyleNet / sustained blocks")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from fyusenet_b200 import capi, hostapi, synthetic      # the product arm never imports oracle/

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = None
    if world > 1 and os.environ.get("FYN_BENCH_NO_NUMA_BIND") is None:
        from fyusenet_b200 import multigpu
        numa = multigpu.bind_to_gpu_numa_node(local_rank)      # before any pinned buffer is allocated
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    weights = synthetic.stylenet_weights(KSIZE)
    img = synthetic.image(HEIGHT, WIDTH, rank)

    # ------------------------------------------------------------------ device-resident arm ("value")
    ctx = capi.Context(local_rank)
    net = hostapi.StyleNet(KSIZE, WIDTH, HEIGHT, upload=False, download=False, device=local_rank)
    net.load_weights(weights)
    tin = ctx.tensor(WIDTH, HEIGHT, 3, 0, capi.ORDER_SHALLOW, capi.F32, 1, packing=3)
    tin.upload(img)
    ctx.stream_sync()
    net.set_input_tensor(tin)
    net.setup()
    stream = net.stream
    for _ in range(args.warmup):
        net.forward()
    net.finish()
    ev0, ev1 = ctx.event_create(), ctx.event_create()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    lib_launch0 = _net_launches(net)
    ctx.event_record(ev0, stream)
    for _ in range(args.steps):
        net.forward()
    ctx.event_record(ev1, stream)
    net.finish()
    barrier()
    clocks = sampler.stop()
    ms = ctx.elapsed_ms(ev0, ev1)
    launches = _net_launches(net) - lib_launch0
    # per-layer times: a second pass with the engine's event pairs between the layers (they cost a few percent, so
    # they stay out of the timed region above)
    net.enable_timings(True)
    for _ in range(args.steps):
        net.forward()
    net.finish()
    layer_ms = {l["name"]: net.layer_timing(l["number"])[0] / args.steps for l in net.layers()}
    families = {l["name"]: l["family"] for l in net.layers()}
    net.enable_timings(False)
    chained = net.chained_layers
    # the dominant layer once more with an event pair around it alone: the other layers then keep their dependent-launch
    # overlap, which is the situation of the timed region above
    top_name = max(layer_ms, key=layer_ms.get)
    top_no = next(l["number"] for l in net.layers() if l["name"] == top_name)
    net.enable_layer_timing(top_no)
    for _ in range(args.steps):
        net.forward()
    net.finish()
    top_ms = net.layer_timing(top_no)[0] / args.steps
    net.enable_timings(False)
    # ... and as a kernel timed alone: the same layer (same shape, weights and input tensor layout) launched back to back
    # through the C ABI between two events -- no event pairs inside the stream, dependent launches overlap as in the frame
    iso_ms = None
    if top_name == "conv1":
        lay = synthetic.stylenet_file_layers(KSIZE)[0]
        nb = lay[3] + lay[3] * lay[1] * lay[1] * lay[2]
        op1 = capi.Conv2d(ctx, weights[:nb], width=WIDTH, height=HEIGHT, in_channels=3, out_channels=lay[3], kernel=lay[1], flags=capi.FLAG_PRE_RELU)
        t1 = ctx.tensor(WIDTH, HEIGHT, lay[3])
        for _ in range(5):
            op1.run(tin, t1)
        ctx.stream_sync()
        ea, eb = ctx.event_create(), ctx.event_create()
        ctx.event_record(ea)
        for _ in range(args.steps):
            op1.run(tin, t1)
        ctx.event_record(eb)
        ctx.event_sync(eb)
        iso_ms = ctx.elapsed_ms(ea, eb) / args.steps
        op1.destroy()
        t1.destroy()
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * args.steps / (ms / 1e3)
    sustained = None
    if not args.no_secondary:
        try:
            sustained = sustained_run(net, ctx, local_rank)
            if world > 1:
                t = torch.tensor([sustained["ms_per_step"]], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                sustained["ms_per_step"] = float(t.item())
                sustained["value"] = world * 1e3 / sustained["ms_per_step"]
        except Exception as exc:                                          # a secondary block never costs the headline line
            sustained = {"error": repr(exc)}

    # ------------------------------------------------------------------ end-to-end arm ("e2e")
    # (a) synchronous API, one frame at a time: setInputBuffer/forward/getOutputBuffer like samples/desktop/stylenet.cpp
    net2 = hostapi.StyleNet(KSIZE, WIDTH, HEIGHT, upload=True, download=True, device=local_rank)
    net2.load_weights(weights)
    net2.setup()
    inbuf = net2.input_buffer()            # pinned host memory owned by the network
    inbuf[:] = img.reshape(-1)
    for _ in range(args.warmup):
        net2.forward()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        net2.forward()                      # H2D copy of the frame + all layers + D2H of the RGBA result + sync
    net2.finish()
    sync_s = time.perf_counter() - t0
    out = net2.output_rgba()[0]
    finite = bool(np.isfinite(out).all())
    out_bytes = int(out.nbytes)
    net2.destroy()
    # (b) the reference's own throughput mechanism, NeuralNetwork::asynchronous(): forward() enqueues, <= 3 sequences in
    # flight; every step still uploads its frame from pinned host memory and downloads its RGBA result.
    net3 = hostapi.StyleNet(KSIZE, WIDTH, HEIGHT, upload=True, download=True, device=local_rank)
    net3.asynchronous()
    net3.load_weights(weights)
    net3.setup()
    for k in range(hostapi.async_slots()):
        net3.input_buffer_slot(k)[:] = img.reshape(-1)
    for _ in range(args.warmup):
        net3.forward()
    net3.finish()
    done0 = net3.async_completed()[0]
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        net3.forward()
    net3.finish()                           # returns when the last download has been delivered
    e2e_s = time.perf_counter() - t0
    delivered = net3.async_completed()[0] - done0
    net3.destroy()
    # (c) the same pipeline with 8-bit frames both ways (StyleNetBase::setByteIO: UBYTE upload -- a reference feature,
    # gpu/uploadlayer.cpp:51-66 -- and RGBA8 download, the samples' host-side quantisation moved to the device): 3 + 4 instead
    # of 12 + 16 bytes per pixel over PCIe
    e2e8 = None
    try:
        net4 = hostapi.StyleNet(KSIZE, WIDTH, HEIGHT, upload=True, download=True, device=local_rank)
        net4.asynchronous()
        net4.set_byte_io(True)
        net4.load_weights(weights)
        net4.setup()
        img8 = np.clip(img * 255.0 + 0.5, 0, 255).astype(np.uint8)
        for k in range(hostapi.async_slots()):
            net4.input_buffer_slot(k)[:] = img8.reshape(-1)
        for _ in range(args.warmup):
            net4.forward()
        net4.finish()
        done8 = net4.async_completed()[0]
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            net4.forward()
        net4.finish()
        e2e8_s = time.perf_counter() - t0
        delivered8 = net4.async_completed()[0] - done8
        out8 = net4.output_rgba()[0]
        if world > 1:
            t = torch.tensor([e2e8_s], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e8_s = float(t[0].item())
        e2e8 = {"value": world * args.steps / e2e8_s, "unit": "frames/s", "ms_per_step": 1e3 * e2e8_s / args.steps,
                "h2d_bytes_per_step": int(img8.nbytes), "d2h_bytes_per_step": int(out8.nbytes), "delivered": int(delivered8),
                "nonzero": bool(out8[..., :3].any()),
                "api": "StyleNet9x9 asynchronous() + setByteIO(): uint8 RGB frame in (value / 255 on the device), uint8 RGBA frame out"}
        net4.destroy()
    except Exception as exc:
        e2e8 = {"error": repr(exc)}
    if world > 1:
        t = torch.tensor([e2e_s, sync_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s, sync_s = float(t[0].item()), float(t[1].item())
    barrier()
    e2e = world * args.steps / e2e_s
    e2e_sync = world * args.steps / sync_s

    # ------------------------------------------------------------------ secondary workloads
    secondary = {}
    if not args.no_secondary:
        from fyusenet_b200 import multigpu

        def max_over_ranks(v):
            if world == 1:
                return float(v)
            t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        comm = None
        try:
            comm = multigpu.make_comm(ctx, rank, world)
        except Exception as exc:
            secondary["comm_error"] = repr(exc)
        jobs = [("resnet50_b512", lambda: secondary_resnet_sharded(ctx, comm, rank, world, local_rank, barrier, max_over_ranks)),
                ("resnet50_b1", lambda: secondary_resnet_b1(ctx, rank, world, local_rank, barrier, max_over_ranks)),
                ("stylenet9x9_4096_bands", lambda: secondary_stylenet_bands(ctx, comm, rank, world, local_rank, barrier, max_over_ranks))]
        for name, job in jobs:
            try:
                if comm is None:
                    raise RuntimeError("no communicator")
                secondary[name] = job()
            except Exception as exc:
                secondary[name] = {"error": repr(exc)}
                if world > 1:
                    break                                                  # ranks may be out of step after a failure: stop here
        if comm is not None:
            secondary["nccl_version"] = comm.info()["nccl_version"]
            comm.destroy()

    if rank == 0:
        hbm, tf_burst, tf_sust, which = peaks()
        alg = layer_algorithmic()
        # dominant kernel = the layer with the largest share of the device time
        conv_ms = {k: v for k, v in layer_ms.items() if k in alg}
        top = top_name if top_name in alg else max(conv_ms, key=conv_ms.get)
        a = alg[top]
        members = [top]
        if chained and top.startswith("res"):
            # the residual trunk runs as ONE kernel (fyn_conv_chain), timed on its first layer: its algorithmic bytes / flops are
            # those of all the layers it computes, each counted as a layer of its own (input + output + residual + weights)
            members = [k for k in alg if k.startswith("res")]
            a = {"bytes": sum(alg[k]["bytes"] for k in members), "flops": sum(alg[k]["flops"] for k in members)}
        t_s = (top_ms if top == top_name else conv_ms[top]) / 1e3
        ai = a["flops"] / a["bytes"]
        ridge = tf_sust * 1e12 / (hbm * 1e9)
        if ai > ridge:
            roof = {"bound": "tensor", "achieved": a["flops"] / t_s / 1e12, "peak": tf_sust, "unit": "TFLOP/s"}
        else:
            roof = {"bound": "hbm", "achieved": a["bytes"] / t_s / 1e9, "peak": hbm, "unit": "GB/s"}
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof["traffic"] = None      # dram bytes read + written per launch of that kernel, from the committed ncu capture
        roof["traffic_source"] = None
        for f in sorted((ROOT / "profiles").glob("r*_top_kernel_ncu.json"), reverse=True):
            cap = json.loads(f.read_text())
            if cap.get("layer") == top and (len(members) == 1) == ("chain" not in cap.get("kernel", "")):
                roof["traffic"] = cap["traffic_bytes_per_launch"]
                roof["traffic_source"] = f"profiles/{f.name} (ncu --set full, one launch)"
                break
        roof["algorithmic_bytes_per_launch"] = a["bytes"]
        roof["algorithmic_flops_per_launch"] = a["flops"]
        roof["kernel"] = top if len(members) == 1 else f"{members[0]} ... {members[-1]} ({len(members)} layers, one persistent kernel)"
        roof["peak_source"] = f"{which} ({'sustained' if roof['bound'] == 'tensor' else 'copy'} figure, kernel timed inside a long step)"
        roof["ms_per_launch"] = t_s * 1e3
        roof["ms_per_launch_all_layers_timed"] = conv_ms[top]
        if iso_ms is not None:
            roof["ms_per_launch_kernel_alone"] = iso_ms
            roof["frac_kernel_alone"] = (a["bytes"] / (iso_ms / 1e3) / 1e9) / hbm if roof["bound"] == "hbm" else (a["flops"] / (iso_ms / 1e3) / 1e12) / tf_burst
        total_layer_ms = sum(layer_ms.values())
        # whole-network roofline: sum_l max(F_l / P, B_l / BW)
        t_lb = sum(max(v["flops"] / (tf_sust * 1e12), v["bytes"] / (hbm * 1e9)) for v in alg.values())
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "StyleNet 9x9 (stylenet9x9 layout, synthetic He weights) 1524x1856 RGB frame, BASELINE configs[1]",
                       "storage": "fp16 activations (reference default), fp32 accumulate", "frames_per_step": 1,
                       "parallelism": f"replicas x{world}" if world > 1 else "single GPU",
                       "cpu_binding": (f"every rank bound to the NUMA node of its GPU (rank 0: node {numa[0]}, {numa[1]} CPUs)" if numa and numa[0] >= 0
                                       else "none (single NUMA node or single process)"),
                       "l2": "per-step working set 713 MB >> 126 MB L2, no explicit flush"},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": int(img.nbytes), "d2h_bytes_per_step": out_bytes,
                    "ms_per_step": 1e3 * e2e_s / args.steps, "finite": finite, "delivered": int(delivered),
                    "api": "StyleNet9x9 asynchronous(): upload/layers/download pipelined on 3 streams, 3 sequences in flight",
                    "sync_value": e2e_sync, "sync_ms_per_step": 1e3 * sync_s / args.steps,
                    "byte_io": e2e8},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "network_roofline": {"t_lower_bound_ms": 1e3 * t_lb, "frac": (1e3 * t_lb) / (ms / args.steps),
                                 "gflop_per_frame": FRAME_GFLOP, "mb_per_frame": FRAME_MB},
            "layers_ms": {k: round(v, 4) for k, v in layer_ms.items()},
            "layer_kernel_family": families,
            "chained_layers": int(chained),
            "layer_ms_sum": total_layer_ms,
            "sustained": sustained,
            "secondary": secondary,
        }
        if not args.no_cpu_baseline and world == 1:
            # bounded sample: full frames until ~12 s of CPU work have been spent
            import fyn_oracle as fo
            cores = fo.set_num_threads(os.cpu_count() or 1)
            _, dt1, _ = cpu_port_frames_per_s(weights, HEIGHT)
            reps = int(min(12, max(2, round(12.0 / max(dt1, 1e-3)))))
            fps, dt, sample = cpu_port_frames_per_s(weights, HEIGHT, repeats=reps)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                    "sample": sample + f", {dt * reps:.1f} s total (reference GL path not runnable here: no EGL/Mesa)"}
        print(json.dumps(line), flush=True)
    net.destroy()
    if world > 1:
        dist.destroy_process_group()


def _net_launches(net) -> int:
    """Kernel launches counted by the C-ABI library on the network's context."""
    import ctypes as C
    from fyusenet_b200 import capi, hostapi
    h = hostapi.lib().fynhost_net_context(net._h)
    n = C.c_uint64()
    capi.check(capi.lib().fyn_launch_count(C.c_void_p(h), C.byref(n), 0))
    return n.value


if __name__ == "__main__":
    main()
