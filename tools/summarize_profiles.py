#!/usr/bin/env python
"""Turns the ncu artefacts of a round into the small tracked summaries under profiles/:

  python tools/summarize_profiles.py <launches.csv> <top_kernel.ncu-rep> <bench.json> <round tag, e.g. r01>

* <tag>_launches.csv            copy of the `ncu --metrics gpu__time_duration.sum --clock-control none` launch list
* <tag>_launches_summary.md     kernel shares of one StyleNet forward (ncu, cold cache, serialised) next to the shares
                                bench.py measured live with CUDA events
* <tag>_top_kernel_ncu.json     DRAM traffic, durations and pipe utilisation of the dominant kernel (`ncu --set full`)
"""
import csv
import json
import shutil
import subprocess
import sys
from pathlib import Path

launches, rep, bench, tag = sys.argv[1:5]
layer_override = sys.argv[5] if len(sys.argv) > 5 else None      # capture of a layer other than the bench's dominant one
out = Path(__file__).resolve().parent.parent / "profiles"
shutil.copy(launches, out / f"{tag}_launches.csv")

rows = list(csv.reader(open(launches)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
idx = {n: i for i, n in enumerate(rows[hi])}
data = [(r[idx["Kernel Name"]], float(r[idx["Metric Value"]]) / 1e3, r[idx["Grid Size"]]) for r in rows[hi + 1:] if len(r) > idx["Metric Value"]]
b = json.load(open(bench))
# one forward = the launches between two occurrences of the first kernel (conv1); layers that launch nothing (a sigmoid
# fused into its producer) are dropped from the bench's layer list by their ~0 event time
period = next(i for i in range(1, len(data)) if data[i][0] == data[0][0] and data[i][2] == data[0][2])
names = [k for k, v in b["layers_ms"].items() if v > 0.005]
n = period
assert len(names) == n, f"{len(names)} timed layers vs {n} launches per forward"
fw = [data[i * n:(i + 1) * n] for i in range(len(data) // n)]
use = [g for g in fw[1:4] if g[0][0] == data[0][0]]       # skip the very first forward (cold instruction caches)
avg = [sum(g[i][1] for g in use) / len(use) for i in range(n)]
tot_ncu, tot_ev = sum(avg), sum(b["layers_ms"][k] for k in names) * 1e3
lines = [f"# {tag}: kernel launches of one StyleNet-9x9 1524x1856 forward", "",
         "`ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py --steps 2 --warmup 3` (cold cache, serialised),",
         f"mean of {len(use)} forwards, beside the per-layer CUDA-event times of the bench run (`{Path(bench).name}`; an event pair per layer adds\n~5 us and breaks the dependent-launch overlap, which is why the frame time below is smaller than the event sum).", "",
         "| layer | kernel | grid | ncu us | ncu share | event us | event share |", "|---|---|---|---:|---:|---:|---:|"]
for i, nm in enumerate(names):
    k = use[0][i][0].replace("void <unnamed>::", "").replace("(<unnamed>::TcArgs)", "")
    ev = b["layers_ms"][nm] * 1e3
    lines.append(f"| {nm} | `{k}` | {use[0][i][2]} | {avg[i]:.1f} | {100 * avg[i] / tot_ncu:.1f}% | {ev:.1f} | {100 * ev / tot_ev:.1f}% |")
lines += [f"| **sum** | | | **{tot_ncu:.1f}** | | **{tot_ev:.1f}** | |", "",
          f"Frame time of the bench run: {b['ms_per_step'] * 1e3:.1f} us ({b['value']:.0f} frames/s); {len(data)} launches captured in total."]
(out / f"{tag}_launches_summary.md").write_text("\n".join(lines) + "\n")

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
h, u, v = r[0], r[1], r[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum"]
m = {}
for i, name in enumerate(h):
    if name in want:
        m[name] = {"value": float(v[i].replace(",", "")), "unit": u[i]}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
traffic = sum(m[k]["value"] * scale[m[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
kname = v[h.index("Kernel Name")] if "Kernel Name" in h else ""
summary = {"kernel": kname, "layer": layer_override or b["roofline"]["kernel"], "report": Path(rep).name, "traffic_bytes_per_launch": traffic, "metrics": m,
           "note": "one launch, ncu --set full --clock-control none; DRAM writes are below the algorithmic output bytes because the 126 MB L2 "
                   "absorbs most of the output (written back after the kernel)"}
name = f"{tag}_top_kernel_ncu.json" if not layer_override else f"{tag}_{layer_override}_kernel_ncu.json"
(out / name).write_text(json.dumps(summary, indent=1) + "\n")
print((out / f"{tag}_launches_summary.md").read_text())
print(json.dumps(summary)[:600])
