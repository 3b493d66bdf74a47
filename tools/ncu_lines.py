#!/usr/bin/env python
"""Summarise an ncu report by CUDA source line: python tools/ncu_lines.py <report.ncu-rep> [top]
(needs kernels compiled with -lineinfo and captured with --import-source on)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
lines = []
hdr = None
fname = ""
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        idx = {h: i for i, h in enumerate(hdr)}
        continue
    if hdr and r[0] not in ("", "Function Name") and len(r) == len(hdr):
        lines.append((fname, r))
samp = idx["# Samples"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[samp] or 0) for _, r in lines)
print(f"total samples {tot}")
for f, r in sorted(lines, key=lambda x: -int(x[1][samp] or 0))[:top]:
    s = int(r[samp] or 0)
    st = sorted(((h, int(r[idx[h]] or 0)) for h in stalls), key=lambda kv: -kv[1])[:2]
    print(f"{s:6d} {100 * s / max(tot, 1):5.1f}%  {f}:{r[0]:>4s}  {r[1].strip()[:90]:90s} {st[0][0]}={st[0][1]} {st[1][0]}={st[1][1]}")
