#!/usr/bin/env python3
"""Per-source-line share of executed instructions and stall samples from an ncu report (needs -lineinfo).
    python tools/ncu_lines.py report.ncu-rep [min_percent]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None; files = {}
cur = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1]; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] != "": files.setdefault(cur, []).append(r)
ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); ith = hdr.index("Avg. Threads Executed")
tot = sum(int(r[ie]) for rs in files.values() for r in rs); tots = sum(int(r[isamp]) for rs in files.values() for r in rs)
print("total warp instructions", tot, "samples", tots)
for fn, rs in files.items():
    print("==", fn)
    for r in rs:
        p = int(r[ie]) / tot * 100; ps = int(r[isamp]) / max(tots, 1) * 100
        if p >= thr or ps >= thr: print(f"{r[0]:>5} inst {p:5.1f}%  stall-samples {ps:5.1f}%  thr {r[ith]:>3}  {r[1][:110]}")
