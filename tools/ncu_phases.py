#!/usr/bin/env python3
"""Share of executed instructions / stall samples per phase of hca_encode_kernel from an ncu report (needs -lineinfo).
    python tools/ncu_phases.py report.ncu-rep [frames]"""
import collections, csv, io, os, subprocess, sys
rep = sys.argv[1]; frames = int(sys.argv[2]) if len(sys.argv) > 2 else 770048
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None; cur = None; files = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1]; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] != "": files.setdefault(cur, []).append(r)
ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
tot = sum(int(r[ie]) for rs in files.values() for r in rs); tots = sum(int(r[isamp]) for rs in files.values() for r in rs)
print("warp instructions", tot, "per frame", round(tot / frames))
src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pycricodecs_b200/csrc/hca_enc_kernels.cu")).read().split("\n")
marks = [("helpers", "namespace {"), ("band_cost (uncounted)", "int band_cost"), ("header_lengths", "void header_lengths"),
         ("prologue", "hca_encode_kernel(HcaEncodeArgs a)"), ("MDCT", "---- MDCT"), ("intensity", "---- intensity stereo"), ("scalefactors", "---- scalefactors"),
         ("hfr averages", "---- HFR group averages"), ("scaled spectra + ranks", "---- scaled spectra"), ("hfr scales", "---- HFR scales"),
         ("bit allocation", "---- bit allocation"), ("final resolutions", "---- final resolutions"), ("header pack", "---- pack (hca.cpp"),
         ("quantise", "Spectra in two phases"), ("pack", "rows in bitstream order"), ("crc + store", "---- CRC16 over")]
anchors = []
for name, text in marks:
    hit = [i + 1 for i, l in enumerate(src) if text in l]
    if hit: anchors.append((hit[0], name))
anchors.sort()
acc = collections.Counter(); accs = collections.Counter()
for fn, rs in files.items():
    for r in rs:
        n = int(r[ie]); s = int(r[isamp]); ln = int(r[0])
        if fn.endswith("hca_enc_kernels.cu"):
            name = ([a for l, a in anchors if l <= ln] or ["top"])[-1]
        else:
            name = fn.split("/")[-1]
        acc[name] += n; accs[name] += s
for k, v in acc.most_common(30): print(f"{v / tot * 100:5.1f}% inst ({v / frames:7.0f}/frame) {accs[k] / tots * 100:5.1f}% stall samples  {k}")
