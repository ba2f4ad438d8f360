#!/usr/bin/env python
"""Stall samples of an .ncu-rep by CUDA source line (needs -lineinfo and --import-source on).
Usage: python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top]"""
import collections, csv, io, os, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg = collections.Counter(); inst = collections.Counter(); why = collections.defaultdict(collections.Counter)
fname, hdr, tot = "", None, 0.0
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) < len(hdr) or r[0] == "" or not r[0].isdigit():
        continue
    if r[ix["# Samples"]] in ("-", ""):
        continue
    s = float(r[ix["# Samples"]] or 0)
    key = "%s:%s  %s" % (fname, r[0], r[1].strip()[:100])
    agg[key] += s; inst[key] += float(r[ix["Instructions Executed"]] or 0); tot += s
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            v = float(r[ix[h]] or 0)
            if v: why[key][h[6:]] += v
print("total samples", tot)
for k, v in agg.most_common(top):
    w = ", ".join("%s %.0f" % (a, b) for a, b in why[k].most_common(3))
    print("%5.1f%% inst %9.0f  %s   [%s]" % (100 * v / max(tot, 1), inst[k], k, w))
