#!/usr/bin/env python
"""Per-CUDA-source-line totals from an `ncu --page source --csv --print-source cuda,sass` export.
Usage: src_lines.py file.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
out = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r; ix = {n: k for k, n in enumerate(hdr)}; continue
    if hdr is None or len(r) < 8: continue
    if r[0] != "" and r[0].isdigit():
        try:
            out.append((int(r[0]), r[1].strip(), int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0)))
        except ValueError:
            pass
ti = sum(o[2] for o in out); ts = sum(o[3] for o in out)
print(f"total warp-instr {ti:,}  samples {ts:,}")
for ln, src, n, s in sorted(out, key=lambda o: -o[2])[:top]:
    print(f"{100*n/max(ti,1):5.1f}% inst {100*s/max(ts,1):5.1f}% stall  L{ln:<4d} {src[:110]}")
