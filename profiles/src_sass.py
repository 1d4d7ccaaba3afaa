#!/usr/bin/env python
"""SASS instructions (with executed counts) under given CUDA source lines of an
`ncu --page source --csv --print-source cuda,sass` export.  Usage: src_sass.py file.csv FILE_SUFFIX LINE [LINE ...]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
suffix = sys.argv[2]
want = {int(v) for v in sys.argv[3:]}
fpath, cur, ix = None, None, None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1]; continue
    if r[0] == "Line No": ix = {n: k for k, n in enumerate(r)}; continue
    if ix is None or len(r) < 8: continue
    if r[0].isdigit():
        cur = int(r[0])
        if fpath.endswith(suffix) and cur in want: print(f"--- L{cur}: {r[1].strip()[:120]}  [{r[ix['Instructions Executed']]} inst]")
        continue
    if fpath and fpath.endswith(suffix) and cur in want:
        print(f"      {r[3].strip():70s} {r[7]}")
