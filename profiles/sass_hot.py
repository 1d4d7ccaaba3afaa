#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass` export: per-kernel instruction mix and the
hottest SASS instructions by executed count / stall samples. Usage: sass_hot.py file.csv [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = None
for i, r in enumerate(rows):
    if r and r[0] == "Address":
        hdr = r; data = rows[i + 1:]; break
ix = {n: k for k, n in enumerate(hdr)}
tot_inst = 0; tot_samp = 0; ops = collections.Counter(); samp_ops = collections.Counter(); wf = 0; wf_ideal = 0
recs = []
for r in data:
    if len(r) < len(hdr) or r[0] == "Address" or not r[0].startswith("0x") and not r[0].isdigit():
        if r and r[0] == "Kernel Name": break
        continue
    try:
        n = int(r[ix["Instructions Executed"]] or 0); s = int(r[ix["# Samples"]] or 0)
    except ValueError:
        continue
    src = r[ix["Source"]].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    ops[op] += n; samp_ops[op] += s; tot_inst += n; tot_samp += s
    try:
        wf += int(r[ix["L1 Wavefronts Shared"]] or 0); wf_ideal += int(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
    except (ValueError, KeyError):
        pass
    recs.append((n, s, src))
print(f"total warp-instructions {tot_inst:,}  samples {tot_samp:,}  smem wavefronts {wf:,} (ideal {wf_ideal:,})")
print("-- opcode mix (executed | stall samples)")
for op, n in ops.most_common(18):
    print(f"  {op:10s} {100*n/tot_inst:5.1f}%   {100*samp_ops[op]/max(tot_samp,1):5.1f}%")
print("-- hottest by stall samples")
for n, s, src in sorted(recs, key=lambda x: -x[1])[:top]:
    print(f"  {100*s/max(tot_samp,1):5.1f}%  {n:>12,}  {src[:100]}")
