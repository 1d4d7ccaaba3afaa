#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel. Usage: launch_summary.py file.csv"""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
    agg.setdefault(row["Kernel Name"].split("(")[0][:70], []).append(v)
LOAD = ("k_gm_build", "k_gm_masks", "k_gm_offsets", "k_gm_values", "k_validate_csr", "k_validate_unique_cols")   # once per dataset (matrix load), not part of an analysis step
is_load = lambda k: any(n in k for n in LOAD)
tot = sum(sum(v) for k, v in agg.items() if not is_load(k))
print(f"{'us total':>12} {'n':>5} {'share':>7}  kernel   (cold-cache, serialised: compare shares; shares are of the analysis steps)")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    if is_load(k):
        print(f"{sum(v):12.1f} {len(v):5d} {'load':>7}  {k}   (once per dataset)")
    else:
        print(f"{sum(v):12.1f} {len(v):5d} {100*sum(v)/tot:6.1f}%  {k}")
