#!/usr/bin/env python
"""Key metrics per kernel from an `ncu -i report.ncu-rep --page raw --csv` export, plus the DRAM traffic JSON
that bench.py's roofline.traffic reads.  Usage: top_kernels.py raw.csv out.txt [traffic.json] ["header line"]"""
import csv, json, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
lines = [sys.argv[4] if len(sys.argv) > 4 else "ncu --set full --clock-control none --import-source on (config B, 1x B200)"]
traffic = {}
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    lines.append("----")
    lines.append(f"  {'Kernel Name':70s} {d['Kernel Name'][:100]} ")
    for k in KEYS:
        if k in d: lines.append(f"  {k:70s} {d[k]} {u[k]}")
    stalls = []
    for k in hdr:
        if k.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in k:
            try: stalls.append((float(d[k].replace(",", "")), k.split("stalled_")[1]))
            except ValueError: pass
    tot = sum(v for v, _ in stalls) or 1.0
    lines.append("  stall samples: " + "  ".join(f"{n}={100 * v / tot:.1f}%" for v, n in sorted(stalls, reverse=True)[:7]))
    name = d["Kernel Name"].replace("void ", "").replace("<unnamed>::", "").replace("unnamed>::", "").split("<")[0].split("(")[0].strip().split("::")[-1]
    rd = float(d["dram__bytes_read.sum"].replace(",", "")) * SCALE[u["dram__bytes_read.sum"]]
    wr = float(d["dram__bytes_write.sum"].replace(",", "")) * SCALE[u["dram__bytes_write.sum"]]
    traffic[name] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr}
open(sys.argv[2], "w").write("\n".join(lines) + "\n")
if len(sys.argv) > 3 and sys.argv[3]:
    json.dump({"source": f"{sys.argv[2]} (ncu --set full, config B, one launch each)", "kernels": traffic},
              open(sys.argv[3], "w"), indent=1)
print("\n".join(lines))
