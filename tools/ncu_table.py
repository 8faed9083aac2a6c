#!/usr/bin/env python
"""One line per kernel from an `ncu --page raw --csv` dump: time, DRAM bytes, pipe / issue utilisation, top stalls.
python tools/ncu_table.py raw.csv [--json out.json]"""
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k, d=0.0):
    try:
        return float(r[ix[k]])
    except Exception:
        return d


stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
out = []
print("%-58s %8s %8s %8s %6s %6s %6s %6s %5s %4s  %s" % ("kernel", "ms", "rd GB", "wr GB", "dram%", "l1dp%", "issue%", "fma%", "warps", "regs", "top stalls (warps per issue)"))
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("ssm::", "")
    short = re.sub(r"__nv_bfloat16", "bf16", short)
    ms = f(r, "gpu__time_duration.sum")
    unit = rows[1][ix["gpu__time_duration.sum"]]
    ms = ms / 1e3 if unit == "us" else (ms / 1e6 if unit == "ns" else ms)

    def gb(k):
        v, u = f(r, k), rows[1][ix[k]]
        return v * {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9, "Tbyte": 1e3}.get(u, 1.0)
    top = sorted(((f(r, s), s.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for s in stalls), reverse=True)[:3]
    rec = {"kernel": short, "ms": ms, "dram_read_gb": gb("dram__bytes_read.sum"), "dram_write_gb": gb("dram__bytes_write.sum"),
           "dram_pct": f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
           "l1_data_pipe_pct": f(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
           "issue_pct": f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "fma_pipe_pct": f(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
           "warps_active_pct": f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
           "registers": int(f(r, "launch__registers_per_thread")), "grid": int(f(r, "launch__grid_size")),
           "top_stalls": {n: round(v, 2) for v, n in top}}
    out.append(rec)
    print("%-58s %8.3f %8.3f %8.3f %6.1f %6.1f %6.1f %6.1f %5.0f %4d  %s" % (
        short[:58], ms, rec["dram_read_gb"], rec["dram_write_gb"], rec["dram_pct"], rec["l1_data_pipe_pct"], rec["issue_pct"],
        rec["fma_pipe_pct"], rec["warps_active_pct"], rec["registers"], ", ".join("%s %.1f" % (n, v) for v, n in top)))
if "--json" in sys.argv:
    json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
