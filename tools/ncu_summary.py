#!/usr/bin/env python
"""Markdown summary of an `ncu --set full` report (one column per profiled launch).

    ncu -i X.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_summary.py raw.csv --title "cfg4 ..." --particles 16777216 [--algo p2g=108,g2p=152]
"""
import argparse
import csv

ROWS = [
    ("duration (us)", "gpu__time_duration.sum", "us"),
    ("DRAM read (MB)", "dram__bytes_read.sum", "MB"),
    ("DRAM write (MB)", "dram__bytes_write.sum", "MB"),
    ("DRAM throughput % of peak", "dram__throughput.avg.pct_of_peak_sustained_elapsed", None),
    ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed", None),
    ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active", None),
    ("FMA pipe %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", None),
    ("ALU pipe %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", None),
    ("XU (MUFU) pipe %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", None),
    ("LSU pipe %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", None),
    ("regs/thread", "launch__registers_per_thread", None),
    ("warps active % (achieved occupancy)", "sm__warps_active.avg.pct_of_peak_sustained_active", None),
    ("eligible warps / cycle / SMSP", "smsp__warps_eligible.avg.per_cycle_active", None),
    ("warp instructions", "smsp__inst_executed.sum", None),
    ("threads per warp instruction", "smsp__thread_inst_executed_per_inst_executed.ratio", None),
    ("L1 sector hit %", "l1tex__t_sector_hit_rate.pct", None),
    ("L2 sector hit %", "lts__t_sector_hit_rate.pct", None),
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
         "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}

ap = argparse.ArgumentParser()
ap.add_argument("raw")
ap.add_argument("--title", default="")
ap.add_argument("--particles", type=float, default=0)
ap.add_argument("--algo", default="p2g=108,g2p=152,g2p_p2g=260")
a = ap.parse_args()
algo = {k: float(v) for k, v in (kv.split("=") for kv in a.algo.split(","))}

rows = list(csv.reader(open(a.raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
names = [r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("nmpm::", "") for r in data]


def val(r, key, want):
    if key not in hdr:
        cand = [h for h in hdr if h.endswith(key)]
        if not cand:
            return None
        key = cand[0]
    i = hdr.index(key)
    try:
        v = float(r[i].replace(",", ""))
    except ValueError:
        return None
    u = units[i]
    if want == "us":
        return v * SCALE.get(u, 1.0)
    if want == "MB":
        return v * SCALE.get(u, 1.0) / 1e6
    return v


print(f"## {a.title}\n")
print("| metric | " + " | ".join(names) + " |")
print("|---|" + "---|" * len(names))
for label, key, want in ROWS:
    cells = []
    for r in data:
        v = val(r, key, want)
        cells.append("—" if v is None else (f"{v:,.0f}" if abs(v) >= 1000 else f"{v:.2f}"))
    print(f"| {label} | " + " | ".join(cells) + " |")
if a.particles:
    ab, ratio, wi = [], [], []
    for n, r in zip(names, data):
        k = "g2p_p2g" if "g2p_p2g" in n else "p2g" if "p2g" in n else "g2p" if "g2p" in n else None
        if k not in algo:
            k = None
        tr = (val(r, "dram__bytes_read.sum", "MB") or 0) + (val(r, "dram__bytes_write.sum", "MB") or 0)
        if k:
            b = algo[k] * a.particles / 1e6
            ab.append(f"{b:,.1f}")
            ratio.append(f"{tr / b:.2f}")
            wi.append(f"{val(r, 'smsp__inst_executed.sum', None) / (a.particles / 32):,.0f}")
        else:
            ab.append("—"), ratio.append("—"), wi.append("—")
    print("| algorithmic bytes (MB) | " + " | ".join(ab) + " |")
    print("| DRAM traffic / algorithmic | " + " | ".join(ratio) + " |")
    print("| warp instructions per 32 particles | " + " | ".join(wi) + " |")
# stall reasons
print("\nTop warp stall reasons (pc sampling, share of stalled samples):\n")
for n, r in zip(names, data):
    st = [(float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for i, h in enumerate(hdr)
          if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued") and r[i] not in ("", "n/a")]
    tot = sum(v for v, _ in st) or 1.0
    print(f"* `{n}`: " + ", ".join(f"{h} {100 * v / tot:.0f}%" for v, h in sorted(st, reverse=True)[:6]))
