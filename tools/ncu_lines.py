#!/usr/bin/env python
"""Attribute the 'Instructions Executed' column of an ncu SASS source page to source file:line.

    ncu -i X.ncu-rep --page source --csv --print-source sass --kernel-name regex:K > sass.csv
    cuobjdump -xelf all libnmpm.so ; nvdisasm --print-line-info *.cubin > all.sass
    python tools/ncu_lines.py sass.csv all.sass <mangled-kernel-substring> [--top 40]

Joins by instruction offset (ncu addresses are absolute; offsets are relative to the first row).
"""
import argparse
import collections
import csv
import re

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("sass")
ap.add_argument("kernel")
ap.add_argument("--top", type=int, default=40)
a = ap.parse_args()

# offset -> (file, line) from nvdisasm
loc = {}
cur = None
infn = False
for ln in open(a.sass, errors="replace"):
    if ln.startswith(".text."):
        infn = a.kernel in ln
        cur = None
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        loc[int(m.group(1), 16)] = (cur, m.group(2).strip())

rows = list(csv.reader(open(a.csv)))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(rows[2][ia], 16)
per_line = collections.Counter()
samp_line = collections.Counter()
per_op = collections.Counter()
tot = 0
for r in rows[2:]:
    if not r or not r[0].startswith("0x"):
        break  # next kernel instance
    off = int(r[ia], 16) - base
    n = int(r[ii])
    tot += n
    where, ins = loc.get(off, (None, r[1]))
    per_line[where] += n
    samp_line[where] += int(r[isamp])
    op = ins.split()[0] if not ins.startswith("@") else ins.split()[1]
    per_op[op.split(".")[0]] += n
print(f"total warp instructions executed: {tot}")
stot = sum(samp_line.values())
print("--- by source line (instr share, stall-sample share)")
for where, n in per_line.most_common(a.top):
    print(f"{n / tot:7.3%} {samp_line[where] / max(1, stot):7.3%}  {where}")
print("--- by opcode")
for op, n in per_op.most_common(25):
    print(f"{n / tot:7.3%}  {op}")
print("--- by enclosing function (nmpm_math.cuh) / 20-line block (other files)")
import bisect
from pathlib import Path
math_src = Path(__file__).resolve().parents[1] / "nuclearmpm_b200" / "csrc" / "nmpm_math.cuh"
starts, names = [], []
for no, ln in enumerate(math_src.read_text().splitlines(), 1):
    m = re.match(r"^(?:NMPM_HD|template|#define NMPM_HESTENES_PAIR|#define NMPM_ROT)", ln)
    m2 = re.search(r"NMPM_HD\s+[\w<>:&\s\*]+?\s(\w+)\(", ln) or re.search(r"#define (NMPM_\w+)", ln)
    if m2:
        starts.append(no)
        names.append(m2.group(1))
def bucket(where):
    if where is None:
        return "unknown"
    f, l = where
    if f == "nmpm_math.cuh" and starts:
        k = bisect.bisect_right(starts, l) - 1
        return "math:" + (names[k] if k >= 0 else "?")
    return f"{f}:{l // 20 * 20}+"
agg, sagg = collections.Counter(), collections.Counter()
for where, n in per_line.items():
    agg[bucket(where)] += n
    sagg[bucket(where)] += samp_line[where]
for k, n in agg.most_common(28):
    print(f"{n / tot:7.3%} {sagg[k] / max(1, stot):7.3%}  {k}")
