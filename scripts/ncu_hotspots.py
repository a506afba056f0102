#!/usr/bin/env python
"""Per-source-line summary of an ncu report's source page.

usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv
       python scripts/ncu_hotspots.py src.csv [top]
Prints, per CUDA source line, the share of executed warp instructions and of stall samples,
with the dominant stall reasons; lines sorted by samples."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]


def num(x):
    try:
        return float(x.split("(")[0])
    except ValueError:
        return 0.0


top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
hdr = None
lines = {}
for r in csv.reader(open(path)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_inst = hdr.index("Instructions Executed")
        i_samp = hdr.index("# Samples")
        stall_cols = [(k, c) for k, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
        i_wf = hdr.index("L1 Wavefronts Shared")
        i_wfx = hdr.index("L1 Wavefronts Shared Excessive")
        continue
    if hdr is None or len(r) < len(hdr) - 2 or not r[0]:
        continue
    try:
        key = (cur_file, int(r[0]))
    except ValueError:
        continue
    d = lines.setdefault(key, {"src": r[1].strip(), "inst": 0.0, "samp": 0.0, "wf": 0.0, "wfx": 0.0,
                               "stalls": defaultdict(float)})
    d["inst"] += num(r[i_inst])
    d["samp"] += num(r[i_samp])
    d["wf"] += num(r[i_wf])
    d["wfx"] += num(r[i_wfx])
    for k, c in stall_cols:
        if k < len(r):
            d["stalls"][c[6:]] += num(r[k])
ti = sum(d["inst"] for d in lines.values()) or 1.0
ts = sum(d["samp"] for d in lines.values()) or 1.0
tw = sum(d["wf"] for d in lines.values())
print(f"total warp instructions {ti:.0f}, samples {ts:.0f}, shared wavefronts {tw:.0f} "
      f"(excessive {sum(d['wfx'] for d in lines.values()):.0f})")
tot_st = defaultdict(float)
for d in lines.values():
    for k, v in d["stalls"].items():
        tot_st[k] += v
print("stall samples:", ", ".join(f"{k} {100 * v / ts:.1f}%" for k, v in sorted(tot_st.items(), key=lambda x: -x[1])[:8]))
for key, d in sorted(lines.items(), key=lambda kv: -kv[1]["samp"])[:top]:
    st = ", ".join(f"{k} {100 * v / max(d['samp'], 1):.0f}%" for k, v in sorted(d["stalls"].items(), key=lambda x: -x[1])[:2])
    print(f"{key[0]}:{key[1]:<5d} inst {100 * d['inst'] / ti:5.1f}%  samp {100 * d['samp'] / ts:5.1f}%  [{st}]  {d['src'][:90]}")

# per-file totals
by_file = defaultdict(lambda: [0.0, 0.0])
for (fn, ln), d in lines.items():
    by_file[fn][0] += d["inst"]
    by_file[fn][1] += d["samp"]
print("per file:", ", ".join(f"{fn} inst {100 * v[0] / ti:.1f}% samp {100 * v[1] / ts:.1f}%" for fn, v in sorted(by_file.items(), key=lambda x: -x[1][0])))
if len(sys.argv) > 3:      # ranges "file:lo-hi,..."
    for spec in sys.argv[3].split(","):
        fn, rng = spec.split(":")
        lo, hi = map(int, rng.split("-"))
        i = sum(d["inst"] for (f_, l_), d in lines.items() if f_ == fn and lo <= l_ <= hi)
        s_ = sum(d["samp"] for (f_, l_), d in lines.items() if f_ == fn and lo <= l_ <= hi)
        print(f"{spec}: inst {100 * i / ti:.1f}% samp {100 * s_ / ts:.1f}%")
if tw:
    print("shared-memory wavefronts by line:")
    for key, d in sorted(lines.items(), key=lambda kv: -kv[1]["wf"])[:max(12, top // 3)]:
        print(f"{key[0]}:{key[1]:<5d} wavefronts {100 * d['wf'] / tw:5.1f}% (excess {100 * d['wfx'] / max(d['wf'], 1):3.0f}%)  {d['src'][:90]}")
