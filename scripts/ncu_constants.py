#!/usr/bin/env python
"""DRAM bytes and executed FP64 flops per launch from `ncu --set full` reports.

usage: python scripts/ncu_constants.py name=report.ncu-rep ...   (prints one JSON object per report)
flops = 2 x fused + 1 x non-fused FP64 thread instructions (predicated-off lanes excluded),
traffic = dram__bytes_read.sum + dram__bytes_write.sum."""
import csv
import json
import subprocess
import sys


def metrics(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def num(m, key):
    v, u = m[key]
    x = float(v.replace(",", ""))
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e-6, "msecond": 1e-3,
             "nsecond": 1e-9, "second": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(u, 1.0)
    return x * scale


for arg in sys.argv[1:]:
    name, report = arg.split("=", 1)
    m = metrics(report)
    dram = num(m, "dram__bytes_read.sum") + num(m, "dram__bytes_write.sum")
    cycles = num(m, "smsp__cycles_elapsed.avg") if "smsp__cycles_elapsed.avg" in m else num(m, "sm__cycles_elapsed.avg")
    fused = num(m, "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed") * cycles
    add = num(m, "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed") * cycles
    mul = num(m, "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed") * cycles
    print(json.dumps({"name": name, "kernel": m["Kernel Name"][0], "duration_us": num(m, "gpu__time_duration.sum") * 1e6,
                      "dram_bytes": dram, "fp64_flop": 2 * fused + add + mul,
                      "fp64_thread_inst": {"dfma": fused, "dadd": add, "dmul": mul},
                      "l1tex_lsu_wavefronts_pct": float(m["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"][0]),
                      "fp64_pipe_pct": float(m["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"][0])}))
