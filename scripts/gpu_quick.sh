#!/bin/bash
# Usage (via gpurun): scripts/gpu_quick.sh [tests] [demo] [man] [planes]  — short summaries only
summ='import json,sys
for line in sys.stdin:
    if line.startswith("{"):
        d=json.loads(line); print(sys.argv[1], "value %.3fM  step %.3f ms  kernel %.3f ms  e2e %.3fM  frac %.4f" % (d["value"]/1e6, d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"]/1e6, d["roofline"]["frac"]), "rows2host %.2fM" % (d.get("e2e_rows_to_host",{}).get("value",0)/1e6), {k: round(v,2) for k,v in d["e2e"].items() if k.endswith("_ms") and isinstance(v, float)})'
for what in "$@"; do
  case $what in
    tests) timeout 900 python -m pytest tests -q -m gpu --timeout 300 -x 2>&1 | tail -6 ;;
    demo) python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extra 2>&1 | python -c "$summ" demo ;;
    man) python bench.py --steps 30 --warmup 5 --basis manuscript --no-cpu-baseline --no-extra 2>&1 | python -c "$summ" manuscript ;;
    planes) UF3B_PLANES=1 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | python -c "$summ" demo-planes ;;
    *) env $what python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra ${BASIS:+--basis $BASIS} 2>&1 | python -c "$summ" "$what" ;;
  esac
done
