#!/bin/bash
# Usage (via gpurun): scripts/gpu_ncu.sh <tag> <kernel regex> [skip] [bench args...]  — one --set full capture
tag=$1; regex=$2; skip=${3:-3}; shift 3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o gpurun_out/${tag} \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra "$@" > gpurun_out/${tag}.log 2>&1
tail -2 gpurun_out/${tag}.log
