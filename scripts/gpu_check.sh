#!/bin/bash
# Usage (on the GPU box, via gpurun): scripts/gpu_check.sh <tag> [tests] [bench] [ncu] [ncufull]
# Writes everything under gpurun_out/<tag>_*.
tag=$1; shift
mkdir -p gpurun_out
for what in "$@"; do
  case $what in
    tests)
      timeout 900 python -m pytest tests -q -m gpu --timeout 300 -x 2>&1 | tail -15 | tee gpurun_out/${tag}_tests.log ;;
    bench)
      python bench.py --steps 50 --warmup 5 > gpurun_out/${tag}_bench_demo.json 2> gpurun_out/${tag}_bench_demo.err
      tail -2 gpurun_out/${tag}_bench_demo.err; cat gpurun_out/${tag}_bench_demo.json
      python bench.py --steps 30 --warmup 5 --basis manuscript --no-cpu-baseline > gpurun_out/${tag}_bench_manuscript.json 2> gpurun_out/${tag}_bench_man.err
      tail -2 gpurun_out/${tag}_bench_man.err; cat gpurun_out/${tag}_bench_manuscript.json ;;
    ncu)
      ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
        python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1 ;;
    ncufull)
      ncu --set full --clock-control none --import-source on -k regex:k_featurize -s 3 -c 1 -f -o gpurun_out/${tag}_featurize \
        python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncufull.log 2>&1
      tail -2 gpurun_out/${tag}_ncufull.log ;;
  esac
done
