#!/bin/bash
# Usage (via gpurun): scripts/gpu_profile_round.sh <tag>  — one `ncu --set full` capture per kernel of
# the round (caches NOT flushed between kernels: the row kernels feed each other through L2).
tag=$1
mkdir -p gpurun_out
cap() {  # name regex skip args...
  name=$1; regex=$2; skip=$3; shift 3
  ncu --set full --clock-control none --cache-control none --import-source on -k regex:$regex -s $skip -c 1 -f \
      -o gpurun_out/${tag}_${name} "$@" > gpurun_out/${tag}_${name}.log 2>&1
  tail -1 gpurun_out/${tag}_${name}.log
}
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra"
export UF3B_BENCH_IN_FLIGHT=1
cap centre_legs k_centre_legs 3 $B
cap rows_nbr k_rows_nbr 3 $B
cap rows_ctr k_rows_ctr 3 $B
cap neighbors k_neighbors 3 $B
cap gram_narrow k_gram_narrow 3 $B
cap coop k_featurize_coop 3 $B --basis manuscript
cap gram k_gram\$ 3 python bench.py --workload fit --basis manuscript --steps 3 --warmup 2
cap multi2 k_rows_multi2 1 python scripts/binary_step.py --steps 1
cap energy_forces k_energy_forces 8 python scripts/md_step.py
