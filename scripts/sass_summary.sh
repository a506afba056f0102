#!/bin/bash
# Usage: scripts/sass_summary.sh > profiles/r02_sass_summary.txt   (runs here: cuobjdump needs no GPU)
# Per kernel of libuf3b.so: counts of the SASS mnemonics that show what the kernel is made of —
# UBLKCP (TMA bulk copies), SYNCS (mbarrier), DMMA (FP64 tensor cores), DFMA/DADD/DMUL (FP64 pipe),
# LDS/STS (shared memory), LDG/STG, RED/ATOMG (global atomics), SHFL, BAR.
lib=${1:-uf3_b200/lib/libuf3b.so}
echo "# SASS mnemonic counts per kernel of $lib ($(date -u +%F), cuobjdump -sass, sm_100a)"
printf "%-60s %6s %6s %6s %6s %6s %6s %6s %6s %6s %6s %6s %6s %6s\n" kernel UBLKCP SYNCS DMMA DFMA DADD DMUL LDS STS LDG STG RED+AT SHFL BAR
cuobjdump -sass "$lib" | awk '
function flush() { if (name != "") printf "%-60s %6d %6d %6d %6d %6d %6d %6d %6d %6d %6d %6d %6d %6d\n", substr(name,1,60), c["UBLKCP"], c["SYNCS"], c["DMMA"], c["DFMA"], c["DADD"], c["DMUL"], c["LDS"], c["STS"], c["LDG"], c["STG"], c["RED"]+c["REDG"]+c["ATOMG"]+c["ATOM"]+c["ATOMS"], c["SHFL"], c["BAR"]; delete c }
/Function :/ { flush(); name=$3; cmd="c++filt -p " name; cmd | getline name; close(cmd) }
/^ +\/\*[0-9a-f]+\*\// { op=$2; sub(/^@!?U?P[0-9T]+ /, "", op); if ($2 ~ /^@/) op=$3; split(op, parts, "."); c[parts[1]]++ }
END { flush() }'
