#!/bin/bash
# guided-step time of library variants at several batch sizes: tools/ab_small.sh "64 1024" base variant ...
sizes=$1; shift
for rnd in 1 2; do
  for v in "$@"; do
    if [ "$v" = base ]; then unset T2S_B200_LIB; else export T2S_B200_LIB=$PWD/t2ms_b200/lib/variants/libt2s_b200_$v.so; fi
    for b in $sizes; do echo "$v (round $rnd) batch $b: $(python tools/step_time.py --batch $b --steps 50 --reps 3 2>&1 | grep PDL | tail -1)"; done
  done
done
