#!/bin/bash
# token kernel times of library variants, two interleaved rounds: tools/ab_tok.sh base ko_LDS ...
for rnd in 1 2; do
  for v in "$@"; do
    if [ "$v" = base ]; then unset T2S_B200_LIB; else export T2S_B200_LIB=$PWD/t2ms_b200/lib/variants/libt2s_b200_$v.so; fi
    echo "$v (round $rnd): $(python tools/tok_time.py 2>&1 | tail -1)"
  done
done
