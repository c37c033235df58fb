#!/bin/bash
# A/B of library variants (tools/build_variant.py): ms per guided step (and optionally the attention kernel time), two
# interleaved rounds.   usage: tools/ab.sh variant1 variant2 ...   ("base" = the in-tree build); AB_ATTN=1 adds attn_time.py
for rnd in 1 2; do
  for v in "$@"; do
    if [ "$v" = base ]; then unset T2S_B200_LIB; else export T2S_B200_LIB=$PWD/t2ms_b200/lib/variants/libt2s_b200_$v.so; fi
    echo "== $v (round $rnd)"
    if [ -n "$AB_ATTN" ]; then python tools/attn_time.py 2>&1 | tail -2; fi
    python tools/step_time.py --reps 3 2>&1 | grep PDL | tail -1
  done
done
unset T2S_B200_LIB
