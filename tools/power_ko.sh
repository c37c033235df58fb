#!/bin/bash
# sustained clock / power of the token MID kernel for knock-out variants: tools/power_ko.sh base ko_STG ...
for v in "$@"; do
  if [ "$v" = base ]; then unset T2S_B200_LIB; else export T2S_B200_LIB=$PWD/t2ms_b200/lib/variants/libt2s_b200_$v.so; fi
  POWER_TAG=$v POWER_LEGS="token MID" python tools/power_by_kernel.py 3 2>&1 | grep "ms per launch"
done
