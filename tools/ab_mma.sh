set -x
python -m pytest tests/test_gpu_dit.py tests/test_gpu_sampling.py -x -q -m gpu 2>&1 | tail -5
for rnd in 1 2; do
  for v in base mma1; do
    if [ "$v" = base ]; then unset T2S_B200_LIB; else export T2S_B200_LIB=$PWD/t2ms_b200/lib/variants/libt2s_b200_$v.so; fi
    echo "== $v (round $rnd)"
    python tools/step_time.py --reps 3 2>&1 | grep PDL | tail -1
  done
done
unset T2S_B200_LIB
python tools/phase_trace.py 2>&1 | head -40
