#!/bin/bash
# A/B against a copy of the previous commit built under .ab_old/ (git archive HEAD ... | tar -x -C .ab_old; build there):
# interleaved rounds of tools/step_time.py, then the phase trace of the current build
for rnd in 1 2; do
  echo "== old (round $rnd)"; (cd .ab_old && python tools/step_time.py --reps 3 2>&1 | grep PDL | tail -1)
  echo "== new (round $rnd)"; python tools/step_time.py --reps 3 2>&1 | grep PDL | tail -1
done
python tools/phase_trace.py 2>&1 | head -32
