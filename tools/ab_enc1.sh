#!/bin/bash
# A/B of builds of the library on the enc = 1 and enc = 0 steps: REPS=2 tools/ab_enc1.sh A.so B.so ...  (alternates; the LAST one stays installed)
reps=${REPS:-2}
for i in $(seq 1 $reps); do
  for L in "$@"; do
    cp $L asr-model_b200/libasrb200.so
    python tools/ab_enc1.py $L 2>/dev/null | tail -1
  done
done
