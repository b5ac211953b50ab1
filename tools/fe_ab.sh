#!/bin/bash
# A/B of builds of the library on the front end alone: REPS=2 tools/fe_ab.sh A.so B.so [C.so ...]  (alternates; the LAST one stays installed)
reps=${REPS:-2}
for i in $(seq 1 $reps); do
  for L in "$@"; do
    cp $L asr-model_b200/libasrb200.so
    python tools/fe_ab.py $L 2>/dev/null | tail -1
  done
done
