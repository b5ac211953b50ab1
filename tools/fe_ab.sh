#!/bin/bash
# A/B of two builds of the library on the front end alone: tools/fe_ab.sh A.so B.so [reps]  (alternates; restores B at the end)
A=$1; B=$2; reps=${3:-2}
for i in $(seq 1 $reps); do
  for L in $A $B; do
    cp $L asr-model_b200/libasrb200.so
    python tools/fe_ab.py $L 2>/dev/null | tail -1
  done
done
cp $B asr-model_b200/libasrb200.so
