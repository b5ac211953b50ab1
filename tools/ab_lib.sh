#!/bin/bash
# A/B of two builds of the library on the headline step: tools/ab_lib.sh A.so B.so [reps]  (alternates, prints ms per step and
# the per-kernel event times; restores A at the end)
A=$1; B=$2; reps=${3:-3}
for i in $(seq 1 $reps); do
  for L in $A $B; do
    cp $L asr-model_b200/libasrb200.so
    python bench.py --extras 0 --no-cpu-baseline --steps 60 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['kernels']
print('$L', round(d['ms_per_step'],4), d['clocks']['sm_mhz'], {n[8:22]: round(v['ms_per_step'],3) for n,v in k.items() if n.startswith('gemm_tc_')})"
  done
done
cp $A asr-model_b200/libasrb200.so
