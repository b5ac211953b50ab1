#!/bin/bash
# A/B of one environment switch on the headline step: tools/ab_env.sh VAR [reps]   (alternates VAR=1 / VAR=0, prints ms per step)
var=$1; reps=${2:-4}
for i in $(seq 1 $reps); do
  for v in 1 0; do
    env $var=$v python bench.py --extras 0 --no-cpu-baseline --steps 60 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$var=$v', round(d['ms_per_step'],4), d['clocks']['sm_mhz'])"
  done
done
