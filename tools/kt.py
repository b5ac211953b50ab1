#!/usr/bin/env python
"""Print step time and the per-kernel table of bench.py JSON lines: python tools/kt.py file.json ..."""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['ms_per_step'], 3), round(d['value']), 'e2e', round(d['e2e']['value']),
              {k.replace('gemm_tc_', ''): round(v['ms_per_step'], 3) for k, v in d['kernels'].items()})
    except Exception as e:
        print(f, 'ERR', e)
