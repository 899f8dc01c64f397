#!/bin/bash
# the headline with the current build, a few times (pair with tools/p1_ab.sh numbers)
for v in 1 2 3; do
timeout 300 python bench.py --no-configs --steps 10 < /dev/null 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.readline()); r=j['roofline']
print('run $v', round(j['value']), 'ms/step', round(j['ms_per_step'],3), 'poly one-engine ms/288', r['kernel_ms_per_288_series_one_engine']['poly'], 'in-step', round(r['kernel_ms_per_step']['poly'],3))"
done
