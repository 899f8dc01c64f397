#!/bin/bash
# A/B of k_poly1's two loops on the bench fleet (one B200): parity test first, then the headline with each.
python -m pytest tests/test_gpu_front.py -x -q -m gpu -k "poly1_static or poly_items" 2>&1 | tail -3
for v in 1 0 1 0; do
ATSC_POLY1_STATIC=$v timeout 300 python bench.py --no-configs --steps 10 < /dev/null 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.readline()); r=j['roofline']
print('static=$v', round(j['value']), 'ms/step', round(j['ms_per_step'],3), 'poly one-engine ms/288', r['kernel_ms_per_288_series_one_engine']['poly'], 'in-step', round(r['kernel_ms_per_step']['poly'],3))"
done
