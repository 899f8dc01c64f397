#!/bin/bash
# headline step vs number of engines / wave size (1152-series step, device resident)
for e in 4 6 8; do for w in 36 72 144; do
  r=$(ATSC_ENGINES=$e ATSC_WAVE_MI=$w timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu --no-configs 2>/dev/null < /dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), round(d['decompress']['value']))")
  echo "engines=$e wave_mi=$w -> $r"
done; done
