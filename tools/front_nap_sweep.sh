#!/bin/bash
for d in 0 32 64 128 400; do
  echo "== ATSC_FRONT_NAP=$d"
  ATSC_FRONT_NAP=$d timeout 60 python tools/prof_front.py ${1:-72} 3 2>&1 < /dev/null | tail -1 | sed 's/.*front/front/'
done
