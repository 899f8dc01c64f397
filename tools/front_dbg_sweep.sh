#!/bin/bash
# timing experiments: which part of k_front costs what (results are WRONG with any bit set)
for d in 0 1 2 4 6 7 3 5; do
  echo "== ATSC_FRONT_DBG=$d (1: no stats, 2: no fold, 4: no polynomial trips)"
  ATSC_FRONT_DBG=$d timeout 100 python tools/prof_front.py ${1:-72} 3 2>&1 < /dev/null | tail -1 | sed 's/.*front/front/'
done
