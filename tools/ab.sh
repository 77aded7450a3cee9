#!/bin/bash
# A/B of kernel variants through env switches: bash tools/ab.sh "<ENV=.. ENV=..>" ...
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg python bench.py --resident-only --steps 60 --warmup 5 --no-clocks ${WL:+--workload $WL} 2>&1 | tail -1
done
