#!/bin/bash
# Round-1 binning chain, own onesweep radix sort vs cub::DeviceRadixSort (GFT_SORT=cub), at the four
# instance counts of the workloads c1 / c2 / c4 / c5 (R ~ 8e4 / 1.2e6 / 9e6 / 2.2e7 per view).
out=gpurun_out
mkdir -p $out
for wl in c1 c2 c4 c5; do
  for s in own cub; do
    echo "== $wl GFT_SORT=$s"
    GFT_SORT=$s python bench.py --resident-only --sequential-views --workload $wl --steps 10 --warmup 3 --no-clocks 2>&1 | tail -1
  done
done > $out/sort_ab_old.log 2>&1
cat $out/sort_ab_old.log
