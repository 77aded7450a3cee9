#!/bin/bash
# ncu capture of the UNMODIFIED reference kernels (oracle/_ref) on one c2 step — SURVEY §8(d):
# "capture the oracle kernels first".  One step = 2 views x (preprocessCUDA, DeviceScan x2,
# duplicateWithKeys, DeviceRadixSort histogram + exclusive sum + 6 onesweep passes,
# identifyTileRanges, renderCUDA fwd, renderCUDA bwd, computeCov2DCUDA, preprocessCUDA bwd).
# usage (repo root, on the GPU box): bash tools/ref_profile.sh <tag>
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
K='regex:CUDA|duplicateWithKeys|identifyTileRanges|DeviceRadixSort|DeviceScan|checkFrustum'
python bench.py --impl reference --resident-only --steps 2 --warmup 3 --no-clocks > $out/ref_plain_$tag.log 2>&1 && \
ncu --set full --metrics smsp__inst_executed_op_global_red.sum,smsp__inst_executed_op_global_atom.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum \
    --clock-control none -k "$K" --launch-skip 108 --launch-count 40 \
    -o /tmp/prof_ref_$tag -f python bench.py --impl reference --resident-only --steps 2 --warmup 3 --no-clocks > $out/ncu_ref_$tag.log 2>&1
echo "reference capture rc=$?"
ncu -i /tmp/prof_ref_$tag.ncu-rep --page raw --csv > $out/prof_ref_${tag}_raw.csv 2>/dev/null
ls -la /tmp/prof_ref_$tag.ncu-rep; wc -l $out/prof_ref_${tag}_raw.csv
