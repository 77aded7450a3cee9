#!/bin/bash
# One GPU visit: parity tests, both bench arms, ncu launch list + full capture of one step.
# usage (from the repo root on the GPU box): bash tools/gpu_round.sh <tag> [quick]
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
if [ "$2" != "quick" ]; then
python -m pytest tests -m gpu -q -rs > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"
tail -3 $out/pytest_gpu_$tag.log
python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > $out/bench_ours_$tag.json 2> $out/bench_ours_$tag.err; echo "ours rc=$?"
cat $out/bench_ours_$tag.json
fi
# one step of ours = 8 kernels (both views in every launch): preprocess_fwd, scan_starts (+ ranges),
# scatter_entries, tile_sort, blend_fwd, blend_bwd, zero_scalars, preprocess_bwd (+ 2 memsets)
K='regex:blend|preprocess|scan_starts|tile_ranges|scatter_entries|tile_sort|zero_scalars'
python bench.py --resident-only --steps 2 --warmup 3 --no-clocks > $out/plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/launches_$tag.csv python bench.py --resident-only --steps 2 --warmup 3 --no-clocks > $out/ncu_l_$tag.log 2>&1
echo "launch list rc=$?"
ncu --set full --metrics smsp__inst_executed_op_global_red.sum,smsp__inst_executed_op_global_atom.sum \
    --clock-control none -k "$K" --launch-skip 24 --launch-count 8 \
    -o /tmp/prof_$tag -f python bench.py --resident-only --steps 2 --warmup 3 --no-clocks > $out/ncu_f_$tag.log 2>&1
echo "full capture rc=$?"
ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > $out/prof_${tag}_raw.csv 2>/dev/null
ls -la /tmp/prof_$tag.ncu-rep
# the two blend kernels once more with source correlation
ncu --set full --clock-control none --import-source on -k regex:blend --launch-skip 6 --launch-count 2 \
    -o $out/prof_${tag}_blend -f python bench.py --resident-only --steps 2 --warmup 3 --no-clocks > $out/ncu_b_$tag.log 2>&1
echo "blend capture rc=$?"
ls -la $out/prof_${tag}_blend.ncu-rep
