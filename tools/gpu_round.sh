#!/bin/bash
# One GPU visit: parity tests, both bench arms, ncu launch list + full capture of one step.
# usage (from the repo root on the GPU box): bash tools/gpu_round.sh <tag> [quick]
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
if [ "$2" != "quick" ]; then
python -m pytest tests -m gpu -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"
tail -3 $out/pytest_gpu_$tag.log
python bench.py --impl reference > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err; echo "ref rc=$?"
python bench.py > $out/bench_ours_$tag.json 2> $out/bench_ours_$tag.err; echo "ours rc=$?"
cat $out/bench_ours_$tag.json
fi
python bench.py --resident-only --steps 2 --warmup 3 > $out/plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $out/launches_$tag.csv python bench.py --resident-only --steps 2 --warmup 3 > $out/ncu_l_$tag.log 2>&1
echo "launch list rc=$?"
# one whole step (26 launches of ours: 2 views x 13, views back to back so the order is fixed),
# all sections, no source (keeps the report small)
ncu --set full --clock-control none -k "regex:blend|preprocess|duplicate|identify|rs_|zero_scalars" --launch-skip 84 --launch-count 28 \
    -o /tmp/prof_$tag -f python bench.py --resident-only --sequential-views --steps 2 --warmup 3 > $out/ncu_f_$tag.log 2>&1
echo "full capture rc=$?"
ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > $out/prof_${tag}_raw.csv 2>/dev/null
ls -la /tmp/prof_$tag.ncu-rep
# the two blend kernels once more with source correlation
ncu --set full --clock-control none --import-source on -k regex:blend --launch-skip 12 --launch-count 2 \
    -o $out/prof_${tag}_blend -f python bench.py --resident-only --sequential-views --steps 2 --warmup 3 > $out/ncu_b_$tag.log 2>&1
echo "blend capture rc=$?"
ls -la $out/prof_${tag}_blend.ncu-rep
