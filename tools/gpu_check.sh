#!/bin/bash
# First validation of a kernel change on the GPU box: the parity
# suite (all failures, not just the first), smoke(), and stage timings with the A/B options.
# usage: bash tools/gpu_check.sh <tag>   (compute-sanitizer is closed on this pool)
tag=${1:-chk}
out=gpurun_out
mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "test_blend_backward_variants_agree or test_tile_sort or test_equal_depths" > $out/ring_$tag.log 2>&1; echo "ring test rc=$?"; tail -3 $out/ring_$tag.log
timeout 1500 python -m pytest tests -m gpu -q -rs --maxfail=25 --tb=short > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"
tail -40 $out/pytest_gpu_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -3 $out/smoke_$tag.log
for wl in c2 c1 c4; do
  python tools/stage_probe.py --workload $wl >> $out/stages_$tag.log 2>&1
  python tools/stage_probe.py --workload $wl --single >> $out/stages_$tag.log 2>&1
done
python tools/stage_probe.py --workload c2_init >> $out/stages_$tag.log 2>&1
python tools/stage_probe.py --workload c5 >> $out/stages_$tag.log 2>&1
for wl in c2 c4 c2_init; do python tools/stage_probe.py --workload $wl --opt sort_adapt=0 >> $out/stages_$tag.log 2>&1; done

cat $out/stages_$tag.log
if [ "$2" == "bench" ]; then
python bench.py --steps 20 --warmup 5 > $out/bench_ours_$tag.json 2> $out/bench_ours_$tag.err; echo "bench ours rc=$?"; tail -3 $out/bench_ours_$tag.err
if [ "$3" == "ref" ]; then
python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err; echo "bench ref rc=$?"; tail -3 $out/bench_ref_$tag.err
fi
cat $out/bench_ours_$tag.json
fi
