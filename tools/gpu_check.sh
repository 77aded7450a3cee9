#!/bin/bash
# First validation of a kernel change on the GPU box: the parity
# suite (all failures, not just the first), smoke(), and stage timings with the A/B options.
# usage: bash tools/gpu_check.sh <tag>   (compute-sanitizer is closed on this pool)
tag=${1:-chk}
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 --tb=short > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"
tail -40 $out/pytest_gpu_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -3 $out/smoke_$tag.log
for wl in c2 c1 c4; do
  python tools/stage_probe.py --workload $wl >> $out/stages_$tag.log 2>&1
  python tools/stage_probe.py --workload $wl --single >> $out/stages_$tag.log 2>&1
done
python tools/stage_probe.py --workload c2 --opt tile_order=0 >> $out/stages_$tag.log 2>&1
python tools/stage_probe.py --workload c4 --opt tile_order=0 >> $out/stages_$tag.log 2>&1
python tools/stage_probe.py --workload c2 --opt sort_radix=0 >> $out/stages_$tag.log 2>&1
python tools/stage_probe.py --workload c2 --opt sort_match=1 >> $out/stages_$tag.log 2>&1
python tools/stage_probe.py --workload c2 --opt sort_cap=2048 >> $out/stages_$tag.log 2>&1
python tools/stage_probe.py --workload c2_init >> $out/stages_$tag.log 2>&1
python tools/stage_probe.py --workload c2_init --opt sort_radix=0 >> $out/stages_$tag.log 2>&1
python tools/stage_probe.py --workload c5 --steps 8 >> $out/stages_$tag.log 2>&1
cat $out/stages_$tag.log
if [ "$2" == "bench" ]; then
python bench.py --steps 20 --warmup 5 > $out/bench_ours_$tag.json 2> $out/bench_ours_$tag.err; echo "bench ours rc=$?"; tail -3 $out/bench_ours_$tag.err
if [ "$3" == "ref" ]; then
python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err; echo "bench ref rc=$?"; tail -3 $out/bench_ref_$tag.err
fi
cat $out/bench_ours_$tag.json
fi
