#!/bin/bash
# First validation of a kernel change on the GPU box: memcheck of the smallest case, the parity
# suite (all failures, not just the first), smoke(), and stage timings with the A/B options.
# usage: bash tools/gpu_check.sh <tag> [nosan]
tag=${1:-chk}
out=gpurun_out
mkdir -p $out
if [ "$2" != "nosan" ]; then
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x \
  -k "test_forward_vs_reference_kernels and (tiny or ragged) or test_backward_vs_reference_kernels and tiny or test_view_batch_equals and shape0" \
  > $out/sanitizer_$tag.log 2>&1; echo "memcheck rc=$?"
tail -5 $out/sanitizer_$tag.log
fi
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 --tb=short > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"
tail -40 $out/pytest_gpu_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -3 $out/smoke_$tag.log
for wl in c2 c1 c4; do
  python tools/stage_probe.py --workload $wl >> $out/stages_$tag.log 2>&1
  python tools/stage_probe.py --workload $wl --single >> $out/stages_$tag.log 2>&1
done
python tools/stage_probe.py --workload c2 --opt bwd_pred=0 >> $out/stages_$tag.log 2>&1
python tools/stage_probe.py --workload c2 --opt pbwd_minb=3 >> $out/stages_$tag.log 2>&1
python tools/stage_probe.py --workload c2_init >> $out/stages_$tag.log 2>&1
python tools/stage_probe.py --workload c2_init --opt sort_cap=4096 >> $out/stages_$tag.log 2>&1
python tools/stage_probe.py --workload c5 --steps 8 >> $out/stages_$tag.log 2>&1
cat $out/stages_$tag.log
