#!/bin/bash
# Multi-GPU visit: exchange probe, then both bench arms under torchrun.  usage: bash tools/multi_gpu_check.sh <N> <tag>
N=${1:-4}
tag=${2:-mg}
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > $out/topo_$tag.txt 2>&1; nproc >> $out/topo_$tag.txt
timeout 600 $TR --master-port 29511 tools/exchange_probe.py > $out/exchange_$tag.log 2>&1; echo "exchange rc=$?"; tail -25 $out/exchange_$tag.log
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > $out/bench_ours_$tag.json 2> $out/bench_ours_$tag.err; echo "ours rc=$?"; tail -5 $out/bench_ours_$tag.err
cat $out/bench_ours_$tag.json
timeout 900 $TR --master-port 29513 bench.py --gpus $N --impl reference --steps 20 --warmup 5 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err; echo "ref rc=$?"; tail -5 $out/bench_ref_$tag.err
cat $out/bench_ref_$tag.json
