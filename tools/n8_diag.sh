#!/bin/bash
# per-rank step times at N ranks, fused NVLS vs NCCL exchange.  usage: bash tools/n8_diag.sh <N>
N=${1:-8}
out=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for mode in auto nccl; do
  if [ $mode == nccl ]; then export GFT_BENCH_EXCHANGE=nccl; fi
  GFT_BENCH_TRACE_RANKS=1 timeout 400 $TR --master-port 2955$N bench.py --gpus $N --steps 20 --warmup 5 --blocks none --resident-only > $out/diag_$mode.json 2> $out/diag_$mode.err
  grep "^step" $out/diag_$mode.err
  tail -1 $out/diag_$mode.json | cut -c1-600
done
