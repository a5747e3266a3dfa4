#!/bin/bash
# multi-GPU checks (run under gpurun --gpus N): replicas bench at C2 and the sharded C5 workload
set -u
N=${1:-2}
TAG=${2:-m}
OUT=gpurun_out
mkdir -p $OUT
for WL in C2 C5; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --workload $WL --steps 20 --warmup 3 > $OUT/${TAG}_bench_${WL}_${N}gpu.json 2> $OUT/${TAG}_bench_${WL}_${N}gpu.err
  echo "$WL x$N exit $?"; tail -c 1500 $OUT/${TAG}_bench_${WL}_${N}gpu.json; tail -5 $OUT/${TAG}_bench_${WL}_${N}gpu.err | cut -c1-300
done
