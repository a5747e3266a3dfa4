#!/bin/bash
# multi-GPU checks (run under gpurun --gpus N): the N-rank sharded parity tests on real peers (both leaf paths), then
# bench lines.   usage: scripts/gpu_multi.sh <N> <tag> [workloads...]     (default workloads: C4 C5)
# A workload suffixed with :rows (e.g. C5:rows) is run with MVIN_B200_XCHG=0: raw-row peer gathers instead of the
# owner-side partial reduction.
set -u
N=${1:-2}
TAG=${2:-m}
shift; shift || true
WLS=${@:-C4 C5}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/${TAG}_smi_${N}gpu.csv 2>&1
timeout 900 python -m pytest tests -m gpu -q -rs -k "multi_rank or virtual_entity" > $OUT/${TAG}_pytest_multi_${N}gpu.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest_multi_${N}gpu.log
tail -8 $OUT/${TAG}_pytest_multi_${N}gpu.log | cut -c1-300
for WLX in $WLS; do
  WL=${WLX%%:*}
  X=1; SUF=""
  if [ "$WLX" != "$WL" ]; then X=0; SUF="_rows"; fi
  MVIN_B200_XCHG=$X timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --workload $WL --steps 20 --warmup 5 --quick > $OUT/${TAG}_bench_${WL}${SUF}_${N}gpu.json 2> $OUT/${TAG}_bench_${WL}${SUF}_${N}gpu.err
  echo "$WLX x$N exit $?"; tail -c 2500 $OUT/${TAG}_bench_${WL}${SUF}_${N}gpu.json; tail -5 $OUT/${TAG}_bench_${WL}${SUF}_${N}gpu.err | cut -c1-300
done
