#!/bin/bash
# ncu --set full capture of selected kernels. usage: scripts/gpu_ncu.sh <tag> <workload> <kernel-regex> [skip] [count]
set -u
TAG=$1; WL=$2; KRE=$3; SKIP=${4:-0}; CNT=${5:-4}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c $CNT -f -o $OUT/${TAG}_full_${WL} \
   python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full_${WL}.log 2>&1
tail -3 $OUT/${TAG}_ncu_full_${WL}.log | cut -c1-300
