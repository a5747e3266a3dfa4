#!/bin/bash
# One gpurun call: parity tests, bench lines, ncu launch list and --set full captures (B200_PROFILING.md recipe).
# usage: scripts/gpu_capture.sh <tag> [workloads...]
set -u
TAG=${1:-r01}
shift || true
WLS=${@:-C2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.csv 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
for WL in $WLS; do
  timeout 600 python bench.py --workload $WL > $OUT/${TAG}_bench_${WL}.json 2> $OUT/${TAG}_bench_${WL}.err
  tail -c 600 $OUT/${TAG}_bench_${WL}.json
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file $OUT/${TAG}_launches_${WL}.csv python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline \
      > $OUT/${TAG}_ncu_launch_${WL}.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'agg_(fwd|bwd)_kernel|leaf_entity|ripple_bwd|user_fwd|transform_(fwd|bwd)' \
      -s 40 -c 12 -f -o $OUT/${TAG}_full_${WL} python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline \
      > $OUT/${TAG}_ncu_full_${WL}.log 2>&1
  # the .ncu-rep files are tens of MB (gpurun brings back <= 64 MiB): summarise on the box, keep the report only if small
  python scripts/ncu_summary.py $OUT/${TAG}_full_${WL}.ncu-rep 6 > $OUT/${TAG}_ncu_full_${WL}.txt 2>/dev/null
  python scripts/traffic_from_ncu.py $OUT/${TAG}_full_${WL}.ncu-rep $WL 2 $OUT/${TAG}_traffic.json > /dev/null 2>&1
  for K in ripple_bwd_kernel user_fwd_kernel "agg_bwd_kernel<(int)"; do
    python scripts/ncu_lines.py $OUT/${TAG}_full_${WL}.ncu-rep "$K" 12 >> $OUT/${TAG}_ncu_lines_${WL}.txt 2>/dev/null
  done
  [ $(stat -c %s $OUT/${TAG}_full_${WL}.ncu-rep) -gt 20000000 ] && rm -f $OUT/${TAG}_full_${WL}.ncu-rep
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
tail -c 400 $OUT/${TAG}_bench_reference.json
