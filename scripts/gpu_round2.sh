#!/bin/bash
# One gpurun call of round 2: parity tests, smoke, bench lines, ncu launch list + DRAM traffic of every kernel of a step,
# ncu --set full of the heavy row kernels (B200_PROFILING.md recipe).
# usage: scripts/gpu_round2.sh <tag> [--no-tests] [--full "C4 C3"] [workloads...]
set -u
TAG=${1:-r02}
shift || true
TESTS=1
FULL=""
if [ "${1:-}" = "--no-tests" ]; then TESTS=0; shift; fi
if [ "${1:-}" = "--full" ]; then FULL=$2; shift; shift; fi
WLS=${@:-C4}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_smi.csv 2>&1
if [ $TESTS = 1 ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -rs --durations=8 > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  tail -14 $OUT/${TAG}_pytest.log
  timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
fi
for WL in $WLS; do
  timeout 900 python bench.py --workload $WL > $OUT/${TAG}_bench_${WL}.json 2> $OUT/${TAG}_bench_${WL}.err
  echo "bench $WL exit $?"; tail -c 1200 $OUT/${TAG}_bench_${WL}.json; tail -3 $OUT/${TAG}_bench_${WL}.err | cut -c1-300
  # launch list + DRAM bytes of every kernel (3 warm-up steps + 2 steps; the last complete step is summarised)
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv \
      --log-file $OUT/${TAG}_launches_${WL}.csv python bench.py --workload $WL --steps 2 --warmup 3 --quick --no-cpu-baseline \
      --no-parity-check > $OUT/${TAG}_ncu_launch_${WL}.log 2>&1
  python scripts/traffic_from_csv.py $OUT/${TAG}_launches_${WL}.csv $WL $OUT/${TAG}_traffic.json > $OUT/${TAG}_traffic_${WL}.txt 2>&1
  tail -5 $OUT/${TAG}_traffic_${WL}.txt
done
for WL in $FULL; do
  timeout 1200 ncu --set full --clock-control none --import-source on \
      -k regex:'agg_(fwd|bwd)|leaf_entity|table_|virt_|rel(q|dv|drk)_tc|ripple_bwd|user_fwd' \
      -s 51 -c 17 -f -o $OUT/${TAG}_full_${WL} python bench.py --workload $WL --steps 2 --warmup 3 --quick --no-cpu-baseline \
      --no-parity-check > $OUT/${TAG}_ncu_full_${WL}.log 2>&1
  # the report is read back and summarised in the container (scripts/ncu_summary.py / ncu_lines.py): ncu --page source on
  # the GPU box stalled two calls for 15+ minutes
  timeout 120 python scripts/ncu_summary.py $OUT/${TAG}_full_${WL}.ncu-rep 10 > $OUT/${TAG}_ncu_full_${WL}.txt 2>/dev/null
  [ $(stat -c %s $OUT/${TAG}_full_${WL}.ncu-rep) -gt 45000000 ] && rm -f $OUT/${TAG}_full_${WL}.ncu-rep
done
