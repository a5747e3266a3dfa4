#!/bin/bash
# A/B of one environment switch on the GPU box: parity tests, then bench lines with VAR unset and VAR=0.
# usage: scripts/gpu_ab.sh <tag> <VAR> [workloads...]
set -u
TAG=$1; VAR=$2; shift 2
WLS=${@:-C2}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
for WL in $WLS; do
  for MODE in on off; do
    if [ $MODE = off ]; then export $VAR=0; else unset $VAR; fi
    timeout 300 python bench.py --workload $WL --steps 200 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_${WL}_${MODE}.json 2> $OUT/${TAG}_${WL}_${MODE}.err
    python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_${WL}_${MODE}.json").read().strip().splitlines()[-1])
    print("$WL $VAR $MODE ms/step", round(d["ms_per_step"], 4), "pairs/s", round(d["value"]), "e2e", round(d["e2e"]["value"]),
          "prefetch", round(d.get("e2e_prefetch", {}).get("value", 0)), "devfeed", round(d.get("e2e_device_feed", {}).get("value", 0)))
except Exception as e:
    print("bench $WL $MODE failed", e); print(open("$OUT/${TAG}_${WL}_${MODE}.err").read()[-1500:])
PY
  done
done
