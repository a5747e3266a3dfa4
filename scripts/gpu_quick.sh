#!/bin/bash
# quick GPU check: parity tests + short bench lines (no ncu).  usage: scripts/gpu_quick.sh <tag> [workloads...]
set -u
TAG=${1:-q}
shift || true
WLS=${@:-C2}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -4 $OUT/${TAG}_pytest.log
for WL in $WLS; do
  timeout 600 python bench.py --workload $WL --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_${WL}.json 2> $OUT/${TAG}_bench_${WL}.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_${WL}.json").read().strip().splitlines()[-1])
    print("$WL", "ms/step", round(d["ms_per_step"], 4), "pairs/s", round(d["value"]), "e2e", round(d["e2e"]["value"]))
    print(" ", d["kernels_ms_per_step"])
except Exception as e:
    print("bench $WL failed", e); print(open("$OUT/${TAG}_bench_${WL}.err").read()[-1500:])
PY
done
