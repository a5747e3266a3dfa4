#!/bin/bash
# One short gpurun call: the variant parity tests (PS_only / HO_only / n_mix_hop > 1 / the other parameter_ablation.py
# settings) and the rest of the GPU suite, each in its own process, side by side on the one GPU; -v on the variants so
# that every verdict is in the log even if the call is cut off.
# usage: scripts/gpu_variants.sh <tag>
set -u
TAG=${1:-r02v}
OUT=gpurun_out
mkdir -p $OUT
( timeout 75 python -u -m pytest tests/test_variants_gpu.py -v --tb=short -p no:cacheprovider > $OUT/${TAG}_variants.log 2>&1
  echo "variants exit $?" >> $OUT/${TAG}_variants.log ) &
( timeout 75 python -u -m pytest tests -m gpu -q --tb=short -rs -p no:cacheprovider --ignore=tests/test_variants_gpu.py > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest.log ) &
wait
grep -c PASSED $OUT/${TAG}_variants.log; grep -E "FAILED|ERROR" $OUT/${TAG}_variants.log | head -40; tail -3 $OUT/${TAG}_variants.log
tail -12 $OUT/${TAG}_pytest.log
