#!/bin/bash
# Quick GPU visit: parity tests (bounded), bench line without the CPU arm, per-launch timing.
# usage: bash tools/gpu_quick.sh <tag> [pytest -k expression]
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
if [ -n "$2" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q -k "$2" > $OUT/${TAG}_pytest.log 2>&1
else
  timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
fi
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -15 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"
cut -c 1-400 $OUT/${TAG}_bench.json
timeout 300 python tools/time_plan.py --top 30 --filter conv > $OUT/${TAG}_time_plan.txt 2>&1
head -34 $OUT/${TAG}_time_plan.txt
