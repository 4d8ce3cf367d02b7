#!/bin/bash
# One development iteration on the GPU box: parity tests, bench line, per-launch time plan, graph timeline and
# (optionally) an `ncu --set full` capture of the kernels matching a regex during one eager step.
# usage: bash tools/gpu_iter.sh <tag> [ncu kernel regex] [ncu launch count]
TAG=${1:-it}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -6 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"
cut -c 1-330 $OUT/${TAG}_bench.json
timeout 300 python tools/time_plan.py --top 30 --filter conv > $OUT/${TAG}_time_plan.txt 2>&1
head -32 $OUT/${TAG}_time_plan.txt
timeout 300 python tools/graph_timeline.py --out $OUT/${TAG}_graph_timeline.json > $OUT/${TAG}_graph_timeline.txt 2>&1
tail -4 $OUT/${TAG}_graph_timeline.txt
if [ -n "$2" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -c ${3:-10} \
     -o $OUT/${TAG}_ncu -f python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > $OUT/${TAG}_ncu.log 2>&1
  tail -3 $OUT/${TAG}_ncu.log
fi
