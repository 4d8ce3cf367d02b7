#!/bin/bash
# One GPU-box visit: parity tests, bench line, per-launch timing, ncu launch list, ncu full capture.
# usage (from the repo root, on the box): bash tools/gpu_round.sh <tag> [skip-tests]
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
if [ "$2" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
  tail -5 $OUT/${TAG}_pytest.log
fi
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 3000 $OUT/${TAG}_bench.json
timeout 300 python tools/time_plan.py --top 30 --filter conv > $OUT/${TAG}_time_plan.txt 2>&1
head -45 $OUT/${TAG}_time_plan.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv \
   --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
python tools/agg_launches.py $OUT/${TAG}_launches.csv 24 > $OUT/${TAG}_launches_summary.txt 2>&1
# ncu --set full of the big-layer kernels in isolation (cold inputs: rotating buffers > L2), third repetition
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc2_kernel|wgrad2_tc_kernel" -s 18 -c 9 \
   -o $OUT/${TAG}_kernels -f python tools/prof_kernels.py 3 > $OUT/${TAG}_ncu_full.log 2>&1
timeout 200 python tools/dbg_timeline2.py > $OUT/${TAG}_tile_kernel_trace.txt 2>&1
timeout 300 python tools/graph_timeline.py --out $OUT/${TAG}_graph_timeline.json > $OUT/${TAG}_graph_timeline.txt 2>&1
ls -la $OUT
