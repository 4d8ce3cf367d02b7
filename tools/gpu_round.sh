#!/bin/bash
# One full evidence visit to the GPU box: parity tests, smoke, bench lines (config 2 with the CPU arm; configs 3 and 5),
# per-launch timing, graph timeline, ncu launch list of one eager step, ncu --set full of the big-layer kernels in isolation.
# usage (from the repo root, on the box): bash tools/gpu_round.sh <tag> [skip-tests]
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
if [ "$2" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
  tail -4 $OUT/${TAG}_pytest.log
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 3500 $OUT/${TAG}_bench.json
timeout 300 python bench.py --config 5 --no-cpu-baseline > $OUT/${TAG}_bench_config5.json 2>> $OUT/${TAG}_bench.err; cut -c 1-400 $OUT/${TAG}_bench_config5.json
timeout 300 python bench.py --config 3 > $OUT/${TAG}_bench_config3.json 2>> $OUT/${TAG}_bench.err; cut -c 1-400 $OUT/${TAG}_bench_config3.json
timeout 300 python tools/bench_aug.py --cpu-baseline > $OUT/${TAG}_bench_aug.json 2>> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench_aug.json
timeout 300 python tools/time_plan.py --top 30 --filter conv > $OUT/${TAG}_time_plan.txt 2>&1
head -34 $OUT/${TAG}_time_plan.txt
timeout 300 python tools/graph_timeline.py --out $OUT/${TAG}_graph_timeline.json > $OUT/${TAG}_graph_timeline.txt 2>&1
python tools/alone_time.py $OUT/${TAG}_graph_timeline.json >> $OUT/${TAG}_graph_timeline.txt 2>&1
tail -40 $OUT/${TAG}_graph_timeline.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv \
   --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
python tools/agg_launches.py $OUT/${TAG}_launches.csv 24 > $OUT/${TAG}_launches_summary.txt 2>&1
head -30 $OUT/${TAG}_launches_summary.txt
# ncu --set full of the big-layer kernels in isolation (cold inputs: rotating buffers > L2), third repetition
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc2_kernel|conv_tc3_kernel|wgrad2_tc_kernel" -s 18 -c 9 \
   -o $OUT/${TAG}_kernels -f python tools/prof_kernels.py 3 > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log
ls -la $OUT | tail -12
