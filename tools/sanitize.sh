#!/bin/bash
# compute-sanitizer passes over the small-shape kernel parity tests (race / barrier checks of the
# hand-rolled mbarrier, named-barrier and TMEM lifetimes).  Logs go to gpurun_out/<tag>_sanitize_<tool>.log; the
# summaries are copied to profiles/ by hand.
# usage: bash tools/sanitize.sh <tag> [pytest -k expression]
TAG=${1:-san}
KEXPR=${2:-"(conv_tc_fwd_3xtf32 or conv_tc_dgrad_bnapply or conv_tc_dgrad_1xtf32 or conv_tc_dgrad_bnstats or conv_wgrad_tc or conv_tc_fused_bn_finalize or pool_upsample_bwd_with_fused or stem_fwd_and_wgrad or conv_skinny) and (shape0 or shape5 or shape15 or shape16 or (shape1 and not (shape10 or shape11 or shape12 or shape13 or shape14 or shape17 or shape18 or shape19)))"}
OUT=gpurun_out
mkdir -p $OUT
for TOOL in memcheck racecheck synccheck; do
  LOG=$OUT/${TAG}_sanitize_${TOOL}.log
  timeout 420 compute-sanitizer --tool $TOOL --print-limit 20 \
      python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "$KEXPR" > $LOG 2>&1
  echo "rc=$?" >> $LOG
  echo "== $TOOL: $(grep -c 'passed\|failed' $LOG) result lines"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" $LOG | tail -5
done
