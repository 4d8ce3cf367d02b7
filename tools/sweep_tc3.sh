#!/bin/bash
# A/B of the conv_tc3 routing mask on the config-2 step (bench without the CPU arm)
for m in 0 13 15 5 4 12 1; do
  echo "== HGK_TC3_MASK=$m"
  HGK_TC3_MASK=$m timeout 200 python bench.py --no-cpu-baseline --steps 20 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])"
done
