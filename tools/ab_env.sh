#!/bin/bash
# A/B of environment knobs on the config-2 step (bench without the CPU arm): each argument is one "VAR=value[,VAR2=value2]"
# setting ("-" = defaults).  usage: bash tools/ab_env.sh - HGK_SPLITK=0 HGK_SPLITK_MT=12
for cfg in "$@"; do
  echo "== $cfg"
  if [ "$cfg" = "-" ]; then envs=""; else envs=$(echo "$cfg" | tr ',' ' '); fi
  env $envs timeout 200 python bench.py --no-cpu-baseline --steps 20 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])"
done
