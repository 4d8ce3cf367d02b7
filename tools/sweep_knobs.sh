#!/bin/bash
# A/B of scheduling / configuration knobs on the bench step (developer tool): bash tools/sweep_knobs.sh
run() { echo -n "$1 : "; env $2 timeout 120 python bench.py --no-cpu-baseline --steps 30 $3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['launches_per_step'])"; }
run "default" "A=1" ""
run "deferred skips" "HGK_DEFER_SKIPS=1" ""
run "deferred skips + 2-CTA tiny kernels" "HGK_DEFER_SKIPS=1 HGK_TC_BIG_OFF=1" ""
run "deferred skips, streams 8/3" "HGK_DEFER_SKIPS=1" "--streams 8 --low-streams 3"
run "deferred skips + 2-CTA tiny, streams 8/3" "HGK_DEFER_SKIPS=1 HGK_TC_BIG_OFF=1" "--streams 8 --low-streams 3"
run "low skips, streams 8/3" "HGK_LOW_SKIPS=1" "--streams 8 --low-streams 3"
