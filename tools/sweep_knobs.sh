#!/bin/bash
run() { echo -n "$1 : "; env $2 timeout 120 python bench.py --no-cpu-baseline --steps 30 $3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['e2e']['value'],1))"; }
run "default" "A=1" ""
run "sub1_below=148" "HGK_TC2_SUB1_BELOW=148" ""
run "sub1_below=1000" "HGK_TC2_SUB1_BELOW=1000" ""
run "sub1_below=4000" "HGK_TC2_SUB1_BELOW=4000" ""
run "streams 8/2" "A=1" "--streams 8 --low-streams 2"
run "streams 8/3" "A=1" "--streams 8 --low-streams 3"
run "streams 4/1" "A=1" "--streams 4 --low-streams 1"
run "streams 6/1" "A=1" "--streams 6 --low-streams 1"
run "streams 6/3" "A=1" "--streams 6 --low-streams 3"
run "default again" "A=1" ""
