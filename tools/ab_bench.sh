#!/bin/bash
# A/B runs of bench.py under environment switches; prints ms/step per setting (measurement helper)
run() { echo -n "$1: "; env $1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-roofline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['ms_per_step'],3), ' clocks', d['clocks']['sm_mhz'])"; }
for s in "$@"; do run "$s"; done
