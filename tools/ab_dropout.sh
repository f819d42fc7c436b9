#!/bin/bash
# step time with and without dropout on the same box (what the counter-based masks cost in total)
for d in 0.1 0.0 0.1 0.0; do echo -n "dropout $d: "; python bench.py --steps 20 --warmup 5 --dropout $d --no-cpu-baseline --no-kernel-roofline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'ms')"; done
