#!/bin/bash
# ncu passes over one steady-state train step of bench.py (B200_PROFILING.md recipe): per-launch durations and DRAM bytes.
# Numbers printed by bench.py under ncu are not bench values. usage: tools/ncu_step.sh <out-prefix>
P=${1:-gpurun_out/r02}
ARGS="bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-roofline"
SKIP=${SKIP:-1900}
COUNT=${COUNT:-540}   # launches of one steady-state step = gpu_launches / steps of a bench.py line
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $SKIP -c $COUNT --csv --log-file ${P}_launches_bench_step.csv python $ARGS > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip $SKIP -c $COUNT --csv --log-file ${P}_dram_bench_step.csv python $ARGS > /dev/null 2>&1
python tools/launch_summary.py ${P}_launches_bench_step.csv > ${P}_launch_summary.txt 2>&1
python - <<PY
import csv, json
rows = [l for l in open("${P}_dram_bench_step.csv") if not l.startswith("==")]
rd = wr = 0.0
n = 0
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for x in csv.DictReader(rows):
    v = float(x["Metric Value"].replace(",", "")) * mult.get(x["Metric Unit"], 1)
    if "read" in x["Metric Name"]: rd += v
    else: wr += v
    n += 1
json.dump({"dram_bytes_per_step": rd + wr, "read": rd, "write": wr, "launches": n // 2,
           "how": "ncu dram__bytes_read.sum + dram__bytes_write.sum summed over a ${COUNT}-launch steady-state window (~one train step) of bench.py, B = 320"},
          open("${P}_step_dram_traffic.json", "w"))
print(open("${P}_step_dram_traffic.json").read())
PY
