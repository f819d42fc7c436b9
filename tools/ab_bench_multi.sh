#!/bin/bash
# A/B of bench.py on N GPUs under environment switches: tools/ab_bench_multi.sh N "ENV=1" "ENV=0" ...
N=$1; shift
port=29600
for s in "$@"; do
  port=$((port+1))
  echo -n "$s: "; env $s python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 5 2>/dev/null | grep "^{" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'ms ', round(d['value']), 'samples/s')"
done
