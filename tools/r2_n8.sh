#!/bin/bash
# BASELINE.json configs #2-#5 on N GPUs of one box (default 8): one bench line per config, as the driver launches bench.py.
# tools/r2_n8.sh [N] ["qvh charades ..."]
set -u
N=${1:-8}
CONFIGS=${2:-"qvh charades anet generate"}
O=gpurun_out
mkdir -p $O
for c in $CONFIGS; do
  ( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py \
      --gpus $N --config $c --steps 6 --warmup 3 --no-eager --no-cpu-baseline ) > $O/n${N}_bench_$c.json 2> $O/n${N}_bench_$c.err
  cut -c1-230 $O/n${N}_bench_$c.json; grep -c . $O/n${N}_bench_$c.err
done
