#!/bin/bash
# usage: tools/scale.sh N "modes"
N=$1; shift
for g in "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 300 --warmup 6 --gather $g 2>&1 | grep -E "^\{|bench\]|Error|error|Traceback" | cut -c1-3000 | python tools/brief.py N$N-$g
done
