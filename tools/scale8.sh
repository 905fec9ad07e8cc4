#!/bin/bash
# N-GPU variants of the headline bench (which gather / overlap is best at this N?)
N=${1:-8}
run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 300 --warmup 6 "$@" 2>&1 | grep -E "^\{|bench\]|Error|error|Traceback" | cut -c1-3000 | tee -a gpurun_out/scale_n${N}.jsonl | python tools/brief.py "N$N-$tag"; }
run fused
run fused-overlap --overlap-encode on
run copy32 --gather copy --gather-ctas 32
run copy32-overlap --gather copy --gather-ctas 32 --overlap-encode on
