#!/bin/bash
# 2 GPUs: copy-engine gather with the peer block cut into 1 / 2 / 3 / 4 pieces (HBMPC_CE_PIECES)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
for p in 1 3 2 4; do
  HBMPC_CE_PIECES=$p timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29615 \
    bench.py --gpus 2 --no-cpu --cfg5 off 2>>gpurun_out/ce_pieces.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'ce_pieces': $p, 'us_per_step': round(d['ms_per_step']*1e3,2), 'value': d['value'], 'mode': d['config']['parallelism']}))" | tee -a gpurun_out/r2o_ce_pieces.jsonl
done
