#!/bin/bash
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 tools/gather_probe.py 2>$OUT/r2o_probe_n$N.err | grep "^{" | tee $OUT/r2o_gather_probe_n$N.json
tail -5 $OUT/r2o_probe_n$N.err | cut -c1-300
