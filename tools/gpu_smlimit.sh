#!/bin/bash
# N > 1: leave a few SMs to the gather's copy kernel (bench.py --sm-limit) -- sweep
N=${1:-4}
mkdir -p gpurun_out
run() {
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 \
    bench.py --gpus $N --no-cpu --cfg5 off "$@" 2>>gpurun_out/smlimit.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', 'us/step',round(d['ms_per_step']*1e3,2),'value %.3e'%d['value'], d['config']['parallelism'], 'nvlink frac', round(d['roofline'].get('nvlink',{}).get('frac',0),3))" || tail -5 gpurun_out/smlimit.err
}
shift
for cfg in "$@"; do run $cfg; done
