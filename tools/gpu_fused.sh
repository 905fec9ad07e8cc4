#!/bin/bash
# fused all-gather (staged full-line stores from the interpolation epilogue) against the copy modes
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -15
for g in fused auto; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --no-cpu --cfg5 off --gather $g 2>gpurun_out/fused_$g.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$g', 'us/step',round(d['ms_per_step']*1e3,2),'value %.3e'%d['value'], d['config']['parallelism'], d['roofline'].get('nvlink',{}).get('frac'))" || tail -5 gpurun_out/fused_$g.err
done
