#!/bin/bash
# sweep of the SM split between the encode and the interpolation launches (bench.py --sm-split)
mkdir -p gpurun_out
for e in 0 88 96 100 104 108 112 118; do
  python bench.py --no-cpu --cfg5 off --sm-split $e --steps 400 2>>gpurun_out/split.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('split',$e,'us/step',round(d['ms_per_step']*1e3,2),'value',d['value'])" | tee -a gpurun_out/split.txt
done
