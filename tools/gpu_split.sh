#!/bin/bash
# sweep of the SM split between the encode and the interpolation launches (bench.py --sm-split)
mkdir -p gpurun_out
for e in ${SPLITS:-0 96 100 102 104 106 108 112}; do
  python bench.py --no-cpu --cfg5 off --sm-split $e 2>>gpurun_out/split.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'encode_sms': $e, 'us_per_step': round(d['ms_per_step']*1e3,2), 'value': d['value']}))" | tee -a gpurun_out/r2f_sm_split.jsonl
done
