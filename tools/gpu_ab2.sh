#!/bin/bash
# same box, same library: bench.py flag variants, alternated
run() {
  python bench.py --no-cpu --cfg5 off "$@" 2>>gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', 'us/step',round(d['ms_per_step']*1e3,2), d['roofline'].get('kernel_ms'))"
}
for rep in 1 2; do
  run --tc-store staged; run --tc-store direct
  run --tc-store staged --batch 1048576 --sets 2; run --tc-store direct --batch 1048576 --sets 2
done
