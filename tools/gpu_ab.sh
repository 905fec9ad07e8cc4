#!/bin/bash
# A/B of builds of the library on the SAME box: tools/ab/lib{A,B,...}.so, alternated
VARIANTS=${VARIANTS:-"A B"}
run() {
  python bench.py --no-cpu --cfg5 off $EXTRA 2>>gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1 us/step',round(d['ms_per_step']*1e3,2), d['roofline'].get('kernel_ms'))"
}
for rep in 1 2; do
  for v in $VARIANTS; do cp tools/ab/lib$v.so honeybadgermpc_b200/libhbmpc_b200.so; run $v; done
done
