#!/bin/bash
# A/B of two builds of the library on the SAME box: tools/ab/libA.so, tools/ab/libB.so, alternated
run() {
  python bench.py --no-cpu --cfg5 off 2>>gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1 us/step',round(d['ms_per_step']*1e3,2), d['roofline'].get('kernel_ms'))"
}
for rep in 1 2 3; do
  for v in A B; do cp tools/ab/lib$v.so honeybadgermpc_b200/libhbmpc_b200.so; run $v; done
done
