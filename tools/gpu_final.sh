#!/bin/bash
# round 2, FINAL single-GPU evidence visit: parity tests, bench lines, ncu (launch list + full capture of the
# tensor-core kernel), sanitizers, the other configs, the protocol bench.  Everything lands in gpurun_out/.
TAG=${1:-r2f}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.txt 2>&1; tail -6 $OUT/${TAG}_pytest.txt
echo "== bench"; timeout 600 python bench.py 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json | python tools/brief2.py
tail -3 $OUT/${TAG}_bench.err | cut -c1-300
echo "== bench --steps 20 --warmup 3 (the driver's flags)"; timeout 600 python bench.py --steps 20 --warmup 3 --cfg5 off 2>>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_driverflags.json | python tools/brief2.py
echo "== bench --serial"; timeout 300 python bench.py --serial --no-cpu --cfg5 off 2>>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_serial.json | python tools/brief2.py
echo "== bench no-tc"; timeout 300 python bench.py --matvec-path no-tc --no-cpu --cfg5 off 2>>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_notc.json | python tools/brief2.py
echo "== bench 1Mi"; timeout 300 python bench.py --batch 1048576 --sets 2 --no-cpu 2>>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_1Mi.json | python tools/brief2.py
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 20 --warmup 3 | tee $OUT/${TAG}_bench_reference.json | cut -c1-300
echo "== other configs"; timeout 600 python tools/bench_configs.py > $OUT/${TAG}_other_configs.jsonl 2>$OUT/${TAG}_other.err; cut -c1-400 $OUT/${TAG}_other_configs.jsonl; tail -3 $OUT/${TAG}_other.err
echo "== protocol"; timeout 600 python tools/bench_protocol.py 2>&1 | tee $OUT/${TAG}_protocol.jsonl | cut -c1-300
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-graph --cfg5 off --min-ms 0 > $OUT/${TAG}_launches.log 2>&1; grep -c "gpu__time_duration" $OUT/${TAG}_launches.csv
echo "== ncu full"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_apply" -s 24 -c 4 -f -o $OUT/${TAG}_prof python bench.py --serial --sets 1 --steps 1 --warmup 3 --no-cpu --no-graph --cfg5 off --min-ms 0 > $OUT/${TAG}_ncu.log 2>&1; tail -2 $OUT/${TAG}_ncu.log | cut -c1-300
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_ntl.py tests/test_gpu_protocol.py -x -q -k "tensor_core or golden or vandermonde_vs_oracle or device_resident or fnt or offline or single_gpu" > $OUT/${TAG}_memcheck.txt 2>&1; tail -6 $OUT/${TAG}_memcheck.txt
echo "== racecheck"; timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_ntl.py tests/test_gpu_robust.py -x -q -k "tensor_core or test_golden or robust_decode_kats or wb_golden" > $OUT/${TAG}_racecheck.txt 2>&1; tail -6 $OUT/${TAG}_racecheck.txt
echo "== probes"; timeout 120 tools/tmem_probe 2000 > $OUT/${TAG}_tmem_probe.txt 2>&1; tail -3 $OUT/${TAG}_tmem_probe.txt
(TC_TRACE=1 timeout 60 tools/tc_probe 6 16 8 303104 0 0 4 20 16 0; TC_TRACE=1 timeout 60 tools/tc_probe 6 6 6 303104 0 0 4 20 12 0; TC_TRACE=1 timeout 60 tools/tc_probe 6 16 8 65536 0 0 4 20 16 0) > $OUT/${TAG}_tc_role_trace.txt 2>&1
for h in 0 4 16 28 64 128 192; do echo "hyp $h"; timeout 60 tools/tc_probe 6 16 8 303104 0 $h 4 20 16 0 | tail -1; done > $OUT/${TAG}_tc_probe_hyp.txt 2>&1
timeout 120 python tools/tile_cost.py --limits 148,74,37 > $OUT/${TAG}_tile_cost.jsonl 2>&1
echo "== store / sm-split variants"; for v in "--tc-store staged" "--sm-split 104" "--serial"; do timeout 300 python bench.py --no-cpu --cfg5 off $v 2>>$OUT/${TAG}_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'variant': '$v', 'us_per_step': round(d['ms_per_step']*1e3,2), 'value': d['value'], 'kernel_ms': d['roofline']['kernel_ms']}))"; done > $OUT/${TAG}_bench_variants.jsonl; cat $OUT/${TAG}_bench_variants.jsonl
