#!/bin/bash
# round 2, visit I: parity tests (device-resident decoder), bench lines, protocol bench, sanitizers
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/r2i_pytest.txt
echo "== bench"; timeout 400 python bench.py 2>$OUT/r2i_bench.err | tee $OUT/r2i_bench.json | cut -c1-2500
tail -5 $OUT/r2i_bench.err
echo "== bench --serial"; timeout 300 python bench.py --serial --no-cpu 2>>$OUT/r2i_bench.err | tee $OUT/r2i_bench_serial.json | cut -c1-700
echo "== bench --matvec-path no-tc"; timeout 300 python bench.py --matvec-path no-tc --no-cpu 2>>$OUT/r2i_bench.err | tee $OUT/r2i_bench_notc.json | cut -c1-700
echo "== bench --no-graph"; timeout 300 python bench.py --no-graph --no-cpu 2>>$OUT/r2i_bench.err | tee $OUT/r2i_bench_nograph.json | cut -c1-700
echo "== bench 1Mi"; timeout 300 python bench.py --batch 1048576 --sets 2 --no-cpu 2>>$OUT/r2i_bench.err | tee $OUT/r2i_bench_1Mi.json | cut -c1-700
tail -5 $OUT/r2i_bench.err
echo "== protocol"; timeout 600 python tools/bench_protocol.py 2>&1 | tee $OUT/r2i_protocol.jsonl | cut -c1-400
echo "== memcheck"; timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_ntl.py tests/test_gpu_protocol.py -x -q -k "tensor_core or golden or vandermonde_vs_oracle or device_resident or fft_vs_oracle" 2>&1 | tail -8 | tee $OUT/r2i_memcheck.txt
echo "== racecheck"; timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_ntl.py tests/test_gpu_robust.py -x -q -k "tensor_core or golden or robust_decode_kats or wb_golden" 2>&1 | tail -8 | tee $OUT/r2i_racecheck.txt
