#!/bin/bash
# round 2, visit H: parity tests on the tensor-core default + first bench lines
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/r2h_smi.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/r2h_pytest.txt
echo "== bench"; timeout 400 python bench.py 2>$OUT/r2h_bench.err | tee $OUT/r2h_bench.json | cut -c1-1500
tail -5 $OUT/r2h_bench.err
echo "== bench --serial"; timeout 300 python bench.py --serial --no-cpu 2>>$OUT/r2h_bench.err | tee $OUT/r2h_bench_serial.json | cut -c1-600
echo "== bench --matvec-path no-tc"; timeout 300 python bench.py --matvec-path no-tc --no-cpu 2>>$OUT/r2h_bench.err | tee $OUT/r2h_bench_notc.json | cut -c1-600
echo "== bench --no-graph"; timeout 300 python bench.py --no-graph --no-cpu 2>>$OUT/r2h_bench.err | tee $OUT/r2h_bench_nograph.json | cut -c1-600
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 20 --warmup 3 | tee $OUT/r2h_bench_reference.json | cut -c1-400
tail -5 $OUT/r2h_bench.err
