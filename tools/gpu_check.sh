#!/bin/bash
# One GPU-box visit: parity tests, headline bench, microbench3, a short memcheck and
# an ncu capture of the two headline kernels.  Everything lands in gpurun_out/.
#   gpurun --timeout 900 -- 'bash tools/gpu_check.sh TAG'
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
echo "== pytest -m gpu"; timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.txt
echo "== bench"; timeout 300 python bench.py 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json | python tools/brief.py default
tail -3 $OUT/${TAG}_bench.err
echo "== bench --serial"; timeout 200 python bench.py --serial --no-cpu 2>>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_serial.json | python tools/brief.py serial
echo "== bench 1Mi (steady state)"; timeout 200 python bench.py --batch 1048576 --sets 1 --steps 30 --warmup 3 --no-cpu 2>>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_1Mi.json | python tools/brief.py 1Mi
echo "== reference arm"; timeout 200 python bench.py --impl reference --steps 20 --warmup 3 | tee $OUT/${TAG}_bench_reference.json | cut -c1-300
echo "== memcheck"; timeout 200 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_ntl.py -x -q -k "fft_vs_oracle or golden or vandermonde_vs_oracle" 2>&1 | tail -8 | tee $OUT/${TAG}_memcheck.txt
echo "== ncu full"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:"ntt16_g4|interp_small" -c 4 -f -o $OUT/${TAG}_prof python bench.py --serial --sets 1 --steps 1 --warmup 3 --no-cpu > $OUT/${TAG}_ncu.log 2>&1; tail -2 $OUT/${TAG}_ncu.log | cut -c1-300
echo "== ncu launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/${TAG}_launches.log 2>&1; grep -c "gpu__time_duration" $OUT/${TAG}_launches.csv
echo "== other configs"; timeout 400 python tools/bench_configs.py > $OUT/${TAG}_other_configs.jsonl 2>$OUT/${TAG}_other.err; cut -c1-330 $OUT/${TAG}_other_configs.jsonl
