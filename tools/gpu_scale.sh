#!/bin/bash
# final scaling visit at N ranks ($1): bench lines of the final code
N=${1:-8}
OUT=gpurun_out; mkdir -p $OUT
echo "== bench N=1 (same box)"; timeout 600 python bench.py --steps 20 --warmup 3 2>$OUT/r2z_bench_n1_on_n$N.err | tee $OUT/r2z_bench_n1_on_n$N.json | python tools/brief2.py
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 2>$OUT/r2z_bench_n${N}.err | tee $OUT/r2z_bench_n${N}.json | python tools/brief2.py
grep -n "Error" $OUT/r2z_bench_n${N}.err | head -3
if [ "$N" -le 4 ]; then
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_protocol.py -m gpu -x -q -k "multi_gpu or party_simulation_nccl" > $OUT/r2z_pytest_multi_n$N.txt 2>&1; grep -v "^frame" $OUT/r2z_pytest_multi_n$N.txt | tail -4
fi
