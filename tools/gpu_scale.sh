#!/bin/bash
# final scaling visit at N ranks ($1): bench lines of the final code
N=${1:-8}
OUT=gpurun_out; mkdir -p $OUT
echo "== bench N=1 (same box)"; timeout 600 python bench.py --steps 20 --warmup 3 2>$OUT/r2g_bench_n1_on_n$N.err | tee $OUT/r2g_bench_n1_on_n$N.json | python tools/brief2.py
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 2>$OUT/r2g_bench_n${N}.err | tee $OUT/r2g_bench_n${N}.json | python tools/brief2.py
grep -n "Error" $OUT/r2g_bench_n${N}.err | head -3
if [ "$N" -le 4 ]; then
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_protocol.py -m gpu -x -q -k "multi_gpu or party_simulation_nccl" > $OUT/r2g_pytest_multi_n$N.txt 2>&1; grep -v "^frame" $OUT/r2g_pytest_multi_n$N.txt | tail -4
fi
if [ "$N" -eq 2 ]; then
echo "== N=2 with the encode / interpolate SM split"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --cfg5 off --sm-split 104 2>>$OUT/r2g_bench_n${N}.err | tee $OUT/r2g_bench_n${N}_split104.json | python tools/brief2.py
echo "== party sim"; for b in 0 ; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/bench_party_sim.py 2>/dev/null | tail -1 | tee -a $OUT/r2g_party_sim.jsonl | cut -c1-200; done
fi
if [ "$N" -eq 8 ]; then
echo "== reference arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus $N --steps 5 --warmup 1 2>/dev/null | tail -1 | tee $OUT/r2g_bench_reference_n8_torchrun.json | cut -c1-300
fi
