#!/bin/bash
# round 2: scaling visit at N ranks ($1) -- bench lines (default gather + mc), N = 1 on the same box,
# the multi-rank tests, the party simulation with and without Byzantine parties
N=${1:-4}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > $OUT/r2q_smi_n$N.txt
if [ "$N" -le 4 ]; then
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_protocol.py -m gpu -x -q -k "multi_gpu or party_simulation_nccl" > $OUT/r2q_pytest_multi_n$N.txt 2>&1; grep -v "^frame" $OUT/r2q_pytest_multi_n$N.txt | tail -6
fi
echo "== bench N=1 (same box)"; timeout 600 python bench.py --steps 20 --warmup 3 2>$OUT/r2q_bench_n1_on_n$N.err | tee $OUT/r2q_bench_n1_on_n$N.json | python tools/brief2.py
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 2>$OUT/r2q_bench_n${N}.err | tee $OUT/r2q_bench_n${N}.json | python tools/brief2.py
grep -n "Error" $OUT/r2q_bench_n${N}.err | head -3
echo "== bench N=$N mc"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --gather mc --cfg5 off --no-cpu 2>$OUT/r2q_bench_n${N}_mc.err | tee $OUT/r2q_bench_n${N}_mc.json | python tools/brief2.py
echo "== reference arm"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 20 --warmup 3 2>/dev/null | tee $OUT/r2q_bench_reference_n$N.json | cut -c1-330
echo "== party sim"; for byz in 0 $(( (N-1)/3 )); do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/bench_party_sim.py --byzantine $byz 2>/dev/null | grep "^{" | tee -a $OUT/r2q_party_sim_n$N.jsonl; done
