#!/bin/bash
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_protocol.py -m gpu -x -q -k "multi_gpu or single_gpu or party_simulation" > $OUT/r2m_pytest_multi_n$N.txt 2>&1; grep -v "^frame" $OUT/r2m_pytest_multi_n$N.txt | tail -12
echo "== bench N=$N gather=auto"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 2>$OUT/r2m_bench_n${N}.err | tee $OUT/r2m_bench_n${N}.json | python tools/brief2.py
grep -n "Error" $OUT/r2m_bench_n${N}.err | head -3
for ctas in 8 32 64; do
echo "== bench N=$N ctas=$ctas"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 20 --warmup 3 --gather-ctas $ctas --cfg5 off --no-cpu 2>$OUT/r2m_bench_n${N}_c$ctas.err | tee $OUT/r2m_bench_n${N}_c$ctas.json | python tools/brief2.py
done
echo "== bench N=$N p2p"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 20 --warmup 3 --gather p2p --cfg5 off --no-cpu 2>$OUT/r2m_bench_n${N}_p2p.err | tee $OUT/r2m_bench_n${N}_p2p.json | python tools/brief2.py
echo "== party sim"; for byz in 0 $(( (N-1)/3 )); do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/bench_party_sim.py --byzantine $byz 2>/dev/null | grep "^{" | tee -a $OUT/r2m_party_sim_n$N.jsonl; done
