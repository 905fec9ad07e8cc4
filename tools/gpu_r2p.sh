#!/bin/bash
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu (all)"; timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/r2p_pytest_n$N.txt 2>&1; grep -v "^frame" $OUT/r2p_pytest_n$N.txt | tail -12
echo "== bench N=1"; timeout 600 python bench.py --steps 20 --warmup 3 2>$OUT/r2p_bench_n1.err | tee $OUT/r2p_bench_n1.json | python tools/brief2.py
echo "== bench N=1 1Mi"; timeout 600 python bench.py --batch 1048576 --sets 2 --no-cpu 2>>$OUT/r2p_bench_n1.err | tee $OUT/r2p_bench_n1_1Mi.json | python tools/brief2.py
tail -3 $OUT/r2p_bench_n1.err | cut -c1-300
echo "== gather probe"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 tools/gather_probe.py 2>$OUT/r2p_probe_n$N.err | grep "^{" | tee $OUT/r2p_gather_probe_n$N.json
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 2>$OUT/r2p_bench_n${N}.err | tee $OUT/r2p_bench_n${N}.json | python tools/brief2.py
grep -n "Error" $OUT/r2p_bench_n${N}.err | head -3
