#!/bin/bash
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $OUT/r2n_pytest_multi_n$N.txt 2>&1; grep -v "^frame" $OUT/r2n_pytest_multi_n$N.txt | tail -12
echo "== bench N=$N gather=auto(ce)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 2>$OUT/r2n_bench_n${N}.err | tee $OUT/r2n_bench_n${N}.json | python tools/brief2.py
grep -n "Error" $OUT/r2n_bench_n${N}.err | head -3
for g in mc nccl; do
echo "== bench N=$N $g"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 20 --warmup 3 --gather $g --cfg5 off --no-cpu 2>$OUT/r2n_bench_n${N}_$g.err | tee $OUT/r2n_bench_n${N}_$g.json | python tools/brief2.py
done
