#!/bin/bash
# round 2, multi-GPU visit: N = $1 ranks
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_protocol.py -m gpu -x -q -k "multi_gpu or single_gpu or party_simulation_nccl" > $OUT/r2k_pytest_multi_n$N.txt 2>&1; grep -v "^frame" $OUT/r2k_pytest_multi_n$N.txt | tail -40
for g in auto fused; do
echo "== bench N=$N gather=$g"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --gather $g 2>$OUT/r2k_bench_n${N}_$g.err | tee $OUT/r2k_bench_n${N}_$g.json | python tools/brief2.py
tail -3 $OUT/r2k_bench_n${N}_$g.err | cut -c1-300
done
echo "== bench N=$N no-graph"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 --no-graph --cfg5 off 2>$OUT/r2k_bench_n${N}_nograph.err | tee $OUT/r2k_bench_n${N}_nograph.json | python tools/brief2.py
echo "== bench N=1 (same box)"; timeout 600 python bench.py --steps 20 --warmup 3 2>$OUT/r2k_bench_n1.err | tee $OUT/r2k_bench_n1_on_n$N.json | python tools/brief2.py
