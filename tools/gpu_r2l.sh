#!/bin/bash
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest (1 GPU part: fnt, offline)"; timeout 900 python -m pytest tests/test_gpu_ntl.py tests/test_gpu_protocol.py -m gpu -x -q -k "fnt or offline or tensor_core or device_resident" > $OUT/r2l_pytest_new.txt 2>&1; tail -15 $OUT/r2l_pytest_new.txt
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_protocol.py -m gpu -x -q -k "multi_gpu or single_gpu or party_simulation_nccl" > $OUT/r2l_pytest_multi_n$N.txt 2>&1; grep -v "^frame" $OUT/r2l_pytest_multi_n$N.txt | tail -30
for g in auto; do
echo "== bench N=$N gather=$g"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --gather $g 2>$OUT/r2l_bench_n${N}_$g.err | tee $OUT/r2l_bench_n${N}_$g.json | python tools/brief2.py
tail -3 $OUT/r2l_bench_n${N}_$g.err | cut -c1-300
done
for ctas in 8 32; do
echo "== bench N=$N ctas=$ctas"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 20 --warmup 3 --gather-ctas $ctas --cfg5 off --no-cpu 2>$OUT/r2l_bench_n${N}_c$ctas.err | tee $OUT/r2l_bench_n${N}_c$ctas.json | python tools/brief2.py
done
