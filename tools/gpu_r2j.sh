#!/bin/bash
# round 2, multi-GPU visit: N = $1 ranks
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/r2j_smi_n$N.txt 2>&1
nvidia-smi topo -m >> $OUT/r2j_smi_n$N.txt 2>&1
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_protocol.py -m gpu -x -q -k "multi_gpu or single_gpu or party_simulation_nccl" 2>&1 | tail -15 | tee $OUT/r2j_pytest_multi_n$N.txt
for g in auto copy; do
echo "== bench N=$N gather=$g"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --gather $g 2>$OUT/r2j_bench_n${N}_$g.err | tee $OUT/r2j_bench_n${N}_$g.json | cut -c1-3000
tail -3 $OUT/r2j_bench_n${N}_$g.err
done
echo "== bench N=1 (same box)"; timeout 600 python bench.py --steps 20 --warmup 3 2>$OUT/r2j_bench_n1.err | tee $OUT/r2j_bench_n1_on_n$N.json | cut -c1-600
echo "== reference arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 20 --warmup 3 2>/dev/null | tee $OUT/r2j_bench_reference_n$N.json | cut -c1-500
