#!/bin/bash
# quick visit of the GPU box: parity tests, headline bench (two streams / one stream), steady state, microbench3
TAG=${1:-ab}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/${TAG}_pytest.txt
echo "== bench (default: two streams)"
timeout 300 python bench.py 2>>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json | python tools/brief.py overlap
echo "== bench --serial"
timeout 200 python bench.py --serial --steps 300 --warmup 10 --no-cpu 2>>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_serial.json | python tools/brief.py serial
echo "== bench 1Mi"
timeout 200 python bench.py --batch 1048576 --sets 1 --steps 30 --warmup 3 --no-cpu 2>>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_1Mi.json | python tools/brief.py 1Mi
echo "== microbench3"; timeout 120 tools/microbench3 2>&1 | tee $OUT/${TAG}_microbench3.txt | tail -11
tail -5 $OUT/${TAG}_bench.err
