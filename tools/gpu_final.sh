#!/bin/bash
# final check of a round on one GPU: parity tests, smoke, default bench line, ncu capture of the two kernels
TAG=${1:-final}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/${TAG}_pytest.txt
echo "== smoke"; timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench"; timeout 300 python bench.py 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json | python tools/brief.py default; tail -2 $OUT/${TAG}_bench.err
echo "== ncu full"; timeout 200 ncu --set full --clock-control none --import-source on -k regex:"ntt16_g4|interp_small" -c 4 -f -o $OUT/${TAG}_prof python bench.py --serial --sets 1 --steps 1 --warmup 3 --no-cpu > $OUT/${TAG}_ncu.log 2>&1; tail -1 $OUT/${TAG}_ncu.log | cut -c1-200
