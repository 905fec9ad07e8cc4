#!/bin/bash
# quick A/B on the GPU box: parity tests, then the headline bench with both interpolation arithmetics
TAG=${1:-ab}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/${TAG}_pytest.txt
for ar in 0 1; do
  echo "== bench HBG_INTERP_ARITH=$ar"
  HBG_INTERP_ARITH=$ar timeout 200 python bench.py --steps 300 --warmup 10 --no-cpu 2>>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_arith$ar.json | python tools/brief.py arith$ar
  HBG_INTERP_ARITH=$ar timeout 200 python bench.py --batch 1048576 --sets 1 --steps 30 --warmup 3 --no-cpu 2>>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_1Mi_arith$ar.json | python tools/brief.py 1Mi-arith$ar
done
echo "== ncu full"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:"ntt16_g4|interp_small" -c 4 -f -o $OUT/${TAG}_prof python bench.py --sets 1 --steps 1 --warmup 3 --no-cpu > $OUT/${TAG}_ncu.log 2>&1; tail -2 $OUT/${TAG}_ncu.log | cut -c1-300
