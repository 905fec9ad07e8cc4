#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for cfg in "6 6 6 65536 16 0 4 20 8" "6 6 6 65536 16 0 4 20 12" "6 6 6 65536 16 0 4 20 16" "6 6 6 65536 16 0 6 20 12" "6 16 8 65536 16 0 4 20 8" "6 16 8 65536 16 0 4 20 16" "6 6 6 1000 16 0 4 20 12" "6 6 6 1048576 16 0 4 10 12" "6 16 8 1048576 16 0 4 10 16" "2 4 4 128 16 0 3 20 8" "16 16 8 5462 16 0 3 20 16"; do
  echo "== tc_probe $cfg"; timeout 60 tools/tc_probe $cfg 2>&1 | tail -6
done | tee $OUT/r2b_tc_probe.txt
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/r2b_pytest.txt
