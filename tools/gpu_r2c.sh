#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for cfg in "6 6 6 65536 0 0 4 20 12" "6 6 6 65536 0 0 6 20 12" "6 6 6 65536 0 0 3 20 12" "6 6 6 1048576 0 0 6 10 12" "6 6 6 1048576 0 2 6 10 12" "6 6 6 1048576 0 4 6 10 12" "6 16 8 65536 0 0 5 20 16" "6 16 8 1048576 0 0 5 10 16" "6 16 8 1048576 0 2 5 10 16" "6 16 8 1048576 0 4 5 10 16" "6 6 6 1000 0 0 4 20 12" "6 6 6 1048576 0 0 6 10 8"; do
  echo "== tc_probe $cfg"; timeout 60 tools/tc_probe $cfg 2>&1 | grep -E "^d=|raw|outputs|time|error|fail"
done | tee $OUT/r2c_tc_probe.txt
