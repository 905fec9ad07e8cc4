#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
# hyp: bits 8.. = pf+1 override
for cfg in "6 6 6 65536 16 0 6 20 12" "6 6 6 1048576 16 0 6 10 12" "6 6 6 1048576 16 512 6 10 12" "6 6 6 1048576 16 1024 6 10 12" "6 6 6 1048576 16 4 6 10 12" "6 6 6 1048576 16 28 6 10 12" "6 6 6 1048576 16 0 4 10 12" "6 6 6 1048576 16 0 5 10 8" "6 6 6 1048576 16 0 6 10 16" "6 16 8 65536 16 0 5 20 16" "6 16 8 1048576 16 0 5 10 16" "6 16 8 1048576 16 4 5 10 16" "6 16 8 1048576 16 0 5 10 8"; do
  echo "== tc_probe $cfg"; timeout 60 tools/tc_probe $cfg 2>&1 | grep -E "^d=|raw acc|outputs|time|error|fail"
done | tee $OUT/r2e_tc_probe.txt
