#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
export TC_TRACE=1
for cfg in "6 6 6 1048576 16 0 6 5 12" "6 6 6 1048576 16 4 6 5 12" "6 6 6 1048576 16 28 6 5 12" "6 16 8 1048576 16 0 5 5 16"; do
  echo "== tc_probe $cfg"; timeout 60 tools/tc_probe $cfg 2>&1 | grep -vE "^smem|raw mismatch|output mismatch"
done | tee $OUT/r2f_tc_trace.txt
