#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for cfg in "6 6 6 65536 0 0 4 20 12" "6 6 6 1000 0 0 4 20 12" "6 6 6 1048576 0 0 4 10 12" "6 6 6 1048576 0 0 6 10 12" "6 6 6 1048576 0 4 4 10 12" "6 16 8 65536 0 0 3 20 16" "6 16 8 1048576 0 0 3 10 16" "6 16 8 1048576 0 4 3 10 16" "2 4 4 128 0 0 3 20 8" "3 5 5 777 0 0 3 20 8" "10 6 6 5000 0 0 3 20 12" "16 2 2 3000 0 0 2 20 8"; do
  echo "== tc_probe $cfg"; timeout 60 tools/tc_probe $cfg 2>&1 | grep -E "^d=|raw acc|raw mism|outputs|time|error|fail" | head -12
done | tee $OUT/r2g_tc_probe.txt
export TC_TRACE=1
for cfg in "6 6 6 1048576 0 0 4 5 12"; do
  echo "== tc_probe $cfg"; timeout 60 tools/tc_probe $cfg 2>&1 | grep -vE "^smem|raw mismatch|output mismatch"
done | tee $OUT/r2g_tc_trace.txt
