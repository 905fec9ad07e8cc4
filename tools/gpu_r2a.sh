#!/bin/bash
# round 2, visit A: tensor-core probe (descriptor hypotheses) + parity tests
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/r2a_smi.txt 2>&1
for cfg in "6 6 6 65536 16 0" "6 6 6 65536 0 0" "6 6 6 65536 16 1" "6 16 8 65536 16 0" "6 6 6 1000 16 0" "6 6 3 65536 16 0"; do
  echo "== tc_probe $cfg"; timeout 60 tools/tc_probe $cfg 2>&1 | tail -12
done | tee $OUT/r2a_tc_probe.txt
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/r2a_pytest.txt
