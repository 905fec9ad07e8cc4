#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest robust + protocol"; timeout 900 python -m pytest tests/test_gpu_robust.py tests/test_gpu_protocol.py -m gpu -x -q > $OUT/r2u_pytest.txt 2>&1; tail -8 $OUT/r2u_pytest.txt
echo "== other configs"; timeout 600 python tools/bench_configs.py > $OUT/r2u_other_configs.jsonl 2>$OUT/r2u_other.err; cut -c1-500 $OUT/r2u_other_configs.jsonl; tail -3 $OUT/r2u_other.err
