#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== other configs"; timeout 600 python tools/bench_configs.py > $OUT/r2t_other_configs.jsonl 2>$OUT/r2t_other.err; cut -c1-700 $OUT/r2t_other_configs.jsonl; tail -3 $OUT/r2t_other.err
