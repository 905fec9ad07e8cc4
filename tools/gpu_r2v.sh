#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/r2v_pytest.txt 2>&1; tail -8 $OUT/r2v_pytest.txt
