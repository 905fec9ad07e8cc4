#!/bin/bash
for v in 0 1 2 3 4 5; do
  HBG_NTT16_VARIANT=$v python bench.py --steps 300 --warmup 10 --no-cpu 2>&1 | tail -1 | python tools/brief.py "variant$v"
done
