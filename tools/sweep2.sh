#!/bin/bash
for v in 1 2 3 6; do
  HBG_INTERP_SPLIT=$v python tools/latency_probe.py | grep -E "^(65536|262144) " | sed "s/^/split$v /"
done
