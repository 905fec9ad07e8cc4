#!/bin/bash
# which role paces a tile: the probe's hypothesis flags on the cfg2 encode / interpolate shapes
# hyp: 2 = epilogue reads TMEM only, 4 = epilogue does nothing, 8 = no loads, 16 = no MMAs,
#      64 = no stores, 128 = no fold / reduce
for shape in "6 16 8 303104 0 H 4 20 16" "6 6 6 303104 0 H 4 20 12"; do
  for h in 0 16 80 144 208 64 128 192; do
    echo "== ${shape/H/$h}"; tools/tc_probe ${shape/H/$h} 2>&1 | grep -E "^time"
  done
done
