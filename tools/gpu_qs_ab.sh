#!/bin/bash
# On the GPU box: the quad slab kernel under a few settings (development).
for t in 2 4 8; do
  echo "tiles $t"; PPO_QS_TILES=$t PPO_TIME_PATH=qslab timeout 90 python tools/time_tslab.py 0 1 8 24 2>&1 | tail -4
done
