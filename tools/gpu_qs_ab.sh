#!/bin/bash
# On the GPU box: library variants (ab/libppo_<name>.so.bin, built by hand) timed back to back (development).
cp ppo_b200/csrc/libppo_b200.so /tmp/libppo_keep.so
for round in 1 2; do
for v in ${AB_VARIANTS:-base}; do
  cp ab/libppo_$v.so.bin ppo_b200/csrc/libppo_b200.so
  echo -n "== $v "; PPO_TIME_PATH=${AB_PATH:-qslab} timeout 60 python tools/time_tslab.py 0 2>&1 | tail -1
done
done
cp /tmp/libppo_keep.so ppo_b200/csrc/libppo_b200.so
