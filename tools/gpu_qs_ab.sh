#!/bin/bash
# On the GPU box: variants of the quad slab kernel built with tools (ab/libppo_<name>.so.bin), timed back to back.
cp ppo_b200/csrc/libppo_b200.so /tmp/libppo_keep.so
for v in base keepr nog4 r104 r104keepr nog4keepr; do
  cp ab/libppo_$v.so.bin ppo_b200/csrc/libppo_b200.so
  echo "== $v"; PPO_TIME_PATH=qslab PPO_TIME_CHECK=1 timeout 60 python tools/time_tslab.py 0 1 2>&1 | tail -3
done
cp /tmp/libppo_keep.so ppo_b200/csrc/libppo_b200.so
