#!/bin/bash
# development: time the TVF slab kernels with individual stages switched off (PPO_TVF_DEBUG bits)
# usage: dbg_tvf.sh <grep pattern for the shape> <bits...>
pat=${1:-256x4096x128 done_p=0.03}; shift
for d in ${@:-0 1 2 4 8 16}; do echo "== PPO_TVF_DEBUG=$d"; PPO_TVF_DEBUG=$d timeout 120 python tools/micro_tvf.py dev 2>&1 | grep "$pat" | grep tslab; done
