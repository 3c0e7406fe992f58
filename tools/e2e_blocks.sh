#!/bin/bash
# development: end-to-end step time of bench.py for several host-pipeline block counts
for b in 6 8 12 16 24; do
  PPO_HOST_BLOCKS=$b python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('blocks $b e2e ms', d['e2e']['ms_per_step'])"
done
