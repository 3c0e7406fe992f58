#!/bin/bash
# On the GPU box: bench both libraries built by tools/ab_build.sh back to back (twice each).
for round in 1 2; do
for v in base var; do
  cp ab/libppo_$v.so.bin ppo_b200/csrc/libppo_b200.so
  python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-other-configs 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); t=d['table_redraw']; print('$v', 'step', round(d['ms_per_step'],4), 'fixed', round(t['fixed_table_ms_per_step'],4), 'kernel', round(t['fixed_table_kernel_ms'],4), 'match', d['e2e']['matches_resident'])"
done
done
cp ab/libppo_base.so.bin ppo_b200/csrc/libppo_b200.so
