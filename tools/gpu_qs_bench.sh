#!/bin/bash
# On the GPU box: the bench pass with PATH_AUTO on the transposed slab kernel and on the quad slab kernel, back to back.
for round in 1 2; do
for v in 0 1; do
  PPO_TVF_QSLAB=$v timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-other-configs 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); t=d['table_redraw']; print('qslab=$v', d['roofline']['kernel'][:32], 'step', round(d['ms_per_step'],4), 'fixed', round(t['fixed_table_ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'fixed kernel', round(t['fixed_table_kernel_ms'],4), 'match', d['e2e']['matches_resident'], 'e2e ms', round(d['e2e']['ms_per_step'],3))"
done
done
