#!/bin/bash
# Does the reference arm (16 host processes) running first change our e2e number on the same box?
B="python bench.py --steps 20 --warmup 3 --no-other-configs --no-cpu-baseline"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('$1 e2e ms', round(d['e2e']['ms_per_step'],2), 'step ms', round(d['ms_per_step'],4))"; }
$B 2>/dev/null | pick ours-1
$B 2>/dev/null | pick ours-2
python bench.py --impl reference --steps 2 --warmup 1 > /dev/null 2>&1; echo "reference arm done"
$B 2>/dev/null | pick ours-after-ref-1
$B 2>/dev/null | pick ours-after-ref-2
sleep 20
$B 2>/dev/null | pick ours-after-sleep
free -g | head -2; nproc
