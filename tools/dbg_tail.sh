#!/bin/bash
# upper bound of what the tail term costs: PPO_TVF_DEBUG=16 skips it (results are wrong for the last rows)
B="python bench.py --steps 100 --warmup 5 --no-other-configs --no-cpu-baseline"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('$1 kernel ms', round(d['roofline']['kernel_ms'],4))"; }
$B 2>/dev/null | pick full
PPO_TVF_DEBUG=16 $B 2>/dev/null | pick no-tail
$B 2>/dev/null | pick full
PPO_TVF_DEBUG=16 $B 2>/dev/null | pick no-tail
