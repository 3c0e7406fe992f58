#!/bin/bash
# quick A/B aid: parity of the slab kernels, then the resident bench line (kernel time, roofline fraction)
python -m pytest tests/test_tvf_gpu.py tests/test_full_size_gpu.py tests/test_value_estimate_gpu.py -x -q 2>&1 | tail -3
B="python bench.py --steps 100 --warmup 5 --no-other-configs --no-cpu-baseline"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('$1 step ms', round(d['ms_per_step'],4), 'kernel ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4), 'e2e ms', round(d['e2e']['ms_per_step'],2))"; }
$B 2>/dev/null | pick run-1
$B 2>/dev/null | pick run-2
