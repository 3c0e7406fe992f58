#!/bin/bash
# N-GPU bench with the moment exchange inside the GAE launch (NVLink peer mailboxes) against the NCCL all-reduce
N=${1:-2}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 100 --warmup 5 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$2', 'value', '%.4g' % d['value'], 'step ms', round(d['ms_per_step'],4), 'launches', d['gpu_launches'], 'e2e ms', round(d['e2e']['ms_per_step'],2))"; }
PPO_PEER_EXCHANGE=0 run 29561 nccl
PPO_PEER_EXCHANGE=1 run 29562 peers
PPO_PEER_EXCHANGE=0 run 29563 nccl
PPO_PEER_EXCHANGE=1 run 29564 peers
