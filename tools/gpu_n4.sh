#!/bin/bash
# On an 4-GPU box: the 4-GPU bench line (weak scaling: 4096 envs per GPU = BASELINE config 5 in total).
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 4 --steps 50 --warmup 5 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "bench n4 rc=$?"
tail -c 600 gpurun_out/bench_n4.json
