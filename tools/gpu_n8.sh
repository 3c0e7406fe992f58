#!/bin/bash
# On an 8-GPU box: the 8-GPU bench line (weak scaling: 4096 envs per GPU = BASELINE config 5 in total).
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$?"
tail -c 600 gpurun_out/bench_n8.json
