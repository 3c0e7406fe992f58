#!/bin/bash
# On a 2-GPU box: the multi-GPU tests, the sharded pass check and the 2-GPU bench line.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_peer_exchange_gpu.py -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
tail -c 1500 gpurun_out/bench_n2.json
