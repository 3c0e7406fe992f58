#!/bin/bash
# On the GPU box: the quad slab kernel beside the transposed slab kernel (time, equality), then its parity suite.
# Every step under a short timeout: a kernel built from hand-rolled barriers can hang instead of failing.
mkdir -p gpurun_out
PPO_TIME_PATH=qslab PPO_TIME_CHECK=1 timeout 90 python tools/time_tslab.py ${QS_DBG:-0 1} 2>&1 | tee gpurun_out/qslab_time.log
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "qslab timing run failed or hung"; exit 1; fi
PPO_TIME_PATH=tslab timeout 90 python tools/time_tslab.py 0 2>&1 | tee -a gpurun_out/qslab_time.log
if [ "$1" == "tests" ] || [ "$2" == "tests" ]; then
  timeout 600 python -m pytest tests/test_tvf_qslab_gpu.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/qslab_tests.log
fi
if [ "$1" == "ncu" ]; then
  PPO_TIME_PATH=qslab timeout 300 ncu --set full --clock-control none --import-source on -k regex:qslab -s 3 -c 1 \
    -o gpurun_out/r02_qslab_4096_adv -f python tools/time_tslab.py 0 > gpurun_out/ncu_qslab.log 2>&1
  tail -3 gpurun_out/ncu_qslab.log
fi
