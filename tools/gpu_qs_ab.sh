#!/bin/bash
# On the GPU box: the quad slab kernel under a few settings (development).
for t in 8 4 2; do
  echo "tiles $t"; PPO_QS_TILES=$t PPO_TIME_PATH=qslab timeout 60 python tools/time_tslab.py 0 1 2>&1 | tail -2
done
if [ "$1" == "ncu" ]; then
  PPO_TIME_PATH=qslab timeout 300 ncu --set full --clock-control none --import-source on -k regex:qslab -s 3 -c 1 \
    -o gpurun_out/r02_qslab_4096_adv -f python tools/time_tslab.py 0 > gpurun_out/ncu_qslab.log 2>&1
  tail -2 gpurun_out/ncu_qslab.log
fi
