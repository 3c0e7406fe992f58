#!/bin/bash
# Full GPU validation of round 2: parity tests, smoke, bench (both arms), ncu launch list of the bench command,
# one full ncu capture of the dominant kernel at the bench shape and one of the quad slab kernel.  Run under gpurun
# from the repo root.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
PPO_TVF_QSLAB=0 timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_tslab.log 2>&1; echo "pytest (PPO_TVF_QSLAB=0) rc=$?" | tee -a gpurun_out/pytest_gpu_tslab.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
PPO_TVF_QSLAB=0 timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-other-configs > gpurun_out/bench_tslab.json 2> gpurun_out/bench_tslab.err; echo "bench (PPO_TVF_QSLAB=0) rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/plain_bench.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu_tslab.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench.json; cat gpurun_out/bench_ref.json
timeout 120 python tools/prof_tvf.py 4096 advanced8 tslab > gpurun_out/plain_tslab.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tvf_tslab -s 2 -c 1 -o gpurun_out/r02_tslab_4096_adv -f \
    python tools/prof_tvf.py 4096 advanced8 tslab > gpurun_out/ncu_tslab.log 2>&1
tail -2 gpurun_out/ncu_tslab.log
PPO_TIME_PATH=qslab timeout 120 python tools/time_tslab.py 0 > gpurun_out/plain_qslab.log 2>&1 && \
PPO_TIME_PATH=qslab timeout 300 ncu --set full --clock-control none --import-source on -k regex:qslab -s 3 -c 1 \
    -o gpurun_out/r02_qslab_4096_adv -f python tools/time_tslab.py 0 > gpurun_out/ncu_qslab.log 2>&1
tail -2 gpurun_out/ncu_qslab.log
timeout 600 python bench.py --scaling strong --steps 30 --warmup 5 --no-cpu-baseline --no-other-configs > gpurun_out/bench_strong.json 2> gpurun_out/bench_strong.err; echo "bench strong rc=$?"
