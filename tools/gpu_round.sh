#!/bin/bash
# One GPU call: parity tests, bench, ncu launch list, one full capture of the dominant kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 ./build/conv_selftest 5328 6 > gpurun_out/conv_selftest.log 2>&1; tail -4 gpurun_out/conv_selftest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --rollouts 20 --replicas 1 --games 5328 > gpurun_out/b_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc3 -s 2 -c 2 -o gpurun_out/conv_tc3_full -f python tools/probe_forward.py 5328 2 > gpurun_out/ncu_full.log 2>&1; tail -n 3 gpurun_out/ncu_full.log
timeout 300 ./build/wgrad_selftest 667 128 42 > gpurun_out/wgrad_selftest.log 2>&1; tail -3 gpurun_out/wgrad_selftest.log
timeout 300 python tools/probe_train.py 4000 5 > gpurun_out/probe_train.log 2>&1; tail -1 gpurun_out/probe_train.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_train.csv python tools/probe_train.py 4000 1 > gpurun_out/train_ncu.log 2>&1
