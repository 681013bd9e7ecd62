#!/bin/bash
# Round-2 evidence in one GPU call: parity tests, conv self-test, bench (+ reference arm), perft probe, launch list of a
# short self-play run.  Outputs land in gpurun_out/ (copy what should be judged into profiles/).
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 ./build/conv_selftest 5328 6 > gpurun_out/conv_selftest.log 2>&1; tail -4 gpurun_out/conv_selftest.log
timeout 300 ./build/conv_selftest 5180 6 8 1 > gpurun_out/conv_selftest_padfree.log 2>&1; tail -4 gpurun_out/conv_selftest_padfree.log
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -n 3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 python tools/probe_perft.py 5 6 5 > gpurun_out/probe_perft6.log 2>&1; tail -1 gpurun_out/probe_perft6.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --rollouts 100 --replicas 1 --games 5328 > gpurun_out/b_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt 2>&1; head -12 gpurun_out/launch_summary.txt
