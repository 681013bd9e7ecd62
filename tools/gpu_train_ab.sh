#!/bin/bash
# Training-step A/B in one GPU call: parity tests of the training path, then probe_train in the three builds of the
# chunk (persistent trunk launches | layer per launch with the fused BN-backward reduction | round-1 two-pass route).
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_train_loop_gpu.py -x -q > gpurun_out/pytest_train.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_train.log
tail -15 gpurun_out/pytest_train.log
for mode in "1 1" "0 1" "0 0"; do
  set -- $mode
  echo "== TAK_TRAIN_PERSIST=$1 TAK_TRAIN_BNB=$2" >> gpurun_out/probe_train_ab.log
  TAK_TRAIN_PERSIST=$1 TAK_TRAIN_BNB=$2 timeout 200 python tools/probe_train.py 4000 8 >> gpurun_out/probe_train_ab.log 2>&1
  TAK_TRAIN_PERSIST=$1 TAK_TRAIN_BNB=$2 timeout 200 python tools/probe_train.py 4000 8 5 >> gpurun_out/probe_train_ab.log 2>&1
done
cat gpurun_out/probe_train_ab.log
