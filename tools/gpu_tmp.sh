set -x
mkdir -p gpurun_out
rm -f gpurun_out/probe_train_ab.log
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_train_loop_gpu.py -x -q > gpurun_out/pytest_train.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_train.log
tail -5 gpurun_out/pytest_train.log
for rep in 1 2; do
for mode in "1 1" "0 1" "1 0" "0 0"; do
  set -- $mode
  echo "== TAK_TRAIN_PREFETCH=$1 TAK_TRAIN_RSTREAM=$2" >> gpurun_out/probe_train_ab.log
  TAK_TRAIN_PREFETCH=$1 TAK_TRAIN_RSTREAM=$2 timeout 200 python tools/probe_train.py 4000 10 2>&1 | cut -c1-110 >> gpurun_out/probe_train_ab.log
done
done
timeout 200 python tools/probe_train.py 4000 10 5 2>&1 | cut -c1-110 >> gpurun_out/probe_train_ab.log
cat gpurun_out/probe_train_ab.log
