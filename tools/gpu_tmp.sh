set -x
mkdir -p gpurun_out
rm -f gpurun_out/probe_train_ab.log
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_train_loop_gpu.py -x -q > gpurun_out/pytest_train.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_train.log
tail -5 gpurun_out/pytest_train.log
cp tak_b200/lib/libtaknative.so /tmp/A.so
for rep in 1 2; do
for v in A B; do
  if [ $v = A ]; then cp /tmp/A.so tak_b200/lib/libtaknative.so; else cp build/dev/libtaknative_B.so tak_b200/lib/libtaknative.so; fi
  echo "== variant $v (A: rolled chunk loop + bias staged per group; B: previous epilogue), fused BN-backward reduction" >> gpurun_out/probe_train_ab.log
  timeout 200 python tools/probe_train.py 4000 10 >> gpurun_out/probe_train_ab.log 2>&1
done
done
cp /tmp/A.so tak_b200/lib/libtaknative.so
echo "== variant A, TAK_TRAIN_BNB=0" >> gpurun_out/probe_train_ab.log
TAK_TRAIN_BNB=0 timeout 200 python tools/probe_train.py 4000 10 >> gpurun_out/probe_train_ab.log 2>&1
timeout 200 python tools/probe_train.py 4000 10 5 >> gpurun_out/probe_train_ab.log 2>&1
cat gpurun_out/probe_train_ab.log
