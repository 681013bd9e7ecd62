set -x
mkdir -p gpurun_out
cp tak_b200/lib/libtaknative.so /tmp/A.so
for rep in 1 2; do
for v in A P; do
  if [ $v = A ]; then cp /tmp/A.so tak_b200/lib/libtaknative.so; else cp build/dev/libtaknative_prev.so tak_b200/lib/libtaknative.so; fi
  echo "== variant $v (A: stage-contiguous X image, 2 bulk copies per stage; P: 17 copies per stage)"
  timeout 200 python tools/probe_forward5.py 4096 2>&1 | tail -1 | cut -c1-200
  timeout 200 python tools/probe_forward5.py 8880 2>&1 | tail -1 | cut -c1-200
done
done
cp /tmp/A.so tak_b200/lib/libtaknative.so
timeout 200 python tools/probe_train.py 4000 10 5 2>&1 | cut -c1-110
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_net_gpu.py tests/test_mcts_gpu.py -x -q 2>&1 | tail -3
