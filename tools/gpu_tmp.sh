set -x
mkdir -p gpurun_out
timeout 120 ./build/wgrad_selftest 667 128 42 2>&1 | tail -3
timeout 120 ./build/wgrad_selftest 37 96 48 2>&1 | tail -3
for rep in 1 2; do
for pf in 1 0; do
echo "== TAK_TRAIN_PREFETCH=$pf"
TAK_TRAIN_PREFETCH=$pf timeout 200 python tools/probe_train.py 4000 10 2>&1 | cut -c1-110
done
done
timeout 200 python tools/probe_train.py 4000 10 5 2>&1 | cut -c1-110
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_net_gpu.py -x -q 2>&1 | tail -3
timeout 300 ./build/conv_selftest 5180 6 8 1 2>&1 | tail -4
