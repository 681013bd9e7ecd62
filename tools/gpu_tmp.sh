set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_comm_gpu.py -x -q > gpurun_out/test_comm_2gpu.log 2>&1; tail -4 gpurun_out/test_comm_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -2 gpurun_out/bench_n2.err; cut -c1-400 gpurun_out/bench_n2.json
