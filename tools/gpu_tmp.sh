set -x
mkdir -p gpurun_out
timeout 200 python tools/probe_train.py 4000 10 2>&1 | cut -c1-110
timeout 200 python tools/probe_train.py 4000 10 5 2>&1 | cut -c1-110
timeout 600 python -m pytest tests/test_train_gpu.py -x -q 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 420 --csv --log-file gpurun_out/launches_train.csv python tools/probe_train.py 4000 2 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/launches_train.csv > gpurun_out/launch_summary_train.txt 2>&1; head -8 gpurun_out/launch_summary_train.txt
