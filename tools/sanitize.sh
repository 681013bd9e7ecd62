#!/bin/bash
# compute-sanitizer over the small parity tests and the tensor-core self-tests (race detection /
# sanitizers).  Logs land in gpurun_out/sanitizer_*.log; a summary line per tool is appended to gpurun_out/sanitizer.txt.
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
: > gpurun_out/sanitizer.txt
run() {  # name, tool, command...
    name=$1; tool=$2; shift 2
    timeout 600 $S --tool $tool --print-limit 20 "$@" > gpurun_out/sanitizer_$name.log 2>&1
    rc=$?
    echo "$name ($tool) rc=$rc: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$name.log | tail -1)" >> gpurun_out/sanitizer.txt
}
run wgrad_memcheck memcheck ./build/wgrad_selftest 7 96 42
run conv_memcheck memcheck ./build/conv_selftest 60 6
run game_memcheck memcheck python -m pytest tests/test_game_gpu.py -x -q -k "wins or tps or error or known"
run mcts_memcheck memcheck python -m pytest tests/test_mcts_gpu.py -x -q -k "reference_mcts or stepwise"
run player_memcheck memcheck python -m pytest tests/test_player_gpu.py tests/test_pit_gpu.py -x -q -k "batch4 or 5-5-4 or equals or counts"
run perft_multi_memcheck memcheck python -m pytest tests/test_game_gpu.py -x -q -k "multi or surface or symmetrical or consistency"
run train5_memcheck memcheck python -m pytest tests/test_train_gpu.py -x -q -k "losses"
run net5_memcheck memcheck python -m pytest tests/test_net_gpu.py -x -q -k "5"
run game_racecheck racecheck python -m pytest tests/test_game_gpu.py -x -q -k "wins or tps or error"
cat gpurun_out/sanitizer.txt
