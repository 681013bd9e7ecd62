#!/bin/bash
# compute-sanitizer over the kernels that are new or rewritten in round 2: perft expansion (k_perft_moves / k_perft_apply),
# device playouts, the fused search loop (k_mcts_step + tower with a device-side board count), the instant-win scan over
# all moves, sampled picks / Dirichlet noise.  Logs: gpurun_out/sanitizer2_*.log, summary: gpurun_out/sanitizer2.txt
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
: > gpurun_out/sanitizer2.txt
run() {  # name, tool, command...
    name=$1; tool=$2; shift 2
    timeout 900 $S --tool $tool --print-limit 20 "$@" > gpurun_out/sanitizer2_$name.log 2>&1
    rc=$?
    echo "$name ($tool) rc=$rc: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer2_$name.log | tail -1) | $(grep -E 'passed|failed' gpurun_out/sanitizer2_$name.log | tail -1)" >> gpurun_out/sanitizer2.txt
}
run perft_memcheck memcheck python -m pytest tests/test_game_gpu.py -x -q -k "perft_known or deep_positions or taller or device_playouts"
run perft_racecheck racecheck python -m pytest tests/test_game_gpu.py -x -q -k "deep_positions or taller"
run perft_initcheck initcheck python -m pytest tests/test_game_gpu.py -x -q -k "deep_positions"
run fused_memcheck memcheck python -m pytest tests/test_mcts_gpu.py -x -q -k "reference_mcts or dummy_net_bit_exact or fused"
run selfplay_memcheck memcheck python -m pytest tests/test_selfplay_gpu.py -x -q -k "instant_win or records"
run noise_memcheck memcheck python -m pytest tests/test_noise_sampling_gpu.py -x -q -k "network_priors"
run fused_racecheck racecheck python -m pytest tests/test_selfplay_gpu.py -x -q -k "matches_reference_loop and 6-6"
cat gpurun_out/sanitizer2.txt
