#!/bin/bash
# compute-sanitizer over the training path as it is after the second half of round 2: dgrad launches with the fused
# BatchNorm-backward sums, the rolled training epilogue, the tiled-TMA wgrad kernel (UTMALDG with out-of-bounds zero
# fill), k_bn_bwd_apply<MASK = false>, the stage-contiguous FC X image.  Summary: gpurun_out/sanitizer2b.txt
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
: > gpurun_out/sanitizer2b.txt
run() {  # name, tool, command...
    name=$1; tool=$2; shift 2
    timeout 900 $S --tool $tool --print-limit 20 "$@" > gpurun_out/sanitizer2b_$name.log 2>&1
    rc=$?
    echo "$name ($tool) rc=$rc: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer2b_$name.log | tail -1) | $(grep -E 'passed|failed|SELFTEST' gpurun_out/sanitizer2b_$name.log | tail -1)" >> gpurun_out/sanitizer2b.txt
}
run wgrad_memcheck memcheck ./build/wgrad_selftest 37 128 42
run wgrad96_memcheck memcheck ./build/wgrad_selftest 19 96 48
run wgrad_initcheck initcheck ./build/wgrad_selftest 37 128 42
run train_memcheck memcheck python -m pytest tests/test_train_gpu.py -x -q -k "losses_match or gradients_match"
run train_initcheck initcheck python -m pytest tests/test_train_gpu.py -x -q -k "losses_match and net6"
cat gpurun_out/sanitizer2b.txt
