#!/bin/bash
# compute-sanitizer over the small-geometry GPU tests (one B200).  memcheck: every kernel family; racecheck: the kernels that
# exchange data through shared memory with generic loads / stores (reductions, BatchNorm, slicer, gather helpers) -- racecheck
# does not model TMA / mbarrier / tcgen05 traffic, so the tensor-core kernels are covered by memcheck and the parity tests.
# Summaries: gpurun_out/sanitizer_*.txt (copy the tails into profiles/).
mkdir -p gpurun_out
export SNB_SANITIZER=1
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # name, tool, timeout, pytest args...
  local name=$1 tool=$2 to=$3; shift 3
  timeout $to $CS --tool $tool --error-exitcode 99 --print-limit 20 python -m pytest "$@" -q -x -p no:cacheprovider > gpurun_out/sanitizer_$name.txt 2>&1
  echo "== $name ($tool) rc=$? : $(grep -E 'passed|failed|error' gpurun_out/sanitizer_$name.txt | tail -1) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$name.txt | tail -1)"
}
run mem_conv memcheck 900 tests/test_gpu_conv.py -k "not 512"
run mem_wgrad memcheck 600 tests/test_gpu_wgrad.py tests/test_gpu_conv_grad.py -k "not bench and not full"
run mem_slicer_reduce memcheck 600 tests/test_gpu_slicer.py tests/test_gpu_reduce.py tests/test_gpu_abn.py -k "not full_size"
run race_reduce racecheck 600 tests/test_gpu_reduce.py tests/test_gpu_abn.py -k "not full_size"
run race_slicer racecheck 600 tests/test_gpu_slicer.py -k "not full_size"
run mem_models memcheck 900 tests/test_gpu_models.py -k "logits_against_reference_vectors or linknet34_against_reference_vectors or zf_unet_against or fused_train_step or fcdensenet67_against"
