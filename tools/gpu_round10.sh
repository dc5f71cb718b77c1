#!/bin/bash
mkdir -p gpurun_out
SNB_CONV_MODE=4 timeout 200 python tools/conv_debug.py > gpurun_out/conv_debug_m4.log 2>&1; echo "conv_debug rc=$?"; tail -12 gpurun_out/conv_debug_m4.log
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_models.py tests/test_gpu_pipeline.py -m gpu -q -x > gpurun_out/t_conv.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/t_conv.log
for M in 3 4; do SNB_CONV_MODE=$M timeout 300 python tools/layer_times.py 13 > gpurun_out/layers_m$M.log 2>&1; done
paste -d'|' gpurun_out/layers_m3.log <(cut -c50- gpurun_out/layers_m4.log)
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pairs.log 2>&1; tail -1 gpurun_out/bench_pairs.log | cut -c1-250
