#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_models.py tests/test_gpu_pipeline.py -m gpu -q -x > gpurun_out/t_conv.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/t_conv.log
SNB_CONV_MODE=3 timeout 300 python tools/layer_times.py 13 > gpurun_out/layers_pool.log 2>&1; cat gpurun_out/layers_pool.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pool.log 2>&1; tail -1 gpurun_out/bench_pool.log | cut -c1-250
