#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x > gpurun_out/t_pipe.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_pipe.log
for B in 1 2 4 7 13; do SNB_CONV_MODE=3 timeout 300 python tools/layer_times.py $B 2>&1 | head -1; done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_graph.log 2>&1; tail -1 gpurun_out/bench_graph.log
