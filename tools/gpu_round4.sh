#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slicer.py -m gpu -q > gpurun_out/t_slicer.log 2>&1; echo "slicer rc=$?"; tail -3 gpurun_out/t_slicer.log
for M in 0 1 2; do
  SNB_CONV_MODE=$M timeout 300 python tools/layer_times.py 13 > gpurun_out/layers_m$M.log 2>&1; echo "layers mode $M rc=$?"
done
paste -d'|' gpurun_out/layers_m0.log <(cut -c52- gpurun_out/layers_m1.log) <(cut -c52- gpurun_out/layers_m2.log)
SNB_CONV_MODE=2 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_m2b.log 2>&1; tail -1 gpurun_out/bench_m2b.log | cut -c1-400
