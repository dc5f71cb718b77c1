#!/bin/bash
# first GPU pass: HBM kernels, conv diagnostics, conv tests, model tests -- each under its own timeout
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_slicer.py tests/test_gpu_reduce.py -m gpu -q > gpurun_out/t_hbm.log 2>&1; echo "hbm rc=$?"
timeout 300 python tools/conv_debug.py > gpurun_out/conv_debug.log 2>&1; echo "conv_debug rc=$?"
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q > gpurun_out/t_conv.log 2>&1; echo "conv rc=$?"
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q > gpurun_out/t_models.log 2>&1; echo "models rc=$?"
tail -5 gpurun_out/t_hbm.log; tail -25 gpurun_out/conv_debug.log; tail -5 gpurun_out/t_conv.log; tail -5 gpurun_out/t_models.log
