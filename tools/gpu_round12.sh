#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_slicer.py tests/test_gpu_pipeline.py -m gpu -q 2>&1 | tail -2
timeout 300 python tools/hbm_kernels.py 2>&1 | tee gpurun_out/hbm_kernels.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:merge_kernel -s 2 -c 1 -f -o gpurun_out/prof_merge_r01 python tools/hbm_kernels.py > gpurun_out/ncu_merge.log 2>&1; echo "ncu merge rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:loss_iou_kernel -s 2 -c 1 -f -o gpurun_out/prof_loss_r01 python tools/hbm_kernels.py > gpurun_out/ncu_loss.log 2>&1; echo "ncu loss rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r12.log 2>&1; tail -1 gpurun_out/bench_r12.log | cut -c1-220
du -sh gpurun_out
