#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"merge_kernel|loss_iou_kernel|split_norm_patch32|split_hwc_kernel|pr_hist" -c 6 -f -o gpurun_out/prof_hbm_r01 python tools/hbm_kernels.py > gpurun_out/ncu_hbm.log 2>&1; echo "ncu hbm rc=$?"
# conv layers of the 3rd plan run (29 conv launches per run): 64->64@512, 256->256@128, 512->512@64, 768->512@64, ConvT 128->32, dec1+head
timeout 900 ncu --set full --clock-control none -k regex:conv_halo -s 46 -c 23 -f -o gpurun_out/prof_conv_r01 python tools/layer_times.py 13 > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
