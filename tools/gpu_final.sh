#!/bin/bash
# Round-end measurement pass on one B200: tests, smoke, headline bench (+ reference arm), secondary workloads,
# ncu launch list of the bench command and ncu --set full captures of the HBM kernels.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/final_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_final.log 2>&1; tail -1 gpurun_out/bench_final.log | cut -c1-250
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.log 2>&1; tail -1 gpurun_out/bench_reference.log | cut -c1-250
for m in unet11 zf_unet fcdensenet67 linknet34; do
  timeout 600 python bench.py --model $m --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$m.log 2>&1
  tail -1 gpurun_out/bench_$m.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['metric'], round(d['value'],1), 'Mpx/s', round(d['ms_per_step'],1), 'ms', 'conv TF/s', round(d['roofline']['achieved'],1))"
done
timeout 600 python bench.py --tta --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tta.log 2>&1; tail -1 gpurun_out/bench_tta.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tta', round(d['value'],1), 'Mpx/s', round(d['ms_per_step'],1), 'ms')"
timeout 600 python bench.py --precision tf32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tf32.log 2>&1; tail -1 gpurun_out/bench_tf32.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tf32', round(d['value'],1), 'Mpx/s', round(d['ms_per_step'],1), 'ms')"
timeout 300 python tools/hbm_kernels.py > gpurun_out/hbm_kernels_final.log 2>&1; cat gpurun_out/hbm_kernels_final.log
(timeout 300 python tools/train_step_bench.py --batch 8; timeout 300 python tools/train_step_bench.py --batch 32; timeout 300 python tools/train_step_bench.py --batch 8 --cpu) > gpurun_out/train_step_final.log 2>&1; cat gpurun_out/train_step_final.log
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r01_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_final.log 2>&1; echo "ncu launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"merge_f32c1|loss_iou_kernel|split_norm_patch32|confusion_kernel" -s 4 -c 4 -f -o gpurun_out/prof_hbm_final python tools/hbm_kernels.py > gpurun_out/ncu_hbm_final.log 2>&1; echo "ncu hbm rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:"conv_scatter" -s 20 -c 3 -f -o gpurun_out/prof_scatter_final python tools/layer_times.py 44 224 fcdensenet67 > gpurun_out/ncu_scatter_final.log 2>&1; echo "ncu scatter rc=$?"
ls -la gpurun_out/*final* | head -20
