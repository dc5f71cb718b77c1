#!/bin/bash
# Round-end measurement pass on one B200: tests (with parity margins), smoke, headline bench (+ reference arm), secondary
# workloads, HBM kernels, training step, role profiles (UNet16 and FCDenseNet67; needs tools/build_rev.py --profile), ncu launch
# list of the bench command and one ncu metrics pass over the conv kernels at the bench's tiles per launch.
# Outputs under gpurun_out/r02_* (copy summaries to profiles/; tools/fill_docs.py fills the documents from them).
R=r02
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/${R}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/${R}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${R}_pytest_gpu.log
grep -E "^margin|max \|p - oracle\||rel-L2|margin" gpurun_out/${R}_pytest_gpu.log > gpurun_out/${R}_parity_margins.txt; wc -l gpurun_out/${R}_parity_margins.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/${R}_bench_final.log 2>&1; tail -1 gpurun_out/${R}_bench_final.log > gpurun_out/${R}_bench_final.json; cut -c1-260 gpurun_out/${R}_bench_final.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.log 2>&1; tail -1 gpurun_out/${R}_bench_reference.log > gpurun_out/${R}_bench_reference.json; cut -c1-260 gpurun_out/${R}_bench_reference.json
: > gpurun_out/${R}_secondary.txt
for m in unet11 zf_unet fcdensenet67 linknet34; do
  timeout 600 python bench.py --model $m --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${R}_bench_$m.log 2>&1
  tail -1 gpurun_out/${R}_bench_$m.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['metric'], '|', round(d['value'],1), 'Mpx/s', round(d['ms_per_step'],1), 'ms', '| conv TF/s', round(d['roofline']['achieved'],1), 'frac', round(d['roofline']['frac'],3), '| tile_batch', d['run']['tile_batch'])" | tee -a gpurun_out/${R}_secondary.txt
done
timeout 600 python bench.py --tta --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${R}_bench_tta.log 2>&1; tail -1 gpurun_out/${R}_bench_tta.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('UNet16 with D4 TTA (8 views) |', round(d['value'],1), 'Mpx/s', round(d['ms_per_step'],1), 'ms')" | tee -a gpurun_out/${R}_secondary.txt
timeout 600 python bench.py --precision tf32 --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${R}_bench_tf32.log 2>&1; tail -1 gpurun_out/${R}_bench_tf32.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('UNet16 tf32 mode |', round(d['value'],1), 'Mpx/s', round(d['ms_per_step'],1), 'ms')" | tee -a gpurun_out/${R}_secondary.txt
timeout 300 python tools/hbm_kernels.py > gpurun_out/${R}_hbm_kernels.txt 2>&1; cut -c1-110 gpurun_out/${R}_hbm_kernels.txt
(timeout 300 python tools/train_step_bench.py --batch 8 --steps 20; timeout 300 python tools/train_step_bench.py --batch 32 --steps 10) > gpurun_out/${R}_train_step.txt 2>&1; cat gpurun_out/${R}_train_step.txt
SNB_B200_LIB=tools/ab/libsnb_b200_profile.so timeout 200 python tools/conv_wait_profile.py 13 > gpurun_out/${R}_conv_roles.txt 2>&1; tail -3 gpurun_out/${R}_conv_roles.txt
SNB_B200_LIB=tools/ab/libsnb_b200_profile.so timeout 200 python tools/conv_wait_profile.py 44 224 fcdensenet67 > gpurun_out/${R}_fcd_roles.txt 2>&1; tail -2 gpurun_out/${R}_fcd_roles.txt | cut -c1-200
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 106 -c 530 --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${R}_ncu_bench.log 2>&1; echo "ncu launch list rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second,launch__registers_per_thread --clock-control none -k regex:"conv_halo|conv_first|conv_igemm" -s 48 -c 24 -f -o gpurun_out/${R}_conv85m python tools/layer_times.py 85 > gpurun_out/${R}_ncu_conv85m.log 2>&1; echo "ncu conv metrics rc=$?"
ls -la gpurun_out/${R}_* | head -40
