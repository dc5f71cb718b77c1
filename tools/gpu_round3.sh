#!/bin/bash
mkdir -p gpurun_out
for M in 1 2; do
  SNB_CONV_MODE=$M timeout 300 python tools/conv_debug.py > gpurun_out/conv_debug_m$M.log 2>&1; echo "conv_debug mode $M rc=$?"
  tail -12 gpurun_out/conv_debug_m$M.log
done
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_models.py tests/test_gpu_pipeline.py -m gpu -q > gpurun_out/t_conv.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/t_conv.log
for M in 0 1 2; do
  SNB_CONV_MODE=$M timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_m$M.log 2>&1; echo "bench mode $M rc=$?"
  python - <<PY
import json
l=[x for x in open('gpurun_out/bench_m$M.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('mode $M', 'Mpx/s', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'conv TF', round(d['roofline']['achieved'],1), 'conv ms', round(d['roofline']['conv_ms_per_step'],2))
else:
    print(open('gpurun_out/bench_m$M.log').read()[-2000:])
PY
done
