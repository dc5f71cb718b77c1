#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_conv.py -m gpu -q -x 2>&1 | tail -3
for A in 4 8; do SNB_A_STAGES=$A timeout 300 python tools/layer_times.py 13 2>&1 | grep -E "total|cout=  32|cout=  64" ; done
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r13.log 2>&1; tail -1 gpurun_out/bench_r13.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms',round(d['ms_per_step'],2),'clocks',d['clocks'])"
