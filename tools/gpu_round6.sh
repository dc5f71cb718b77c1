#!/bin/bash
mkdir -p gpurun_out
for B in 13 26 39 57 169; do
  timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --batch $B > gpurun_out/bench_B$B.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/bench_B$B.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('batch $B', 'Mpx/s', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'conv TF', round(d['roofline']['achieved'],1), 'conv ms', round(d['roofline']['conv_ms_per_step'],2), 'pool ms', round(d['roofline']['pool_ms_per_step'],2), 'e2e', round(d['e2e']['value'],1))
else:
    print(open('gpurun_out/bench_B$B.log').read()[-1500:])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 460 --csv --log-file gpurun_out/launches_r01b.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
