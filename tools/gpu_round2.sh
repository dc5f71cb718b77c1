#!/bin/bash
# full GPU test suite, smoke, first bench line, ncu launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/t_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
for B in 13 ; do
  timeout 900 python bench.py --steps 3 --warmup 3 --batch $B > gpurun_out/bench_b$B.log 2>&1; echo "bench b$B rc=$?"; tail -2 gpurun_out/bench_b$B.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 460 --csv --log-file gpurun_out/launches_r01.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/ncu_bench.log
