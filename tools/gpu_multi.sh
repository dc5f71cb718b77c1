#!/bin/bash
# Multi-GPU measurement pass on one box with N GPUs (gpurun --gpus N -- bash tools/gpu_multi.sh N): NCCL parity tests, the
# weak-scaling bench (one image per rank per step), one image tile-sharded (strong scaling), BASELINE configs[3] on a
# shortened image list with --verify (masks and counts against the single-rank run).  Outputs: gpurun_out/r02_*_nN.*
N=${1:-2}
R=r02
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
nvidia-smi -L | head -$N
timeout 600 python -m pytest tests/test_gpu_multirank.py -q 2>&1 | tail -2
timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/${R}_bench_n$N.log 2>&1; grep '^{' gpurun_out/${R}_bench_n$N.log | tail -1 > gpurun_out/${R}_bench_n$N.json; cut -c1-330 gpurun_out/${R}_bench_n$N.json
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 6 --warmup 3 --shard tile --no-secondary > gpurun_out/${R}_bench_tile_n$N.log 2>&1; grep '^{' gpurun_out/${R}_bench_tile_n$N.log | tail -1 > gpurun_out/${R}_bench_tile_n$N.json; cut -c1-330 gpurun_out/${R}_bench_tile_n$N.json
timeout 900 $TR --master-port 29513 tools/run_config3.py --images $((3 * N)) --verify > gpurun_out/${R}_config3_n$N.log 2>&1; tail -4 gpurun_out/${R}_config3_n$N.log
timeout 900 $TR --master-port 29514 tools/run_config3.py --images $N --shard tile --verify > gpurun_out/${R}_config3_tile_n$N.log 2>&1; tail -3 gpurun_out/${R}_config3_tile_n$N.log
