#!/bin/bash
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_final.log 2>&1; echo "bench2 rc=$?"
timeout 200 $TR --master-port 29515 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_2gpu_ref.log 2>&1; echo "ref rc=$?"
timeout 200 $TR --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_gpu_check_2gpu.log 2>&1; echo "check rc=$?"
tail -2 gpurun_out/bench_2gpu_final.log | cut -c1-400; tail -4 gpurun_out/multi_gpu_check_2gpu.log
