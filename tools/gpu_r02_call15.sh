#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_8gpu_final.log 2>&1; echo "bench8 rc=$?"
tail -1 gpurun_out/bench_8gpu_final.log | cut -c1-300
