#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_4gpu_final.log 2>&1; echo "bench4 rc=$?"
tail -1 gpurun_out/bench_4gpu_final.log | cut -c1-300
