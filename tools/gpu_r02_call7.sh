#!/bin/bash
# 8 GPUs: 4K inference strong scaling (default / no overlap), the driver's bench line at N = 8
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29512 tools/bench_infer4k.py --batched > gpurun_out/infer4k_8gpu.json 2> gpurun_out/infer4k_8gpu.err; echo rc=$?
timeout 240 $TR --master-port 29513 tools/bench_infer4k.py --batched --no-overlap > gpurun_out/infer4k_8gpu_nooverlap.json 2>> gpurun_out/infer4k_8gpu.err; echo rc=$?
timeout 300 $TR --master-port 29514 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu.log 2>&1; echo "bench8 rc=$?"
timeout 200 python tools/bench_infer4k.py --streams > gpurun_out/infer4k_1gpu_streams.json 2>> gpurun_out/infer4k_8gpu.err; echo rc=$?
timeout 200 python tools/bench_infer4k.py > gpurun_out/infer4k_1gpu_serial.json 2>> gpurun_out/infer4k_8gpu.err; echo rc=$?
cat gpurun_out/infer4k_8gpu.json gpurun_out/infer4k_8gpu_nooverlap.json | cut -c1-200
