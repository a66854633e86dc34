#!/bin/bash
# Round-2 GPU call 4 (2 GPUs): host overhead of a rank's inference step, strip exactness, 4K inference with / without overlap, bench at 2
set -x
mkdir -p gpurun_out
python tools/strip_host_overhead.py --ranks 8 > gpurun_out/host_overhead_r8.json 2> gpurun_out/host_overhead.err; echo rc=$?
python tools/strip_host_overhead.py --ranks 2 > gpurun_out/host_overhead_r2.json 2>> gpurun_out/host_overhead.err; echo rc=$?
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_gpu_check_2gpu.log 2>&1; echo "check rc=$?"
$TR --master-port 29512 tools/bench_infer4k.py --batched > gpurun_out/infer4k_2gpu_overlap.json 2> gpurun_out/infer4k_2gpu.err; echo rc=$?
$TR --master-port 29513 tools/bench_infer4k.py --batched --no-overlap > gpurun_out/infer4k_2gpu_nooverlap.json 2>> gpurun_out/infer4k_2gpu.err; echo rc=$?
python tools/bench_infer4k.py > gpurun_out/infer4k_1gpu.json 2>> gpurun_out/infer4k_2gpu.err; echo rc=$?
$TR --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; echo "bench2 rc=$?"
tail -3 gpurun_out/multi_gpu_check_2gpu.log
